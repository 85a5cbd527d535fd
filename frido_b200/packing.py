"""Weight packing through the C ABI (csrc/pack.cu; SURVEY.md §8b `pack_weights`).

Checkpoint tensors keep the reference's layouts (Conv2d OIHW, Linear [out,in]); the engines read K-major rows
[C_out][tap][C_in] with fused operands concatenated.  Every function here turns device tensors into a packed device tensor
with native launches only (`frido_pack_permute3`, `frido_pack_conv_weight`, `frido_fold_self_attention`, `frido_vec_add`) on
the current stream - no ATen kernels - so the same packing is reachable from a non-Python host.  There is no CPU path:
for CPU tensors (plans are host bookkeeping and may be BUILT, never run, without a device) the destination is allocated
with the right shape and left unwritten - nothing is computed on the host.
"""
import torch

from . import _lib as L


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _f32(t):
    t = t.detach()
    if t.dtype != torch.float32:
        raise L.FridoError("weight packing expects fp32 checkpoint tensors")
    return t


def _permute3(src, s, dst, d_off, d, n):
    if not src.is_cuda:   # shape bookkeeping only (see the module docstring)
        return
    L.check(L.lib().frido_pack_permute3(src.data_ptr(), s[0], s[1], s[2], dst.data_ptr() + 4 * d_off, d[0], d[1], d[2],
                                        n[0], n[1], n[2], _stream()), "pack_permute3")


def conv_weight(w, out=None, row0=0, col0=0):
    """OIHW -> rows [O][kh*kw][I] (K-major for the implicit GEMM), optionally into a block of `out` at (row0, col0)."""
    w = _f32(w)
    O, I, KH, KW = w.shape
    w = w.contiguous()
    if out is None:
        out = torch.empty(O, KH * KW * I, dtype=torch.float32, device=w.device)
    ld = out.stride(0)
    if not w.is_cuda:
        return out
    L.check(L.lib().frido_pack_conv_weight(w.data_ptr(), O, I, KH, KW, out.data_ptr() + 4 * (row0 * ld + col0), ld, _stream()),
            "pack_conv_weight")
    return out


def place(m, out, row0=0, col0=0, transpose=False, row_step=1):
    """out[row0 + i*row_step, col0 + j] = m[i, j] (or m[j, i] with transpose): concatenation and interleaving by strides."""
    m = _f32(m)
    if m.dim() == 1:
        m = m.view(1, -1)
    R, Cn = (m.shape[1], m.shape[0]) if transpose else (m.shape[0], m.shape[1])
    s_r, s_c = (m.stride(1), m.stride(0)) if transpose else (m.stride(0), m.stride(1))
    ld = out.stride(0) if out.dim() == 2 else out.numel()
    _permute3(m, (0, s_r, s_c), out, row0 * ld + col0, (0, ld * row_step, 1), (1, R, Cn))
    return out


def cat_rows(mats, device=None):
    """torch.cat(mats, 0) of 2-D (or 1-D) fp32 tensors, as native strided copies."""
    mats = [_f32(m) for m in mats]
    if mats[0].dim() == 1:
        n = sum(m.numel() for m in mats)
        out = torch.empty(n, dtype=torch.float32, device=mats[0].device)
        off = 0
        for m in mats:
            place(m, out.view(1, -1), 0, off)
            off += m.numel()
        return out
    K = mats[0].shape[1]
    out = torch.empty(sum(m.shape[0] for m in mats), K, dtype=torch.float32, device=mats[0].device)
    r = 0
    for m in mats:
        place(m, out, r, 0)
        r += m.shape[0]
    return out


def conv_rows(convs_w):
    """Several conv weights with the same input, concatenated on C_out (SPADE gamma | beta): [sum O][tap][I]."""
    ws = [_f32(w) for w in convs_w]
    K = ws[0].shape[1] * ws[0].shape[2] * ws[0].shape[3]
    out = torch.empty(sum(w.shape[0] for w in ws), K, dtype=torch.float32, device=ws[0].device)
    r = 0
    for w in ws:
        conv_weight(w, out, r, 0)
        r += w.shape[0]
    return out


def conv_plus_side(w3, w1):
    """[O][9 I | I_x]: a 3x3 conv's rows with the 1x1 skip_connection's weights appended (pyunet.py:248,299 as extra K steps)."""
    w3, w1 = _f32(w3), _f32(w1)
    O, I = w3.shape[0], w3.shape[1]
    T = w3.shape[2] * w3.shape[3]
    Ix = w1.shape[1]
    out = torch.empty(O, T * I + Ix, dtype=torch.float32, device=w3.device)
    conv_weight(w3, out, 0, 0)
    place(w1.reshape(O, Ix), out, 0, T * I)
    return out


def interleave_rows(a, b):
    """Rows (a_0, b_0, a_1, b_1, ...): GEGLU value / gate pairs next to each other (attention.py:42-44)."""
    a, b = _f32(a), _f32(b)
    one_d = a.dim() == 1
    if one_d:
        a, b = a.view(-1, 1), b.view(-1, 1)
    out = torch.empty(2 * a.shape[0], a.shape[1], dtype=torch.float32, device=a.device)
    place(a, out, 0, 0, row_step=2)
    place(b, out, 1, 0, row_step=2)
    return out.view(-1) if one_d else out


def transpose(m):
    m = _f32(m)
    out = torch.empty(m.shape[1], m.shape[0], dtype=torch.float32, device=m.device)
    return place(m, out, transpose=True)


def copy(t):
    """Packed copy of a tensor (vectors, [C,C] views of 1x1 convs): detached, dense, same values."""
    t = _f32(t)
    t2 = t.reshape(1, -1) if t.dim() != 2 else t
    out = torch.empty(t2.shape, dtype=torch.float32, device=t.device)
    place(t2, out)
    return out.view(t.shape)


def vec_add(a, b):
    a, b = _f32(a), _f32(b)
    a, b = a.contiguous(), b.contiguous()
    out = torch.empty_like(a)
    if not a.is_cuda:
        return out
    L.check(L.lib().frido_vec_add(a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(), _stream()), "vec_add")
    return out


def fold_self_attention(wq, wk, wv, wo):
    """(Wk^T Wq, Wo Wv) with fp64 products and sums, rounded once (attention.py:172-191 re-associated; see unet.py)."""
    wq, wk, wv, wo = (_f32(w) for w in (wq, wk, wv, wo))
    Cd = wq.shape[0]
    assert all(tuple(w.shape) == (Cd, Cd) for w in (wq, wk, wv, wo)), "self-attention folds need square projections"
    wq, wk, wv, wo = (w.contiguous() for w in (wq, wk, wv, wo))
    a = torch.empty(Cd, Cd, dtype=torch.float32, device=wq.device)
    v = torch.empty_like(a)
    if not wq.is_cuda:
        return a, v
    L.check(L.lib().frido_fold_self_attention(wq.data_ptr(), wk.data_ptr(), wv.data_ptr(), wo.data_ptr(), Cd, a.data_ptr(),
                                              v.data_ptr(), _stream()), "fold_self_attention")
    return a, v
