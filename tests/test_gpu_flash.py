"""GPU parity of the streaming-softmax tcgen05 attention kernel (csrc/attn_flash.cu, frido_attn_flash) against an fp64
PyTorch statement of  res + bias + softmax(scale q k^T) v  (attention.py:170-193 with one head of C channels).

Tolerance: BF16x3 products are good to ~2^-16 relative per term, P is split the same way and everything accumulates in
fp32; the scores here are O(1..10), so |err| <= 2e-4 * max|out| (measured ~2e-5) is the bound written in the assert."""
import ctypes as C
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _pair(t):
    from frido_b200 import _lib as L
    hi = torch.empty(t.shape, dtype=torch.bfloat16, device=t.device)
    lo = torch.empty_like(hi)
    s = torch.cuda.current_stream().cuda_stream
    L.check(L.lib().frido_split_bf16(t.data_ptr(), hi.data_ptr(), lo.data_ptr(), t.numel(), s), "split")
    return hi, lo


def _flash(q, k, v, scale, bias, res, layout):
    """layout 'cbn': V^T stored [C][B*N] (what the UNet plan produces); 'bcn': [B][C][N]."""
    from frido_b200 import _lib as L
    B, N, Cd = q.shape
    if layout == "cbn":
        vt = v.permute(2, 0, 1).contiguous()  # [C, B, N]
        vt_sb, vt_ld = N, B * N
    else:
        vt = v.permute(0, 2, 1).contiguous()  # [B, C, N]
        vt_sb, vt_ld = Cd * N, N
    qp, kp, vp = _pair(q.contiguous()), _pair(k.contiguous()), _pair(vt)
    out = torch.full((B, N, Cd), float("nan"), device=q.device)
    p = L.FlashParams()
    p.q_hi, p.q_lo, p.q_sb, p.q_ld = qp[0].data_ptr(), qp[1].data_ptr(), N * Cd, Cd
    p.k_hi, p.k_lo, p.k_sb, p.k_ld = kp[0].data_ptr(), kp[1].data_ptr(), N * Cd, Cd
    p.vt_hi, p.vt_lo, p.vt_sb, p.vt_ld = vp[0].data_ptr(), vp[1].data_ptr(), vt_sb, vt_ld
    p.B, p.N, p.C, p.scale = B, N, Cd, scale
    p.bias = bias.data_ptr() if bias is not None else None
    if res is not None:
        p.res, p.r_sb, p.r_ld = res.data_ptr(), N * Cd, Cd
    p.out, p.o_sb, p.o_ld = out.data_ptr(), N * Cd, Cd
    L.check(L.lib().frido_attn_flash(C.byref(p), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "attn_flash")
    torch.cuda.synchronize()
    return out


def _ref(q, k, v, scale, bias, res):
    q, k, v = q.double().cpu(), k.double().cpu(), v.double().cpu()
    o = torch.softmax(scale * q @ k.transpose(1, 2), dim=-1) @ v
    if bias is not None:
        o = o + bias.double().cpu()
    if res is not None:
        o = o + res.double().cpu()
    return o


CASES = [
    # B, N, C, forced output split (0 = launcher's choice), V layout
    (1, 128, 64, 0, "cbn"),      # one key block, smallest head
    (2, 256, 576, 0, "cbn"),     # the 16x16 level of the layout2img UNet: 3 output slices of 192
    (2, 256, 576, 2, "cbn"),     # DV = 288 -> two 144-wide MMAs per key step
    (1, 1024, 384, 0, "cbn"),    # the 32x32 level: full 384-wide accumulator, 8 key blocks
    (2, 1024, 384, 2, "bcn"),    # same, sliced, batched V layout
    (1, 512, 960, 0, "bcn"),     # 512^2 config's 16x16 level
    (3, 384, 96, 0, "cbn"),      # odd block count, narrow head
    (1, 4096, 128, 0, "bcn"),    # long sequence (decoder-like)
]


@pytest.fixture(params=["single", "pair"])
def pair_mode(request):
    """single-CTA kernel / CTA-pair kernel (tcgen05.mma.cta_group::2; shapes with an odd number of query tiles fall back)."""
    old = os.environ.get("FRIDO_FLASH_PAIR")
    os.environ["FRIDO_FLASH_PAIR"] = "1" if request.param == "pair" else "0"
    yield request.param
    if old is None:
        os.environ.pop("FRIDO_FLASH_PAIR", None)
    else:
        os.environ["FRIDO_FLASH_PAIR"] = old


@pytest.mark.parametrize("case", CASES)
def test_flash_matches_fp64(dev, case, pair_mode):
    B, N, Cd, split, layout = case
    g = torch.Generator().manual_seed(1000 + N + Cd)
    q = torch.randn(B, N, Cd, generator=g).to(dev)
    k = torch.randn(B, N, Cd, generator=g).to(dev)
    v = torch.randn(B, N, Cd, generator=g).to(dev)
    bias = torch.randn(Cd, generator=g).to(dev)
    res = torch.randn(B, N, Cd, generator=g).to(dev)
    scale = 2.0 * Cd ** -0.5   # logits with std 2: a peaked but not one-hot softmax
    old = os.environ.get("FRIDO_FLASH_SPLIT")
    if split:
        os.environ["FRIDO_FLASH_SPLIT"] = str(split)
    try:
        out = _flash(q, k, v, scale, bias, res, layout)
    finally:
        if old is None:
            os.environ.pop("FRIDO_FLASH_SPLIT", None)
        else:
            os.environ["FRIDO_FLASH_SPLIT"] = old
    ref = _ref(q, k, v, scale, bias, res)
    err = (out.double().cpu() - ref).abs().max().item()
    assert err <= 2e-4 * ref.abs().max().item(), (case, err)


def test_flash_rescale_path(dev, pair_mode):
    """Scores that keep growing along the key axis: every key block raises the row maximum by far more than 2^8, so the
    accumulator rescale runs at every block; plus a no-bias / no-residual launch."""
    B, N, Cd = 1, 512, 64
    g = torch.Generator().manual_seed(7)
    q = torch.ones(B, N, Cd) + 0.1 * torch.randn(B, N, Cd, generator=g)
    ramp = torch.linspace(0.0, 1.5, N).view(1, N, 1)   # +24 (natural log units) per 128-key block
    k = ramp * torch.ones(B, N, Cd) + 0.1 * torch.randn(B, N, Cd, generator=g)
    v = torch.randn(B, N, Cd, generator=g)
    q, k, v = q.to(dev), k.to(dev), v.to(dev)
    scale = 1.0
    out = _flash(q, k, v, scale, None, None, "cbn")
    ref = _ref(q, k, v, scale, None, None)
    err = (out.double().cpu() - ref).abs().max().item()
    assert err <= 5e-4 * ref.abs().max().item(), err   # logits up to ~100: their 2^-17 relative error shows in the weights
    # and the opposite: the maximum sits in the first block, later blocks underflow to zero
    out = _flash(q, k.flip(1).contiguous(), v, scale, None, None, "cbn")
    ref = _ref(q, k.flip(1), v, scale, None, None)
    err = (out.double().cpu() - ref).abs().max().item()
    assert err <= 5e-4 * ref.abs().max().item(), err


def test_flash_rejects_bad_shapes(dev):
    from frido_b200 import _lib as L
    assert not L.lib().frido_attn_flash_eligible(1, 100, 64)
    assert not L.lib().frido_attn_flash_eligible(1, 128, 48)
    assert L.lib().frido_attn_flash_eligible(16, 1024, 384)
    p = L.FlashParams()
    with pytest.raises(L.FridoError):
        L.check(L.lib().frido_attn_flash(C.byref(p), None), "attn_flash")
