"""Host runtime: flat op programs over static device buffers.

A `Program` is built once per (network, stage, batch, geometry), owns every
intermediate buffer it touches, and is executed by the native executor
`frido_run_program` (one C call) — normally inside a captured CUDA graph, so a
sampler step is a single graph launch.  Nothing here computes on the host.
"""
import ctypes as C
import os

import torch

from . import _lib as L


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def default_engine():
    """'bf16x3' = tcgen05 error-compensated BF16x3 (hi*hi + lo*hi + hi*lo, fp32 accumulate; products ~2^-16) for every
               eligible conv/linear with shared weights, 3xTF32 for the attention matmuls (default);
    'tc3'    = tcgen05 error-compensated 3xTF32 everywhere eligible (products ~2^-21, 2x slower than bf16x3);
    'tc'     = tcgen05 single-pass TF32 (operands rounded to 10 mantissa bits, ~1e-3 relative);
    'simt'   = everything on the fp32 SIMT engine (numerical yardstick / debugging).
    Shapes the tensor-core engine does not take always run on the SIMT engine."""
    e = os.environ.get("FRIDO_ENGINE", "bf16x3")
    if e not in ("bf16x3", "tc3", "tc", "simt"):
        raise ValueError(f"FRIDO_ENGINE={e!r}: expected bf16x3, tc3, tc or simt")
    return e


_SK_WS = {}


def sk_workspace(device):
    """One zero-initialised stream-K workspace per device (FridoConvParams.sk_ws): launches on a stream execute in order and
    every launch leaves its arrival counters at zero, so all programs of a device share it.
    CONTRACT (single stream per device): programs of one device must not run CONCURRENTLY on different streams - the arrival
    counters and partial-accumulator slots are shared, and the library's launch-chaining flag is per process.  The host
    runtime only ever enqueues on torch's current stream (and the graphs captured from it); a host that wants two concurrent
    streams on one device passes each its own `sk_ws` (frido_workspace_bytes) in FridoConvParams."""
    device = torch.device(device)
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    ws = _SK_WS.get(key)
    if ws is None:
        ws = torch.zeros(L.SK_WS_BYTES, dtype=torch.uint8, device=device)
        _SK_WS[key] = ws
    return ws


def round_tf32_(t):
    """In-place round-to-nearest (ties away) to TF32 of a device tensor: MMA operands are rounded once
    by their producer so the tensor core's mantissa truncation never bites."""
    if t.is_cuda:
        s = torch.cuda.current_stream(t.device).cuda_stream
        L.check(L.lib().frido_round_tf32(C.c_void_p(t.data_ptr()), C.c_void_p(t.data_ptr()), t.numel(), C.c_void_p(s)), "round_tf32")
    return t


class Src:
    """A strided view used as a conv A-source: element strides (image,row,col,channel)."""

    __slots__ = ("t", "C", "sb", "sy", "sx", "sc", "off")

    def __init__(self, t, C_, sb, sy, sx, sc, off=0):
        self.t, self.C, self.sb, self.sy, self.sx, self.sc, self.off = t, C_, sb, sy, sx, sc, off

    @staticmethod
    def nhwc(t, H, W, C_=None, sub=1, c_off=0, c_total=None):
        """t is a contiguous [B,H,W,Ct] tensor; take channels [c_off, c_off+C_); `sub`
        subsamples rows/cols (nearest down-resize: picks index i*sub)."""
        Ct = c_total if c_total is not None else t.shape[-1]
        C_ = C_ if C_ is not None else Ct
        return Src(t, C_, H * W * Ct, W * Ct * sub, Ct * sub, 1, c_off)

    @staticmethod
    def nchw(t, H, W, c0, c1):
        Ct = t.shape[1]
        return Src(t, c1 - c0, Ct * H * W, W, 1, H * W, c0 * H * W)

    @property
    def ptr(self):
        return self.t.data_ptr() + 4 * self.off


class Program:
    def __init__(self, device, name="", engine=None):
        self.device = torch.device(device)
        self.name = name
        self.ops = []
        self.tags = []
        self.keep = []  # every tensor referenced by an op (keeps storage alive)
        self._pool = {}
        self._arr = None
        self.graph = None
        self.flops = 0  # algorithmic multiply-add*2 of conv/linear/matmul ops
        self.tc_flops = 0
        self.mma_weights = []  # tensors consumed as the W operand of a tcgen05 conv
        self.engine = default_engine() if engine is None else engine
        self.R = 1 if self.engine == "tc" else 0  # single-pass TF32: producers of MMA operands round to TF32
        self.tc_code = {"tc": 1, "tc3": 2, "bf16x3": 3}.get(self.engine, 0)
        self.split_list = []   # (fp32 packed weight, bf16 hi, bf16 lo) of every BF16x3 conv
        self._split_ids = {}

    # ---- buffers -------------------------------------------------------
    def buf(self, *shape, dtype=torch.float32, zero=False):
        n = 1
        for s in shape:
            n *= int(s)
        key = (n, dtype)
        lst = self._pool.get(key)
        if lst and not zero:
            t = lst.pop().view(*shape)
        else:
            t = (torch.zeros if zero else torch.empty)(*shape, dtype=dtype, device=self.device)
        self.keep.append(t)
        return t

    def release(self, t):
        """Return a buffer to the pool: later ops may overwrite it (stream order makes that safe)."""
        self._pool.setdefault((t.numel(), t.dtype), []).append(t.view(-1))

    def hold(self, *ts):
        for t in ts:
            if t is not None:
                self.keep.append(t)

    # ---- op emitters ---------------------------------------------------
    def _add(self, kind, params, tag):
        self.ops.append(L.make_op(kind, params, len(self.tags)))
        self.tags.append(tag)
        self._arr = None

    def conv(self, a0, w, out, *, B, Hin, Win, Hout, Wout, Cout, ksize=1, stride=1, pad=0, ups=1, a1=None,
             bias=None, rowvec=None, rowvec_sb=0, res=None, alpha=1.0, act=L.ACT_NONE, o_sb=None, o_sp=None, o_sn=1,
             w_sb=0, w_ld=0, w_off=0, out_off=0, round_tf32=0, engine=None, csum=None, out_pair=None, w_pair=None,
             alg_flops=None, side=None, out_u8=None, u8_mode=0, nrm=None, presplit=False, tag="conv"):
        """Returns True when `csum` (per-channel GroupNorm sums of the output, [B,Cout,2] fp64) was attached to the op:
        only the tcgen05 engines accumulate it, for dense NHWC outputs with >= 32 pixels per image."""
        if engine is None:
            engine = self.tc_code if (self.tc_code and self._tc_ok(a0, a1, B, Hin, Win, Hout, Wout, Cout, ksize, stride, pad, ups, w,
                                                               w_sb, w_ld, w_off, act, o_sb, o_sp, o_sn, out, out_off, res)) else 0
        if side is not None and engine == 0:
            raise L.FridoError("conv: the fused 1x1 side input needs the tcgen05 engine (check Program.tc_eligible first)")
        if engine != 0 and act == L.ACT_GEGLU and os.environ.get("FRIDO_EXACT_ERF", "0") != "1":
            act = L.ACT_GEGLU_FAST  # tensor-core epilogue: erf by A&S 7.1.26 (|err| < 5e-7), see csrc/common.cuh
        if engine == 0 and act == L.ACT_GEGLU_FAST:
            act = L.ACT_GEGLU
        if engine == 3 and w_sb and w_pair is None:
            engine = 2  # per-image "weights" are activations (attention) without a bf16 pair copy -> 3xTF32
        p = L.ConvParams()
        p.a0, p.c0 = a0.ptr, a0.C
        p.a0_sb, p.a0_sy, p.a0_sx, p.a0_sc = a0.sb, a0.sy, a0.sx, a0.sc
        if a1 is not None:
            p.a1, p.c1 = a1.ptr, a1.C
            p.a1_sb, p.a1_sy, p.a1_sx, p.a1_sc = a1.sb, a1.sy, a1.sx, a1.sc
        c_side = 0
        if side is not None:  # (x0, x1 or None): 1x1 side input accumulated into the same output (ResBlock skip_connection)
            x0, x1 = side
            p.x0, p.cx0, p.x0_sb, p.x0_sy, p.x0_sx = x0.ptr, x0.C, x0.sb, x0.sy, x0.sx
            c_side = x0.C
            self.hold(x0.t)
            if x1 is not None:
                p.x1, p.cx1, p.x1_sb, p.x1_sy, p.x1_sx = x1.ptr, x1.C, x1.sb, x1.sy, x1.sx
                c_side += x1.C
                self.hold(x1.t)
        p.B, p.Hin, p.Win, p.ups = B, Hin, Win, ups
        p.ksize, p.stride, p.pad, p.Hout, p.Wout = ksize, stride, pad, Hout, Wout
        p.w, p.w_sb, p.w_ld, p.Cout = w.data_ptr() + 4 * w_off, w_sb, w_ld, Cout
        if engine == 3 and w_pair is not None:  # operand already stored as a bf16 hi/lo pair by its producer's epilogue
            p.w, p.w_lo = w_pair[0].data_ptr() + 2 * w_off, w_pair[1].data_ptr() + 2 * w_off
            self.hold(w_pair[0], w_pair[1])
        elif engine == 3:
            ent = self._split_ids.get(id(w))
            if ent is None:
                ent = (w, torch.empty(w.shape, dtype=torch.bfloat16, device=w.device),
                       torch.empty(w.shape, dtype=torch.bfloat16, device=w.device))
                self._split_ids[id(w)] = ent
                self.split_list.append(ent)
            p.w, p.w_lo = ent[1].data_ptr() + 2 * w_off, ent[2].data_ptr() + 2 * w_off
            self.hold(ent[1], ent[2])
        p.bias, p.rowvec, p.rowvec_sb, p.res = _ptr(bias), _ptr(rowvec), rowvec_sb, _ptr(res)
        p.alpha, p.act = alpha, act
        n_out = Cout // 2 if act in (L.ACT_GEGLU, L.ACT_GEGLU_FAST) else Cout
        p.out = out.data_ptr() + 4 * out_off
        p.o_sp = n_out if o_sp is None else o_sp
        p.o_sb = Hout * Wout * p.o_sp if o_sb is None else o_sb
        p.o_sn = o_sn
        p.round_tf32, p.engine = round_tf32, engine
        if engine in (1, 2, 3) and out.is_cuda:
            ws = sk_workspace(self.device)
            p.sk_ws, p.sk_ws_bytes = ws.data_ptr(), ws.numel()
            self.hold(ws)
        if out_pair is not None:
            p.out_hi, p.out_lo = out_pair[0].data_ptr() + 2 * out_off, out_pair[1].data_ptr() + 2 * out_off
            self.hold(out_pair[0], out_pair[1])
        if nrm is not None:  # normalise-on-load: (ab [B,C,2], gb [B,HW,2C] or None, silu) - see nf_eligible
            if engine != 3:
                raise L.FridoError("conv: normalise-on-load needs the BF16x3 tcgen05 engine (check Program.nf_eligible first)")
            ab, gb, silu = nrm
            p.nrm_ab, p.nrm_gb, p.nrm_silu = ab.data_ptr(), _ptr(gb), int(silu)
            self.hold(ab, gb)
        if presplit:  # a0 | a1 were written in the engine's operand form (norm_act / upsample2x with out_split)
            if engine != 3 or nrm is not None:
                raise L.FridoError("conv: a pre-split operand needs the BF16x3 tcgen05 engine (check Program.nf_eligible first)")
            p.a_presplit = 1
        if out_u8 is not None:  # uint8 NHWC copy of the outputs (decoder head: sample_diffusion.py:103-121 fused into conv_out)
            if engine != 0:
                raise L.FridoError("conv: out_u8 is a feature of the small-Cout head kernels (SIMT engine)")
            p.out_u8, p.u8_mode = out_u8.data_ptr(), u8_mode
            self.hold(out_u8)
        tw = min(128, 1 << max(Wout - 1, 0).bit_length())
        th = min(128 // tw, 1 << max(Hout - 1, 0).bit_length())
        csum_ok = csum is not None and engine in (1, 2, 3) and o_sn == 1 and act not in (L.ACT_GEGLU, L.ACT_GEGLU_FAST) and tw * th >= 32
        if csum_ok:
            p.chan_sums = csum.data_ptr()
            self.hold(csum)
        self.hold(a0.t, None if a1 is None else a1.t, w, out, bias, rowvec, res)
        # algorithmic FLOPs: zero-padded rows/columns added only to fit the tensor-core tile are not counted
        fl = alg_flops if alg_flops is not None else 2 * B * Hout * Wout * Cout * (ksize * ksize * (a0.C + (a1.C if a1 is not None else 0)) + c_side)
        self.flops += fl
        if engine in (1, 2, 3):
            self.tc_flops += fl
            self.mma_weights.append(w)
        self._add(L.OP_CONV, p, tag)
        return csum_ok

    def tc_eligible(self, a0, a1, out, *, B, Hin, Win, Hout, Wout, Cout, ksize, stride=1, pad=0, res=None):
        """Would conv() run this shape on the tcgen05 engine?  (weights assumed freshly allocated, i.e. aligned)"""
        return bool(self.tc_code) and self._tc_ok(a0, a1, B, Hin, Win, Hout, Wout, Cout, ksize, stride, pad, 1, out, 0, 0, 0,
                                                  L.ACT_NONE, None, None, 1, out, 0, res)

    def nf_eligible(self, a0, a1, out, *, B, H, W, Cout, ksize):
        """Can conv() apply the GroupNorm (+SPADE) (+SiLU) of its input on load (csrc/conv_nf.cu)?  BF16x3 engine, 3x3 / 1x1
        stride-1 conv, and a 128-pixel tile (at most 16 wide) whose halo fits the shared-memory slot with <= 4 images."""
        if self.tc_code != 3:
            return False
        if not self.tc_eligible(a0, a1, out, B=B, Hin=H, Win=W, Hout=H, Wout=W, Cout=Cout, ksize=ksize, pad=ksize // 2):
            return False
        tw = min(16, 1 << max(W - 1, 0).bit_length())
        th = min(128 // tw, 1 << max(H - 1, 0).bit_length())
        tb = 128 // (tw * th)
        hp = ksize // 2
        return tb <= 4 and (tw + 2 * hp) * (th + 2 * hp) * tb <= 208

    def gn_finalize(self, ab, gamma, beta, *, B, HW, c0, c1=0, eps, sums=None, csum0=None, csum1=None, groups=32, tag="gn_finalize"):
        p = L.GnFinalizeParams()
        p.c0, p.c1, p.B, p.HW, p.groups, p.eps = c0, c1, B, HW, groups, eps
        p.sums, p.csum0, p.csum1, p.gamma, p.beta, p.ab = _ptr(sums), _ptr(csum0), _ptr(csum1), gamma.data_ptr(), beta.data_ptr(), ab.data_ptr()
        self.hold(sums, csum0, csum1, gamma, beta, ab)
        self._add(L.OP_GN_FINALIZE, p, tag)

    @staticmethod
    def _tc_ok(a0, a1, B, Hin, Win, Hout, Wout, Cout, ksize, stride, pad, ups, w, w_sb, w_ld, w_off, act, o_sb, o_sp, o_sn,
               out, out_off, res):
        """Shape contract of conv2d_tc (csrc/conv_tc.cu); everything else runs on the SIMT engine."""
        asym = ksize == 3 and stride == 2 and pad == 0  # encoder Downsample (taming model.py:68-72)
        if ups != 1 or ksize not in (1, 3) or (pad != ksize // 2 and not asym):
            return False
        if not (stride == 1 or (stride == 2 and ksize == 3 and os.environ.get('FRIDO_TC_STRIDE2', '1') == '1')):
            return False
        want = ((Hin - 2) // 2 + 1, (Win - 2) // 2 + 1) if asym else ((Hin + stride - 1) // stride, (Win + stride - 1) // stride)
        if (Hout, Wout) != want:
            return False
        if stride == 2 and (2 * min(128, 1 << (Wout - 1).bit_length()) > 256):
            return False
        if a0.C % 32 or (a1 is not None and a1.C % 32) or Cout % 64:
            return False
        if Hout * Wout * B < 128:  # less than one tile of rows: latency-bound, SIMT is as good
            return False
        srcs = [a0] + ([a1] if a1 is not None else [])
        for s in srcs:
            if s.sc != 1 or s.sx % 4 or s.sy % 4 or s.sb % 4 or (s.ptr & 15):
                return False
        Ktot = ksize * ksize * sum(s.C for s in srcs)
        ld = w_ld or Ktot
        if ld % 4 or w_sb % 4 or ((w.data_ptr() + 4 * w_off) & 15):
            return False
        if w_sb:  # per-image weights: a 128-row tile must not straddle images
            tw = min(128, 1 << (Wout - 1).bit_length())
            th = min(128 // tw, 1 << (Hout - 1).bit_length())
            if tw * th != 128:
                return False
        n_out = Cout // 2 if act in (L.ACT_GEGLU, L.ACT_GEGLU_FAST) else Cout
        o_sp_ = n_out if o_sp is None else o_sp
        o_sb_ = Hout * Wout * o_sp_ if o_sb is None else o_sb
        if act in (L.ACT_GEGLU, L.ACT_GEGLU_FAST) and o_sn != 1:
            return False
        if o_sn == 1 and (o_sp_ % 4 or o_sb_ % 4):
            return False
        if (out.data_ptr() + 4 * out_off) & 15 or (res is not None and res.data_ptr() & 15):
            return False
        return True

    def embed_tokens(self, tokens, tok_emb, pos_emb, out, *, B, Lseq, D, tag="embed"):
        p = L.EmbedParams()
        p.tokens, p.B, p.L, p.D, p.vocab = tokens.data_ptr(), B, Lseq, D, tok_emb.shape[0]
        p.tok_emb, p.pos_emb, p.out = tok_emb.data_ptr(), pos_emb.data_ptr(), out.data_ptr()
        self.hold(tokens, tok_emb, pos_emb, out)
        self._add(L.OP_EMBED, p, tag)

    def mha_small(self, qkv, out, *, B, Lseq, H, Dh, scale, tag="mha"):
        p = L.MhaParams()
        p.qkv, p.B, p.L, p.H, p.Dh, p.scale, p.out = qkv.data_ptr(), B, Lseq, H, Dh, scale, out.data_ptr()
        self.hold(qkv, out)
        self.flops += 4 * B * H * Lseq * Lseq * Dh
        self._add(L.OP_MHA, p, tag)

    def attn_small(self, q, k, v, out, *, B, N, Nk, Cdim, scale, q_off=0, q_sb, q_ld, k_off=0, k_sb, k_ld, v_off=0, v_sb, v_ld,
                   ln=None, ln_eps=1e-5, bias=None, res=None, ln2=None, ln2_eps=1e-5, out2=None, tag="attn_small"):
        """Fused softmax(scale LN(q) k^T) v + bias + res for short key sequences (csrc/attn.cu); offsets/strides in
        floats; ln = (gamma, beta) applies a LayerNorm to the query rows first; res is a dense [B,N,C] tensor."""
        p = L.AttnParams()
        p.q, p.q_sb, p.q_ld = q.data_ptr() + 4 * q_off, q_sb, q_ld
        p.k, p.k_sb, p.k_ld = k.data_ptr() + 4 * k_off, k_sb, k_ld
        p.v, p.v_sb, p.v_ld = v.data_ptr() + 4 * v_off, v_sb, v_ld
        p.B, p.N, p.Nk, p.C, p.scale = B, N, Nk, Cdim, scale
        p.out, p.o_sb, p.o_ld = out.data_ptr(), N * Cdim, Cdim
        if ln is not None:
            p.ln_gamma, p.ln_beta, p.ln_eps = ln[0].data_ptr(), ln[1].data_ptr(), ln_eps
            self.hold(ln[0], ln[1])
        p.bias = _ptr(bias)
        if res is not None:
            p.res, p.r_sb, p.r_ld = res.data_ptr(), N * Cdim, Cdim
        if ln2 is not None:  # second output: LayerNorm of the produced rows (the block's next norm)
            p.ln2_gamma, p.ln2_beta, p.ln2_eps, p.out2 = ln2[0].data_ptr(), ln2[1].data_ptr(), ln2_eps, out2.data_ptr()
            self.hold(ln2[0], ln2[1], out2)
        self.hold(q, k, v, out, bias, res)
        self.flops += 4 * B * N * Nk * Cdim
        self._add(L.OP_ATTN, p, tag)

    @staticmethod
    def flash_eligible(B, N, Cdim):
        """Streaming-softmax tcgen05 attention (csrc/attn_flash.cu): whole 128-token tiles, channels in 32s; FRIDO_FLASH=0 keeps
        the QK^T -> softmax -> PV launches."""
        return os.environ.get("FRIDO_FLASH", "1") == "1" and bool(L.lib().frido_attn_flash_eligible(B, N, Cdim))

    def flash(self, q_pair, k_pair, vt_pair, out, *, B, N, Cdim, scale, vt_sb, vt_ld, bias=None, res=None, q_off=0, q_ld=None,
              q_sb=None, k_off=0, k_ld=None, k_sb=None, tag="attn.flash"):
        """out = res + bias + softmax(scale q k^T) v in one launch; q_pair / k_pair = (hi, lo) bf16 tensors holding the rows
        [B,N,C] at element offset q_off / k_off with row stride q_ld / k_ld (default: dense), vt_pair the values
        channel-major: element (b, c, key) at b * vt_sb + c * vt_ld + key."""
        p = L.FlashParams()
        q_ld, k_ld = q_ld or Cdim, k_ld or Cdim
        p.q_hi, p.q_lo = q_pair[0].data_ptr() + 2 * q_off, q_pair[1].data_ptr() + 2 * q_off
        p.q_sb, p.q_ld = q_sb or N * q_ld, q_ld
        p.k_hi, p.k_lo = k_pair[0].data_ptr() + 2 * k_off, k_pair[1].data_ptr() + 2 * k_off
        p.k_sb, p.k_ld = k_sb or N * k_ld, k_ld
        p.vt_hi, p.vt_lo, p.vt_sb, p.vt_ld = vt_pair[0].data_ptr(), vt_pair[1].data_ptr(), vt_sb, vt_ld
        p.B, p.N, p.C, p.scale = B, N, Cdim, scale
        p.bias = _ptr(bias)
        if res is not None:
            p.res, p.r_sb, p.r_ld = res.data_ptr(), N * Cdim, Cdim
        p.out, p.o_sb, p.o_ld = out.data_ptr(), N * Cdim, Cdim
        self.hold(*q_pair, *k_pair, *vt_pair, out, bias, res)
        self.flops += 4 * B * N * N * Cdim
        self.tc_flops += 4 * B * N * N * Cdim
        self._add(L.OP_FLASH, p, tag)

    def upsample2x(self, x, out, *, B, H, W, Cdim, round_tf32=0, out_split=0, tag="upsample2x"):
        p = L.UpsampleParams()
        p.x, p.B, p.H, p.W, p.C, p.round_tf32, p.out, p.out_split = x.data_ptr(), B, H, W, Cdim, round_tf32, out.data_ptr(), int(out_split)
        self.hold(x, out)
        self._add(L.OP_UPSAMPLE, p, tag)

    def linear(self, a, w, out, *, M, K, N, bias=None, res=None, act=L.ACT_NONE, a_ld=None, a_off=0, out_ld=None,
               rowvec=None, round_tf32=0, engine=None, out_pair=None, tag="linear"):
        """out[M,N] = act(a[M,K] @ w[N,K]^T + bias (+rowvec) (+res))  — rows are 'pixels' of one image."""
        a_ld = K if a_ld is None else a_ld
        src = Src(a, K, 0, 0, a_ld, 1, a_off)
        self.conv(src, w, out, B=1, Hin=1, Win=M, Hout=1, Wout=M, Cout=N, bias=bias, res=res, act=act,
                  rowvec=rowvec, o_sp=out_ld, round_tf32=round_tf32, engine=engine, out_pair=out_pair, tag=tag)

    def zero(self, t, tag="zero"):
        p = L.ZeroParams()
        p.ptr, p.nbytes = t.data_ptr(), t.numel() * t.element_size()
        self.hold(t)
        self._add(L.OP_ZERO, p, tag)

    def chan_stats(self, a0, c0, csum, *, B, HW, tag="chan_stats"):
        """Per-channel (sum, sum of squares) of an NHWC tensor into csum [B,c0,2] fp64 (zero on entry): the statistics a
        tcgen05 conv epilogue would have produced, for tensors that come from the SIMT engine."""
        self.gn_stats(a0, c0, csum, B=B, HW=HW, groups=0, tag=tag)

    def gn_stats(self, a0, c0, sums, *, B, HW, a1=None, c1=0, groups=32, tag="gn_stats"):
        p = L.GnStatsParams()
        p.a0, p.a1, p.c0, p.c1, p.B, p.HW, p.groups, p.sums = _ptr(a0), _ptr(a1), c0, c1, B, HW, groups, sums.data_ptr()
        self.hold(a0, a1, sums)
        self._add(L.OP_GN_STATS, p, tag)

    def norm_act(self, a0, c0, sums, gamma, beta, out, *, B, HW, eps, a1=None, c1=0, gb=None, silu=1, groups=32,
                 round_tf32=0, csum0=None, csum1=None, out_split=0, tag="norm_act"):
        p = L.NormActParams()
        p.a0, p.a1, p.c0, p.c1, p.B, p.HW, p.groups = _ptr(a0), _ptr(a1), c0, c1, B, HW, groups
        p.sums, p.eps, p.gamma, p.beta, p.gb = _ptr(sums), eps, gamma.data_ptr(), beta.data_ptr(), _ptr(gb)
        p.csum0, p.csum1 = _ptr(csum0), _ptr(csum1)
        self.hold(csum0, csum1)
        p.silu, p.round_tf32, p.out, p.out_split = silu, round_tf32, out.data_ptr(), int(out_split)
        self.hold(a0, a1, sums, gamma, beta, gb, out)
        self._add(L.OP_NORM_ACT, p, tag)

    def layernorm(self, x, gamma, beta, out, *, rows, Cdim, eps=1e-5, round_tf32=0, out_pair=None, tag="layernorm"):
        p = L.LayerNormParams()
        p.x, p.rows, p.C, p.eps, p.gamma, p.beta = x.data_ptr(), rows, Cdim, eps, gamma.data_ptr(), beta.data_ptr()
        p.round_tf32, p.out = round_tf32, out.data_ptr()
        if out_pair is not None:
            p.out_hi, p.out_lo = out_pair[0].data_ptr(), out_pair[1].data_ptr()
            self.hold(out_pair[0], out_pair[1])
        self.hold(x, gamma, beta, out)
        self._add(L.OP_LAYERNORM, p, tag)

    def softmax(self, s, *, rows, n, ld, scale, out=None, round_tf32=0, tag="softmax"):
        out = s if out is None else out
        p = L.SoftmaxParams()
        p.s, p.rows, p.n, p.ld, p.scale, p.round_tf32, p.out = s.data_ptr(), rows, n, ld, scale, round_tf32, out.data_ptr()
        self.hold(s, out)
        self._add(L.OP_SOFTMAX, p, tag)

    def time_embed(self, ts, out, *, B, dim, max_period=10000.0, tag="time_embed"):
        import math
        half = dim // 2
        # util.py:160-162, evaluated by the same CPU torch ops as the reference, then shipped to the device
        freqs = torch.exp(-math.log(max_period) * torch.arange(start=0, end=half, dtype=torch.float32) / half).to(self.device)
        p = L.TimeEmbedParams()
        p.t, p.B, p.dim, p.max_period, p.freqs, p.out = ts.data_ptr(), B, dim, max_period, freqs.data_ptr(), out.data_ptr()
        self.hold(ts, out, freqs)
        self._add(L.OP_TIME_EMBED, p, tag)

    def step_begin(self, step, t_table, ts, *, B, T, use_next=0, tag="step_begin"):
        p = L.StepBeginParams()
        p.step, p.t_table, p.use_next, p.T, p.ts, p.B = step.data_ptr(), t_table.data_ptr(), use_next, T, ts.data_ptr(), B
        self.hold(step, t_table, ts)
        self._add(L.OP_STEP_BEGIN, p, tag)

    def update(self, x, eps, coef, step, x_prev, *, B, c_start, c_end, HW, eps_uncond=None, cfg_scale=1.0, advance=1,
               plms_order=0, plms_mode=0, hist=None, eps_save=None, noise=None, seed=0, seed_dev=None, temperature=1.0, x_dup=None,
               pred_x0=None, tag="sampler_update"):
        p = L.UpdateParams()
        p.x, p.eps, p.eps_uncond, p.cfg_scale = x.data_ptr(), eps.data_ptr(), _ptr(eps_uncond), cfg_scale
        p.B, p.c_start, p.c_end, p.HW = B, c_start, c_end, HW
        p.coef, p.step, p.advance = coef.data_ptr(), step.data_ptr(), advance
        p.plms_order, p.plms_mode, p.hist, p.eps_save = plms_order, plms_mode, _ptr(hist), _ptr(eps_save)
        p.noise, p.seed, p.seed_dev, p.temperature = _ptr(noise), seed, _ptr(seed_dev), temperature
        p.x_prev, p.x_dup, p.pred_x0 = x_prev.data_ptr(), _ptr(x_dup), _ptr(pred_x0)
        self.hold(x, eps, eps_uncond, coef, step, hist, eps_save, noise, seed_dev, x_prev, x_dup, pred_x0)
        self._add(L.OP_UPDATE, p, tag)

    def blend(self, x, x0, mask, sqrt_acp, sqrt_1m_acp, step, t_table, *, B, Cdim, HW, T, noise=None, x_dup=None, seed=0,
              seed_dev=None, tag="mask_blend"):
        p = L.BlendParams()
        p.x, p.x_dup, p.x0, p.mask, p.noise = x.data_ptr(), _ptr(x_dup), x0.data_ptr(), mask.data_ptr(), _ptr(noise)
        p.sqrt_acp, p.sqrt_1m_acp, p.step, p.t_table, p.T = sqrt_acp.data_ptr(), sqrt_1m_acp.data_ptr(), step.data_ptr(), t_table.data_ptr(), T
        p.B, p.C, p.HW, p.seed, p.seed_dev = B, Cdim, HW, seed, _ptr(seed_dev)
        self.hold(x, x_dup, x0, mask, noise, sqrt_acp, sqrt_1m_acp, step, t_table, seed_dev)
        self._add(L.OP_BLEND, p, tag)

    def snap(self, x, *, B, Ctot, H, W, c_start, c_end, n, tag="stage_snap"):
        p = L.SnapParams()
        p.x, p.B, p.C, p.H, p.W, p.c_start, p.c_end, p.n = x.data_ptr(), B, Ctot, H, W, c_start, c_end, n
        self.hold(x)
        self._add(L.OP_SNAP, p, tag)

    def conv_transpose2d(self, x, w, bias, out, *, B, H, W, Cin, Cout, x_ld=0, x_off=0, tag="conv_transpose2d"):
        p = L.ConvT2dParams()
        p.x, p.B, p.H, p.W, p.Cin, p.Cout = x.data_ptr() + 4 * x_off, B, H, W, Cin, Cout
        p.w, p.bias, p.out, p.x_ld = w.data_ptr(), _ptr(bias), out.data_ptr(), x_ld
        self.hold(x, w, bias, out)
        self.flops += 2 * B * H * W * 16 * Cin * Cout
        self._add(L.OP_CONVT, p, tag)

    def assemble_latent(self, h, out, *, B, H, W, e, sh, scale, C_total, c_off, tag="assemble_latent"):
        p = L.AssembleParams()
        p.h, p.B, p.H, p.W, p.e, p.sh, p.scale, p.out, p.C_total, p.c_off = h.data_ptr(), B, H, W, e, sh, scale, out.data_ptr(), C_total, c_off
        self.hold(h, out)
        self._add(L.OP_ASSEMBLE, p, tag)

    def vq(self, z, codebook, out, indices, *, B, C_total, HW, c_start, e_dim, scale_factor, out_C, out_coff, z_nhwc=0, tag="vq"):
        p = L.VqParams()
        p.z, p.B, p.C_total, p.HW, p.c_start, p.e_dim = z.data_ptr(), B, C_total, HW, c_start, e_dim
        p.scale_factor, p.codebook, p.n_e = scale_factor, codebook.data_ptr(), codebook.shape[0]
        p.out, p.out_C, p.out_coff, p.indices, p.z_nhwc = out.data_ptr(), out_C, out_coff, indices.data_ptr(), z_nhwc
        self.hold(z, codebook, out, indices)
        self._add(L.OP_VQ, p, tag)

    def prepare_weights(self):
        """(Re)compute the bf16 hi/lo copies of every weight a BF16x3 conv reads (after packing / re-packing)."""
        for w, hi, lo in self.split_list:
            if w.is_cuda:
                s = torch.cuda.current_stream(w.device).cuda_stream
                L.check(L.lib().frido_split_bf16(C.c_void_p(w.data_ptr()), C.c_void_p(hi.data_ptr()), C.c_void_p(lo.data_ptr()),
                                                 w.numel(), C.c_void_p(s)), "split_bf16")

    # ---- execution -----------------------------------------------------
    def _array(self):
        if self._arr is None:
            self._arr = (L.Op * len(self.ops))(*self.ops)
        return self._arr

    def run(self, stream=None):
        """Enqueue all ops on `stream` (default: torch's current stream)."""
        if not self.ops:
            return
        s = torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        rc = L.lib().frido_run_program(C.cast(self._array(), C.c_void_p), len(self.ops), C.c_void_p(s))
        L.check(rc, f"program {self.name}")

    def capture(self):
        """Capture the program into a CUDA graph (after one eager warm-up run)."""
        self.run()
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.run()
        self.graph = g
        return g

    def replay(self):
        if self.graph is None:
            self.run()
        else:
            self.graph.replay()

    def __len__(self):
        return len(self.ops)
