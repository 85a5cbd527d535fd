"""PyUNetModel — drop-in mirror of frido/modules/diffusionmodules/pyunet.py:447-950.

Same constructor keywords, attributes and state-dict keys as the reference;
`forward(x, timesteps, context, y, stage)` returns the same tensor.  The body
of forward is a *program* of libfrido_b200 launches (see UNetPlan):

  * activations NHWC fp32; tokens of a SpatialTransformer are the same memory;
  * skip concat, nearest x2 upsampling, SPADE's nearest down-resize are folded
    into conv addressing (no copies);
  * step-invariant work is hoisted into a per-stage PROLOGUE program: the
    SPADE gamma/beta maps of all norm sites (they depend only on the frozen
    coarse channels, ddim.py:246,266 + pyunet.py:906-911) and the
    cross-attention K/V of the context (attention.py:175-176);
  * the 22 ResBlock timestep projections run as one GEMM.
"""
import math
import os

import torch
from torch import nn

from . import _lib as L
from . import modules as M
from . import packing as PK
from .program import Program, Src, round_tf32_


class PyUNetModel(nn.Module):
    def __init__(self, image_size, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions,
                 dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2, num_classes=None,
                 use_checkpoint=False, use_fp16=False, num_heads=-1, num_head_channels=-1, num_heads_upsample=-1,
                 use_scale_shift_norm=False, use_embed=False, num_stage=1, resblock_updown=False,
                 use_new_attention_order=False, use_spatial_transformer=False, transformer_depth=1, context_dim=None,
                 n_embed=None, legacy=True, use_split_head=False, split_embed_dim_list=[], use_SPADE_norm=False,
                 use_pos_embed=False, use_mscond=False, use_stage_expert=False):
        super().__init__()
        # the shipped configs (configs/frido/**) use exactly this subset; anything else is refused loudly
        unsupported = dict(dims=(dims, 2), num_classes=(num_classes, None), use_scale_shift_norm=(use_scale_shift_norm, False),
                           resblock_updown=(resblock_updown, False), n_embed=(n_embed, None), legacy=(legacy, True),
                           use_pos_embed=(use_pos_embed, False), use_mscond=(use_mscond, False),
                           use_stage_expert=(use_stage_expert, False), use_fp16=(use_fp16, False),
                           conv_resample=(conv_resample, True))
        for k, (v, want) in unsupported.items():
            if v != want:
                raise NotImplementedError(f"PyUNetModel({k}={v!r}) is outside the B200 hot path (supported: {want!r})")
        if not use_spatial_transformer or context_dim is None:
            raise NotImplementedError("only use_spatial_transformer=True with a context_dim is supported")
        if not use_split_head:
            raise NotImplementedError("only use_split_head=True (Frido split heads) is supported")
        if isinstance(context_dim, (list, tuple)):
            context_dim = list(context_dim)[0]
        split = [int(v) for v in split_embed_dim_list]
        assert len(split) != 0, "specify split head embed dim."
        assert sum(split) == in_channels
        self.image_size, self.in_channels, self.model_channels = image_size, in_channels, model_channels
        self.out_channels, self.num_res_blocks = out_channels, num_res_blocks
        self.attention_resolutions = [int(a) for a in attention_resolutions]
        self.dropout, self.channel_mult = dropout, [int(c) for c in channel_mult]
        self.num_classes, self.num_stage = num_classes, num_stage
        self.use_split_head, self.split_embed_dim_list = use_split_head, split
        self.use_SPADE_norm, self.context_dim = use_SPADE_norm, context_dim
        self.transformer_depth = transformer_depth
        self.dtype = torch.float32
        spade = use_SPADE_norm
        mc = model_channels
        ted = mc * 4
        self.time_embed = M.Seq(nn.Linear(mc, ted), M.Marker(), nn.Linear(ted, ted))
        if num_stage > 1:
            self.stage_emb = nn.Embedding(num_stage, ted)
        if spade:
            self.pre_input_cond_blocks = nn.ModuleList(
                [M.Seq(nn.Conv2d(sum(split[: i + 1]), mc, 3, padding=1)) for i in range(len(split) - 1)])
            self.pre_input_blocks = nn.ModuleList([M.Seq(nn.Conv2d(split[i], mc, 3, padding=1)) for i in range(len(split))])
        else:
            self.pre_input_blocks = nn.ModuleList(
                [M.Seq(nn.Conv2d(sum(split[: i + 1]), mc, 3, padding=1)) for i in range(len(split))])

        def st(ch):
            return M.SpatialTransformer(ch, mc, transformer_depth, context_dim, spade)

        self.input_blocks = nn.ModuleList([])
        chans = [mc]
        ch, ds = mc, 1
        for level, mult in enumerate(self.channel_mult):
            for _ in range(num_res_blocks):
                layers = [M.ResBlock(ch, mc, ted, mult * mc, spade)]
                ch = mult * mc
                if ds in self.attention_resolutions:
                    layers.append(st(ch))
                self.input_blocks.append(M.Seq(*layers))
                chans.append(ch)
            if level != len(self.channel_mult) - 1:
                self.input_blocks.append(M.Seq(M.Downsample(ch, ch)))
                chans.append(ch)
                ds *= 2
        self.middle_block = M.Seq(M.ResBlock(ch, mc, ted, ch, spade), st(ch), M.ResBlock(ch, mc, ted, ch, spade))
        self.output_blocks = nn.ModuleList([])
        for level, mult in list(enumerate(self.channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                ich = chans.pop()
                layers = [M.ResBlock(ch + ich, mc, ted, mc * mult, spade)]
                ch = mc * mult
                if ds in self.attention_resolutions:
                    layers.append(st(ch))
                if level and i == num_res_blocks:
                    layers.append(M.Upsample(ch, ch))
                    ds //= 2
                self.output_blocks.append(M.Seq(*layers))
        self.out = nn.ModuleList([M.Seq(M.gn(ch, 1e-5), M.Marker(), M._zero(nn.Conv2d(mc, split[i], 3, padding=1)))
                                  for i in range(len(split))])
        self._plans = {}
        self._pack_version = 0
        self._fingerprint = None

    # ------------------------------------------------------------------
    def invalidate(self):
        """Weights changed under us (EMA swap writes through .data.copy_, ema.py:51,76 —
        invisible to pointer/version checks): every plan re-packs at its next use (`repack_if_stale`)."""
        self._pack_version += 1
        self._fingerprint = None

    def _weights_fingerprint(self):
        """L1 and L2 norm of every parameter tensor (two multi-tensor passes over the weights, ~1 ms for 2 GB, one small
        device->host copy): what `invalidate_if_changed` compares to notice in-place rewrites nobody announced."""
        ps = [p.detach() for p in self.parameters()]
        return torch.stack(torch._foreach_norm(ps, 2) + torch._foreach_norm(ps, 1)).double().cpu()

    def invalidate_if_changed(self):
        """Called at the start of every `sampler.sample()`: the reference lets anyone rewrite the weights in place between
        calls (`LitEma.copy_to/restore`, frido.py:182-194).  `ema_scope` and `load_state_dict` invalidate explicitly; this
        catches the rest without paying a full re-pack (2 GB of packing + bf16 splitting per plan) on every call.
        FRIDO_ALWAYS_REPACK=1 restores the unconditional re-pack."""
        if os.environ.get("FRIDO_ALWAYS_REPACK", "0") == "1":
            self.invalidate()
            return
        fp = self._weights_fingerprint()
        old = getattr(self, "_fingerprint", None)
        if old is None or old.shape != fp.shape or not torch.equal(old, fp):
            self._pack_version += 1
        self._fingerprint = fp

    def plan(self, stage, B, H, W, L_ctx):
        key = (stage, B, H, W, L_ctx)
        p = self._plans.get(key)
        if p is None:
            p = UNetPlan(self, stage, B, H, W, L_ctx)
            self._plans[key] = p
        return p

    @torch.no_grad()
    def forward(self, x, timesteps=None, context=None, y=None, stage=None, **kwargs):
        assert y is None, "must specify y if and only if the model is class-conditional"
        if not x.is_cuda:
            raise L.FridoError("PyUNetModel runs on a CUDA device only (no CPU path)")
        stage = 0 if stage is None else int(stage)
        B, C, H, W = x.shape
        assert C >= sum(self.split_embed_dim_list[: stage + 1])
        plan = self.plan(stage, B, H, W, context.shape[1])
        plan.repack_if_stale()
        plan.x_in[:, : plan.c_end].copy_(x[:, : plan.c_end])
        plan.ts.copy_(timesteps.to(torch.int64))
        plan.ctx.copy_(context)
        plan.prologue.run()
        plan.step.run()
        return plan.eps.clone()


def fold_self_attention(ca):
    """Weight products of a self-attention CrossAttention module (attention.py:172-191, re-associated; x = LN(h)):
         sim   = (x Wq^T)(x Wk^T)^T       = (x A^T) x^T        with A   = Wk^T Wq   (keys are x itself)
         to_out(P (x Wv^T)) - b_o         = P (x Wv'^T)        with Wv' = Wo Wv     (values carry to_out)
    Products are taken in fp64 and rounded once; returns (A, Wv') as [C,C] fp32 linear weights ([out,in])."""
    return PK.fold_self_attention(ca.to_q.weight, ca.to_k.weight, ca.to_v.weight, ca.to_out[0].weight)


def fold_cross_attention_weights(ca):
    """Linear weights that turn K = ctx Wk^T and V = ctx Wv^T into the operands of the fused cross-attention kernel:
         K' = K Wq   (sim = (x Wq^T) K^T = x K'^T)        -> weight Wq^T   ([out,in] = [C,C])
         V' = V Wo^T (to_out(P V) - b_o = P V')           -> weight Wo"""
    return PK.transpose(ca.to_q.weight), PK.copy(ca.to_out[0].weight)


def _pack_conv(w):
    """OIHW -> [O][kh*kw][I] (K-major rows for the implicit GEMM); a native launch (csrc/pack.cu)."""
    return PK.conv_weight(w)


class UNetPlan:
    """Programs + static buffers for one (stage, batch, geometry)."""

    def __init__(self, net: PyUNetModel, stage, B, H, W, Lc):
        self.net, self.stage, self.B, self.H, self.W, self.Lc = net, stage, B, H, W, Lc
        dev = next(net.parameters()).device
        self.dev = dev
        split = net.split_embed_dim_list
        self.spade = net.use_SPADE_norm
        self.c_cond = sum(split[:stage]) if self.spade else 0
        self.c_end = sum(split[: stage + 1])
        self.e_s = split[stage]
        self.packers = []
        self.version = net._pack_version
        f32 = dict(dtype=torch.float32, device=dev)
        self.x_in = torch.zeros(B, self.c_end, H, W, **f32)
        self.ts = torch.zeros(B, dtype=torch.int64, device=dev)
        self.ctx = torch.zeros(B, Lc, net.context_dim, **f32)
        self.eps = torch.zeros(B, self.e_s, H, W, **f32)
        self.Lp = (Lc + 63) // 64 * 64  # padded key count: zero keys, so cross-attention can run on the tensor-core engine
        self.prologue = Program(dev, f"unet.s{stage}.prologue")
        self.step = Program(dev, f"unet.s{stage}.step")
        self._gn_slots = []
        self._mma_ids = set()
        self._build()
        self._finalize_weights()

    # ---- weight packing (re-runnable: same destination pointers) -------
    def _packed(self, fn):
        dst = fn().to(self.dev).contiguous()
        self.packers.append((dst, fn))
        return dst

    def repack(self):
        for dst, fn in self.packers:
            PK.place(fn().view(1, -1), dst.view(1, -1))
            if id(dst) in self._mma_ids:
                round_tf32_(dst)
        self.prologue.prepare_weights()
        self.step.prepare_weights()
        self.version = self.net._pack_version

    def _finalize_weights(self):
        """Weights that feed the tcgen05 engine are rounded to TF32 once, here (and on every repack)."""
        self._mma_ids = {id(t) for prog in (self.prologue, self.step) if prog.R for t in prog.mma_weights}
        for dst, _ in self.packers:
            if id(dst) in self._mma_ids:
                round_tf32_(dst)
        self.prologue.prepare_weights()
        self.step.prepare_weights()

    def repack_if_stale(self):
        if self.version != self.net._pack_version:
            self.repack()

    def _conv_w(self, conv):
        return self._packed(lambda: _pack_conv(conv.weight))

    def _vec(self, p):
        return self._packed(lambda: PK.copy(p))

    # ---- helpers ---------------------------------------------------------
    def _csum_new(self, C):
        """Carve a [B,C,2] fp64 block for producer-side GroupNorm sums out of the per-step zeroed pool."""
        n = self.B * C * 2
        if self._csum_used + n > self._csum_pool.numel():
            return None
        t = self._csum_pool[self._csum_used:self._csum_used + n].view(self.B, C, 2)
        self._csum_used += n
        return t

    def _conv_stats(self, prog, a0, w, out, Cout, **kw):
        """conv whose output feeds a GroupNorm: ask the epilogue for the per-channel sums; remember them by tensor."""
        cs = self._csum_new(Cout) if prog is self.step else None
        if prog.conv(a0, w, out, Cout=Cout, csum=cs, **kw):
            self._csum[id(out)] = cs

    def _gn_slot(self, prog):
        t = self._sums_all[len(self._gn_slots)]
        self._gn_slots.append(t)
        return t

    def _norm(self, prog, xs, cs, h, w, norm, eps, silu, site, out_split=0):
        """GroupNorm(+SPADE)(+SiLU) over the (possibly concatenated) NHWC sources xs -> new buffer (out_split: written in the
        BF16x3 engine's operand form for a conv that takes `presplit=True`)."""
        B = self.B
        hw = h * w
        C = sum(cs)
        a1 = xs[1] if len(xs) > 1 else None
        c1 = cs[1] if len(xs) > 1 else 0
        cs0 = self._csum.get(id(xs[0]))
        cs1 = self._csum.get(id(a1)) if a1 is not None else None
        fused = cs0 is not None and (a1 is None or cs1 is not None)
        sums = None
        if not fused:  # producer could not provide the statistics (SIMT-engine producer): separate reduction pass
            cs0 = cs1 = None
            sums = self._gn_slot(prog)
            prog.gn_stats(xs[0], cs[0], sums, B=B, HW=hw, a1=a1, c1=c1)
        is_spade = isinstance(norm, M.SPADE)
        g = norm.param_free_norm if is_spade else norm
        gb = self._spade_site(norm, C, h, w) if (is_spade and self.c_cond) else None
        out = prog.buf(B, hw, C)
        prog.norm_act(xs[0], cs[0], sums, self._vec(g.weight), self._vec(g.bias), out, B=B, HW=hw, eps=eps, a1=a1,
                      c1=c1, gb=gb, silu=silu, round_tf32=prog.R, csum0=cs0, csum1=cs1, out_split=out_split, tag=site)
        return out

    def _norm_mode(self, prog, xs, cs, h, w, norm, out, Cout, ksize, side=False):
        """How the GroupNorm (+SPADE) (+SiLU) in front of a conv reaches the tensor core (FRIDO_FUSE_NORM):
          'fused' (1)  the conv normalises on load (csrc/conv_nf.cu): no norm_act pass, no normalised tensor in memory;
          'split' (2)  norm_act writes the engine's bf16 hi | lo operand form and the conv's halo-resident path just feeds it;
          'plain' (0)  norm_act writes fp32, the conv re-fetches and splits the tile per filter tap (csrc/conv_tc.cu).
        Default 'auto': fused where there are no SPADE maps to stream (stage 0, and models without SPADE: measured 12.56 ms
        per step against 12.74 with only the single-N-tile sites fused and 12.85 plain), plain where there are (a fused conv
        re-reads the maps per N tile and per halo: 14.4 against 13.1 ms)."""
        mode = os.environ.get("FRIDO_FUSE_NORM", "auto")
        a1 = Src.nhwc(xs[1], h, w) if len(xs) > 1 else None
        if mode == "0" or any(c % 32 for c in cs) or not prog.nf_eligible(Src.nhwc(xs[0], h, w), a1, out, B=self.B, H=h, W=w, Cout=Cout, ksize=ksize):
            return "plain"
        if ksize == 1:
            return "fused" if os.environ.get("FRIDO_FUSE_NORM_1X1", "0") == "1" else "plain"
        if side and mode == "auto" and os.environ.get("FRIDO_FUSE_NORM_SIDE", "1") == "0":
            return "plain"   # A/B aid: convs that carry the fused 1x1 skip input stay on conv_tc.cu
        if mode == "1":
            return "fused"
        if mode == "2":
            return "split"
        spade = isinstance(norm, M.SPADE) and bool(self.c_cond)
        wide = os.environ.get("FRIDO_FUSE_NORM_WIDE", "1") == "1"   # A/B aid: 0 = fuse only single-N-tile convs (C_out <= 192)
        if not spade and (wide or Cout <= 192):
            return "fused"
        # SPADE sites: a separate norm_act pass streams the maps once; its conv then runs on the CTA-pair kernel of conv_tc.cu
        # (measured 13.27 ms per stage-1 step) or, with FRIDO_SPADE_SPLIT=1, reads split operands through conv_nf.cu (13.43)
        return "split" if os.environ.get("FRIDO_SPADE_SPLIT", "0") == "1" else "plain"

    def _norm_on_load(self, prog, xs, cs, h, w, norm, eps, silu, out, Cout, ksize, side=False):
        """GroupNorm(+SPADE)(+SiLU) handed to the consuming conv instead of a norm_act pass (csrc/conv_nf.cu): emits the
        tiny statistics -> (scale, shift) kernel and returns (`nrm` argument of Program.conv, buffer to release after the
        conv), or None when the conv cannot normalise on load (then `_norm` materialises the tensor as before)."""
        B = self.B
        if self._norm_mode(prog, xs, cs, h, w, norm, out, Cout, ksize, side) != "fused":
            return None
        hw = h * w
        C = sum(cs)
        x1 = xs[1] if len(xs) > 1 else None
        c1 = cs[1] if len(xs) > 1 else 0
        cs0 = self._csum.get(id(xs[0]))
        cs1 = self._csum.get(id(x1)) if x1 is not None else None
        sums = None
        if cs0 is None or (x1 is not None and cs1 is None):  # no producer statistics: separate reduction pass
            cs0 = cs1 = None
            sums = self._gn_slot(prog)
            prog.gn_stats(xs[0], cs[0], sums, B=B, HW=hw, a1=x1, c1=c1)
        is_spade = isinstance(norm, M.SPADE)
        g = norm.param_free_norm if is_spade else norm
        gb = self._spade_site(norm, C, h, w) if (is_spade and self.c_cond) else None
        ab = prog.buf(B, C, 2)
        prog.gn_finalize(ab, self._vec(g.weight), self._vec(g.bias), B=B, HW=hw, c0=cs[0], c1=c1, eps=eps, sums=sums, csum0=cs0,
                         csum1=cs1, tag="gn_finalize")
        return (ab, gb, silu), ab

    def _spade_site(self, sp, C, h, w):
        """Prologue: gamma|beta = conv3x3(ReLU(conv3x3(nearest_resize(h_cond)))) (spade_norm.py:52-55)."""
        P, B = self.prologue, self.B
        hw = h * w
        sub = self.H // h
        assert sub * h == self.H and sub * w == self.W
        mc = self.net.model_channels
        nh = sp.mlp_shared[0].weight.shape[0]
        actv = P.buf(B, hw, nh)
        P.conv(Src.nhwc(self.h_cond, self.H, self.W, mc, sub=sub), self._conv_w(sp.mlp_shared[0]), actv, B=B, Hin=h,
               Win=w, Hout=h, Wout=w, Cout=nh, ksize=3, pad=1, bias=self._vec(sp.mlp_shared[0].bias), act=L.ACT_RELU,
               round_tf32=P.R, tag="spade.shared")
        wgb = self._packed(lambda: PK.conv_rows([sp.mlp_gamma.weight, sp.mlp_beta.weight]))
        bgb = self._packed(lambda: PK.cat_rows([sp.mlp_gamma.bias, sp.mlp_beta.bias]))
        gb = torch.empty(B, hw, 2 * C, dtype=torch.float32, device=self.dev)  # lives across steps: not pooled
        P.conv(Src.nhwc(actv, h, w), wgb, gb, B=B, Hin=h, Win=w, Hout=h, Wout=w, Cout=2 * C, ksize=3, pad=1, bias=bgb,
               tag="spade.gamma_beta")
        P.release(actv)
        return gb

    def _resblock(self, rb, xs, cs, h, w):
        S, B = self.step, self.B
        hw = h * w
        cout = rb.out_channels
        h1 = S.buf(B, hw, cout)
        off = self._emb_off[id(rb)]
        kw1 = dict(B=B, Hin=h, Win=w, Hout=h, Wout=w, ksize=3, pad=1, bias=self._vec(rb.in_layers[2].bias),
                   rowvec=self.emb_all[:, off:], rowvec_sb=self.emb_total, tag="res.conv1")
        nf = self._norm_on_load(S, xs, cs, h, w, rb.in_layers[0], 1e-5, 1, h1, cout, 3)
        if nf is not None:  # GroupNorm -> [SPADE] -> SiLU applied by conv1 on load (pyunet.py:209-212 as one launch)
            self._conv_stats(S, Src.nhwc(xs[0], h, w), self._conv_w(rb.in_layers[2]), h1, cout,
                             a1=Src.nhwc(xs[1], h, w) if len(xs) > 1 else None, nrm=nf[0], **kw1)
            S.release(nf[1])
        else:
            sp = self._norm_mode(S, xs, cs, h, w, rb.in_layers[0], h1, cout, 3) == "split"
            t1 = self._norm(S, xs, cs, h, w, rb.in_layers[0], 1e-5, 1, "res.norm1", out_split=int(sp))
            self._conv_stats(S, Src.nhwc(t1, h, w), self._conv_w(rb.in_layers[2]), h1, cout, presplit=sp, **kw1)
            S.release(t1)
        a1 = Src.nhwc(xs[1], h, w) if len(xs) > 1 else None
        out = S.buf(B, hw, cout)
        side2 = isinstance(rb.skip_connection, nn.Conv2d) and os.environ.get("FRIDO_FUSE_SKIP", "1") == "1"
        nf = self._norm_on_load(S, [h1], [cout], h, w, rb.out_layers[0], 1e-5, 1, out, cout, 3, side=side2)
        if nf is not None:
            src2, t2 = Src.nhwc(h1, h, w), None
        else:
            sp2 = self._norm_mode(S, [h1], [cout], h, w, rb.out_layers[0], out, cout, 3, side=side2) == "split"
            t2 = self._norm(S, [h1], [cout], h, w, rb.out_layers[0], 1e-5, 1, "res.norm2", out_split=int(sp2))
            S.release(h1)
            src2 = Src.nhwc(t2, h, w)
        nrm2 = nf[0] if nf is not None else None
        sp2 = nf is None and sp2
        sk = None
        is_conv = isinstance(rb.skip_connection, nn.Conv2d)
        if (is_conv and os.environ.get("FRIDO_FUSE_SKIP", "1") == "1" and all(c % 32 == 0 for c in cs) and
                S.tc_eligible(src2, None, out, B=B, Hin=h, Win=w, Hout=h, Wout=w, Cout=cout, ksize=3, pad=1)):
            # the 1x1 skip_connection conv (pyunet.py:248,299) rides on conv2's K loop: its input channels are extra
            # K steps read at the output pixel, its weights extra columns, the biases add up - no separate launch, no
            # round trip of the skip tensor through HBM
            sc_, c2_ = rb.skip_connection, rb.out_layers[3]
            w_cat = self._packed(lambda: PK.conv_plus_side(c2_.weight, sc_.weight))
            b_cat = self._packed(lambda: PK.vec_add(c2_.bias, sc_.bias))
            self._conv_stats(S, src2, w_cat, out, cout, B=B, Hin=h, Win=w, Hout=h, Wout=w, ksize=3, pad=1,
                             bias=b_cat, side=(Src.nhwc(xs[0], h, w), a1), nrm=nrm2, presplit=sp2, tag="res.conv2+skip")
        else:
            if is_conv:
                sk = S.buf(B, hw, cout)
                S.conv(Src.nhwc(xs[0], h, w), self._conv_w(rb.skip_connection), sk, B=B, Hin=h, Win=w, Hout=h, Wout=w,
                       Cout=cout, a1=a1, bias=self._vec(rb.skip_connection.bias), tag="res.skip")
                res = sk
            else:
                assert len(xs) == 1
                res = xs[0]
            self._conv_stats(S, src2, self._conv_w(rb.out_layers[3]), out, cout, B=B, Hin=h, Win=w, Hout=h, Wout=w,
                             ksize=3, pad=1, bias=self._vec(rb.out_layers[3].bias), res=res, nrm=nrm2, presplit=sp2, tag="res.conv2")
        if nf is not None:
            S.release(nf[1])
            S.release(h1)
        else:
            S.release(t2)
        if sk is not None:
            S.release(sk)
        return out

    def _attention(self, x, C, N, ca, ctx_kv, tag):
        """x: LayerNorm-ed tokens [B,N,C]; returns attention output before to_out."""
        S, B = self.step, self.B
        scale = float(C) ** -0.5
        if ctx_kv is None and self._small_attn(N, C):
            # fewer than 128 tokens per image (8x8 level): one q|k|v GEMM + the fused short-sequence attention kernel
            wqkv = self._packed(lambda: PK.cat_rows([ca.to_q.weight, ca.to_k.weight, ca.to_v.weight]))
            qkv = S.buf(B, N, 3 * C)
            S.linear(x, wqkv, qkv, M=B * N, K=C, N=3 * C, tag=tag + ".qkv")
            o = S.buf(B, N, C)
            S.attn_small(qkv, qkv, qkv, o, B=B, N=N, Nk=N, Cdim=C, scale=scale, q_sb=N * 3 * C, q_ld=3 * C, k_off=C,
                         k_sb=N * 3 * C, k_ld=3 * C, v_off=2 * C, v_sb=N * 3 * C, v_ld=3 * C, tag=tag + ".fused")
            S.release(qkv)
            return o
        if ctx_kv is None:  # self-attention: fused q|k projection, V written transposed
            wqk = self._packed(lambda: PK.cat_rows([ca.to_q.weight, ca.to_k.weight]))
            qk = S.buf(B, N, 2 * C)
            # BF16x3: K and V^T later act as the W operand of QK^T / PV, so their producers also emit the bf16 hi/lo pair
            pair = S.tc_code == 3 and N >= 128
            qk_pair = (S.buf(B, N, 2 * C, dtype=torch.bfloat16), S.buf(B, N, 2 * C, dtype=torch.bfloat16)) if pair else None
            S.linear(x, wqk, qk, M=B * N, K=C, N=2 * C, round_tf32=S.R, out_pair=qk_pair, tag=tag + ".qk")
            vT = S.buf(B, C, N)
            vT_pair = (S.buf(B, C, N, dtype=torch.bfloat16), S.buf(B, C, N, dtype=torch.bfloat16)) if pair else None
            S.conv(Src(x, C, N * C, 0, C, 1), self._vec(ca.to_v.weight), vT, B=B, Hin=1, Win=N, Hout=1, Wout=N, Cout=C,
                   o_sb=C * N, o_sp=1, o_sn=N, round_tf32=S.R, out_pair=vT_pair, tag=tag + ".vT")
            Nk, Nkp = N, N
            q_src = Src(qk, C, N * 2 * C, 0, 2 * C, 1)
            k_t, k_off, k_sb, k_ld = qk, C, N * 2 * C, 2 * C
            v_t, v_sb = vT, C * N
            k_pair, v_pair = qk_pair, vT_pair
        else:
            kc, vTc, k_pair, v_pair = ctx_kv
            q = S.buf(B, N, C)
            S.linear(x, self._vec(ca.to_q.weight), q, M=B * N, K=C, N=C, round_tf32=S.R, tag=tag + ".q")
            Nk, Nkp = self.Lc, self.Lp
            q_src = Src(q, C, N * C, 0, C, 1)
            k_t, k_off, k_sb, k_ld = kc, 0, self.Lp * C, C
            v_t, v_sb = vTc, C * self.Lp
            qk = q
        # scores: pad columns [Nk, Nkp) stay zero forever (zero-initialised, never written)
        sc = torch.zeros(B, N, Nkp, dtype=torch.float32, device=self.dev) if Nkp != Nk else S.buf(B, N, Nkp)
        S.hold(sc)
        # cross-attention: computing the (zero) padded key columns too makes C_out a multiple of 64 -> tensor-core eligible
        n_cols = Nkp if (Nkp != Nk and S.tc_code and N >= 128) else Nk
        S.conv(q_src, k_t, sc, B=B, Hin=1, Win=N, Hout=1, Wout=N, Cout=n_cols, w_sb=k_sb, w_ld=k_ld, w_off=k_off,
               o_sb=N * Nkp, o_sp=Nkp, w_pair=k_pair, alg_flops=2 * B * N * Nk * C, tag=tag + ".qk^T")
        S.softmax(sc, rows=B * N, n=Nk, ld=Nkp, scale=scale, round_tf32=S.R, tag=tag + ".softmax")
        o = S.buf(B, N, C)
        # K runs over the padded key count when that makes the tensor-core engine eligible (pad columns are zero)
        Kpv = Nkp if Nkp % 32 == 0 else Nk
        S.conv(Src(sc, Kpv, N * Nkp, 0, Nkp, 1), v_t, o, B=B, Hin=1, Win=N, Hout=1, Wout=N, Cout=C, w_sb=v_sb, w_ld=Nkp,
               round_tf32=S.R, w_pair=v_pair, alg_flops=2 * B * N * Nk * C, tag=tag + ".pv")
        S.release(qk)
        if ctx_kv is None:
            S.release(vT)
            if qk_pair is not None:
                for t_ in qk_pair + vT_pair:
                    S.release(t_)
        if Nkp == Nk:
            S.release(sc)
        return o

    @staticmethod
    def _fold_self_attn():
        import os
        return os.environ.get("FRIDO_ATTN_FOLD", "1") == "1"

    def _self_attention_folded(self, hcur, blk, C, N):
        """h1 = hcur + to_out(attn1(LN(hcur))) (attention.py:323) with the weight products folded at pack time
        (attention.py:172-191 re-associated; x = LN(hcur)):
          sim = (x Wq^T)(x Wk^T)^T = (x (Wk^T Wq)^T) x^T        -> one C x C projection, the keys are x itself
          to_out(P (x Wv^T)) = P (x (Wo Wv)^T) + b_o            -> the values carry to_out; bias + residual ride on PV
        which removes the to_out GEMM and the key projection from every step."""
        S, B = self.step, self.B
        ca = blk.attn1
        scale = float(C) ** -0.5
        def fold_a():  # [C,C] weight of x -> x (Wk^T Wq)^T
            return fold_self_attention(ca)[0]

        def fold_v():  # [C,C] weight of x -> x (Wo Wv)^T
            return fold_self_attention(ca)[1]

        b_o = self._vec(ca.to_out[0].bias)
        ln = S.buf(B, N, C)
        h1 = S.buf(B, N, C)
        if self._small_attn(N, C):  # fewer than 128 tokens per image (8x8 level): fused SIMT attention kernel
            S.layernorm(hcur, self._vec(blk.norm1.weight), self._vec(blk.norm1.bias), ln, rows=B * N, Cdim=C)
            w_av = self._packed(lambda: PK.cat_rows(list(fold_self_attention(ca))))
            qv = S.buf(B, N, 2 * C)
            S.linear(ln, w_av, qv, M=B * N, K=C, N=2 * C, tag="attn1.qv")
            S.attn_small(qv, ln, qv, h1, B=B, N=N, Nk=N, Cdim=C, scale=scale, q_sb=N * 2 * C, q_ld=2 * C, k_sb=N * C, k_ld=C,
                         v_off=C, v_sb=N * 2 * C, v_ld=2 * C, bias=b_o, res=hcur, tag="attn1.fused")
            S.release(qv); S.release(ln)
            return h1
        pair = S.tc_code == 3
        bf = dict(dtype=torch.bfloat16)
        ln_pair = (S.buf(B, N, C, **bf), S.buf(B, N, C, **bf)) if pair else None
        S.layernorm(hcur, self._vec(blk.norm1.weight), self._vec(blk.norm1.bias), ln, rows=B * N, Cdim=C, round_tf32=S.R,
                    out_pair=ln_pair)
        flash = pair and S.flash_eligible(B, N, C)
        t = S.buf(B, N, C)
        t_pair = (S.buf(B, N, C, **bf), S.buf(B, N, C, **bf)) if flash else None
        S.linear(ln, self._packed(fold_a), t, M=B * N, K=C, N=C, round_tf32=S.R, out_pair=t_pair, tag="attn1.q")
        # V'^T for all images at once, [C, B*N]: the folded value weights (Wo Wv) are the A operand (C rows) and the
        # tokens x play the weight matrix ([B*N rows][K=C], bf16 pair from the LayerNorm) - a dense row-major store
        # instead of a transposing epilogue.  Image b's V'^T is the column block [b*N, (b+1)*N).
        vT = S.buf(C, B * N)
        vT_pair = (S.buf(C, B * N, **bf), S.buf(C, B * N, **bf)) if pair else None
        S.conv(Src(self._packed(fold_v), C, 0, 0, C, 1), ln, vT, B=1, Hin=1, Win=C, Hout=1, Wout=C, Cout=B * N, w_ld=C,
               round_tf32=S.R, out_pair=vT_pair, w_pair=ln_pair, tag="attn1.vT")
        if flash:  # scores, softmax and PV in one launch (csrc/attn_flash.cu): no [B,N,N] tensor
            S.flash(t_pair, ln_pair, vT_pair, h1, B=B, N=N, Cdim=C, scale=scale, vt_sb=N, vt_ld=B * N, bias=b_o, res=hcur,
                    tag="attn1.flash")
            for t_ in (t, vT, ln) + ln_pair + vT_pair + t_pair:
                S.release(t_)
            return h1
        sc = S.buf(B, N, N)
        S.conv(Src(t, C, N * C, 0, C, 1), ln, sc, B=B, Hin=1, Win=N, Hout=1, Wout=N, Cout=N, w_sb=N * C, w_ld=C,
               o_sb=N * N, o_sp=N, w_pair=ln_pair, tag="attn1.qk^T")
        S.softmax(sc, rows=B * N, n=N, ld=N, scale=scale, round_tf32=S.R, tag="attn1.softmax")
        S.conv(Src(sc, N, N * N, 0, N, 1), vT, h1, B=B, Hin=1, Win=N, Hout=1, Wout=N, Cout=C, w_sb=N, w_ld=B * N,
               w_pair=vT_pair, bias=b_o, res=hcur, tag="attn1.pv")
        for t_ in (t, vT, sc, ln) + (ln_pair + vT_pair if pair else ()):
            S.release(t_)
        return h1

    def _small_attn(self, Nk, C):
        """Key sequences too short for a 128-row tensor-core tile go to the fused SIMT attention kernel (csrc/attn.cu)."""
        import os
        return Nk <= 64 and C <= 1024 and C % 4 == 0 and os.environ.get("FRIDO_ATTN_SMALL", "1") == "1"

    def _ctx_folded(self, ca, C):
        """Prologue, short condition: the score and output operands of the fused cross-attention kernel with the
        step-invariant projections folded in (attention.py:172-191 re-associated):
          sim = (x Wq^T)(ctx Wk^T)^T = x ((ctx Wk^T) Wq)^T        -> k' = (ctx Wk^T) Wq    [B,Lc,C]
          to_out(P (ctx Wv^T)) = P ((ctx Wv^T) Wo^T) + b_o        -> v' = (ctx Wv^T) Wo^T  [B,Lc,C]"""
        P, B, Lc = self.prologue, self.B, self.Lc
        D = self.net.context_dim
        kf = torch.empty(B, Lc, C, dtype=torch.float32, device=self.dev)  # live across steps: not pooled
        vf = torch.empty(B, Lc, C, dtype=torch.float32, device=self.dev)
        kc = P.buf(B, Lc, C)
        P.linear(self.ctx, self._vec(ca.to_k.weight), kc, M=B * Lc, K=D, N=C, tag="ctx.k")
        P.linear(kc, self._packed(lambda: fold_cross_attention_weights(ca)[0]), kf, M=B * Lc, K=C, N=C, tag="ctx.k.Wq")
        P.linear(self.ctx, self._vec(ca.to_v.weight), kc, M=B * Lc, K=D, N=C, tag="ctx.v")
        P.linear(kc, self._packed(lambda: fold_cross_attention_weights(ca)[1]), vf, M=B * Lc, K=C, N=C, tag="ctx.v.Wo")
        P.release(kc)
        return kf, vf

    def _ctx_kv(self, ca, C):
        """Prologue: K = ctx Wk^T [B,Lp,C] (rows >= Lc zero), V^T [B,C,Lp] (attention.py:175-176)."""
        P, B, Lc, Lp = self.prologue, self.B, self.Lc, self.Lp
        D = self.net.context_dim
        kc = torch.zeros(B, Lp, C, dtype=torch.float32, device=self.dev)
        vT = torch.zeros(B, C, Lp, dtype=torch.float32, device=self.dev)
        k_pair = v_pair = None
        if self.step.tc_code == 3:  # bf16 hi/lo copies (zero pad rows/columns stay zero)
            k_pair = tuple(torch.zeros(B, Lp, C, dtype=torch.bfloat16, device=self.dev) for _ in range(2))
            v_pair = tuple(torch.zeros(B, C, Lp, dtype=torch.bfloat16, device=self.dev) for _ in range(2))
        P.conv(Src(self.ctx, D, Lc * D, 0, D, 1), self._vec(ca.to_k.weight), kc, B=B, Hin=1, Win=Lc, Hout=1, Wout=Lc,
               Cout=C, o_sb=Lp * C, o_sp=C, round_tf32=P.R, out_pair=k_pair, tag="ctx.k")
        P.conv(Src(self.ctx, D, Lc * D, 0, D, 1), self._vec(ca.to_v.weight), vT, B=B, Hin=1, Win=Lc, Hout=1, Wout=Lc,
               Cout=C, o_sb=C * Lp, o_sp=1, o_sn=Lp, round_tf32=P.R, out_pair=v_pair, tag="ctx.vT")
        return kc, vT, k_pair, v_pair

    def _transformer(self, st, x, C, h, w):
        S, B = self.step, self.B
        N = h * w
        hcur = S.buf(B, N, C)
        w_in = self._packed(lambda: PK.copy(st.proj_in.weight.detach().view(C, C)))
        nf = self._norm_on_load(S, [x], [C], h, w, st.norm, 1e-6, 0, hcur, C, 1)
        if nf is not None:  # GroupNorm [+SPADE] applied by proj_in on load (attention.py:296-298 as one launch)
            S.conv(Src.nhwc(x, h, w), w_in, hcur, B=B, Hin=h, Win=w, Hout=h, Wout=w, Cout=C, bias=self._vec(st.proj_in.bias),
                   nrm=nf[0], tag="st.proj_in")
            S.release(nf[1])
        else:
            t = self._norm(S, [x], [C], h, w, st.norm, 1e-6, 0, "st.norm")
            S.linear(t, w_in, hcur, M=B * N, K=C, N=C, bias=self._vec(st.proj_in.bias), tag="st.proj_in")
            S.release(t)
        for blk in st.transformer_blocks:
            if self._fold_self_attn():
                h1 = self._self_attention_folded(hcur, blk, C, N)
            else:
                ln = S.buf(B, N, C)
                S.layernorm(hcur, self._vec(blk.norm1.weight), self._vec(blk.norm1.bias), ln, rows=B * N, Cdim=C, round_tf32=S.R)
                o = self._attention(ln, C, N, blk.attn1, None, "attn1")
                S.release(ln)
                h1 = S.buf(B, N, C)
                S.linear(o, self._vec(blk.attn1.to_out[0].weight), h1, M=B * N, K=C, N=C,
                         bias=self._vec(blk.attn1.to_out[0].bias), res=hcur, tag="attn1.out")
                S.release(o)
            S.release(hcur)
            ln3 = None
            if self._small_attn(self.Lc, C):
                # h2 = h1 + to_out(attn2(LN(h1), ctx)) in one launch: LayerNorm, 26-key attention against the folded
                # operands, output bias and residual (attention.py:324)
                kf, vf = self._ctx_folded(blk.attn2, C)
                h2 = S.buf(B, N, C)
                ln3 = S.buf(B, N, C)  # LayerNorm(h2) for the feed-forward, produced by the same launch
                S.attn_small(h1, kf, vf, h2, B=B, N=N, Nk=self.Lc, Cdim=C, scale=float(C) ** -0.5, q_sb=N * C, q_ld=C,
                             k_sb=self.Lc * C, k_ld=C, v_sb=self.Lc * C, v_ld=C,
                             ln=(self._vec(blk.norm2.weight), self._vec(blk.norm2.bias)),
                             bias=self._vec(blk.attn2.to_out[0].bias), res=h1,
                             ln2=(self._vec(blk.norm3.weight), self._vec(blk.norm3.bias)), out2=ln3, tag="attn2.block")
                S.release(h1)
            else:
                ln = S.buf(B, N, C)
                S.layernorm(h1, self._vec(blk.norm2.weight), self._vec(blk.norm2.bias), ln, rows=B * N, Cdim=C, round_tf32=S.R)
                o = self._attention(ln, C, N, blk.attn2, self._ctx_kv(blk.attn2, C), "attn2")
                S.release(ln)
                h2 = S.buf(B, N, C)
                S.linear(o, self._vec(blk.attn2.to_out[0].weight), h2, M=B * N, K=C, N=C,
                         bias=self._vec(blk.attn2.to_out[0].bias), res=h1, tag="attn2.out")
                S.release(o); S.release(h1)
            if ln3 is not None:
                ln = ln3
            else:
                ln = S.buf(B, N, C)
                S.layernorm(h2, self._vec(blk.norm3.weight), self._vec(blk.norm3.bias), ln, rows=B * N, Cdim=C, round_tf32=S.R)
            proj = blk.ff.net[0].proj
            inner = proj.weight.shape[0] // 2

            def _inter(p=proj, inner=inner):  # GEGLU (attention.py:42-44): rows (value_j, gate_j) interleaved
                return PK.interleave_rows(p.weight.detach()[:inner], p.weight.detach()[inner:])

            def _inter_b(p=proj, inner=inner):
                return PK.interleave_rows(p.bias.detach()[:inner], p.bias.detach()[inner:])

            ff = S.buf(B, N, inner)
            S.linear(ln, self._packed(_inter), ff, M=B * N, K=C, N=2 * inner, bias=self._packed(_inter_b), act=L.ACT_GEGLU,
                     round_tf32=S.R, tag="ff.geglu")
            S.release(ln)
            h3 = S.buf(B, N, C)
            S.linear(ff, self._vec(blk.ff.net[2].weight), h3, M=B * N, K=inner, N=C, bias=self._vec(blk.ff.net[2].bias),
                     res=h2, tag="ff.out")
            S.release(ff); S.release(h2)
            hcur = h3
        out = S.buf(B, N, C)
        # same GEMM as a 1x1 conv over [B,h,w,C] (so the epilogue can attribute rows to images for the GroupNorm sums)
        self._conv_stats(S, Src.nhwc(hcur, h, w), self._packed(lambda: PK.copy(st.proj_out.weight.detach().view(C, C))), out, C,
                         B=B, Hin=h, Win=w, Hout=h, Wout=w, bias=self._vec(st.proj_out.bias), res=x, tag="st.proj_out")
        S.release(hcur)
        return out

    def _run_seq(self, seq, xs, cs, h, w):
        """One TimestepEmbedSequential (pyunet.py:75-91). xs: list of NHWC sources (skip concat)."""
        S, B = self.step, self.B
        for layer in seq:
            if isinstance(layer, M.ResBlock):
                x = self._resblock(layer, xs, cs, h, w)
                xs, cs = [x], [layer.out_channels]
            elif isinstance(layer, M.SpatialTransformer):
                x = self._transformer(layer, xs[0], cs[0], h, w)
                xs = [x]
            elif isinstance(layer, M.Downsample):
                ho, wo = (h + 1) // 2, (w + 1) // 2
                x = S.buf(B, ho * wo, cs[0])
                self._conv_stats(S, Src.nhwc(xs[0], h, w), self._conv_w(layer.op), x, cs[0], B=B, Hin=h, Win=w, Hout=ho, Wout=wo,
                                 ksize=3, stride=2, pad=1, bias=self._vec(layer.op.bias), tag="down")
                xs, h, w = [x], ho, wo
            elif isinstance(layer, M.Upsample):
                x = S.buf(B, 4 * h * w, cs[0])
                if S.tc_code and cs[0] % 64 == 0:
                    # tensor-core path: materialise the nearest x2 copy (TF32-rounded), then a plain 3x3 conv
                    up = S.buf(B, 4 * h * w, cs[0])
                    # the x2 copy is written in the engine's operand form when the conv can take its halo-resident path
                    sp = (os.environ.get("FRIDO_FUSE_NORM", "auto") != "0" and cs[0] % 32 == 0 and
                          S.nf_eligible(Src.nhwc(up, 2 * h, 2 * w), None, x, B=B, H=2 * h, W=2 * w, Cout=cs[0], ksize=3))
                    S.upsample2x(xs[0], up, B=B, H=h, W=w, Cdim=cs[0], round_tf32=S.R, out_split=int(sp))
                    self._conv_stats(S, Src.nhwc(up, 2 * h, 2 * w), self._conv_w(layer.conv), x, cs[0], B=B, Hin=2 * h,
                                     Win=2 * w, Hout=2 * h, Wout=2 * w, ksize=3, pad=1, bias=self._vec(layer.conv.bias),
                                     presplit=sp, tag="up")
                    S.release(up)
                else:
                    S.conv(Src.nhwc(xs[0], h, w), self._conv_w(layer.conv), x, B=B, Hin=h, Win=w, Hout=2 * h, Wout=2 * w,
                           Cout=cs[0], ksize=3, pad=1, ups=2, bias=self._vec(layer.conv.bias), tag="up")
                xs, h, w = [x], 2 * h, 2 * w
            else:
                raise TypeError(type(layer))
        return xs[0], cs[0], h, w

    # ---- the program -----------------------------------------------------
    def _build(self):
        net, B, H, W, s = self.net, self.B, self.H, self.W, self.stage
        P, S = self.prologue, self.step
        mc = net.model_channels
        ted = mc * 4
        # GN statistics slots: one [B,32,2] fp64 block per norm site, zeroed once per step
        n_norm = sum(1 for m in net.modules() if isinstance(m, (nn.GroupNorm,)))
        self._sums_all = torch.zeros(n_norm + 2, B, 32, 2, dtype=torch.float64, device=self.dev)
        S.zero(self._sums_all, tag="gn.zero")
        # producer-side GroupNorm statistics: per-channel (sum, sumsq) blocks written by conv epilogues, zeroed per step
        tot_c = sum(m.num_channels for m in net.modules() if isinstance(m, nn.GroupNorm))
        self._csum_pool = torch.zeros(B * tot_c * 2 + 16, dtype=torch.float64, device=self.dev)
        self._csum_used = 0
        self._csum = {}
        S.zero(self._csum_pool, tag="gn.zero")
        # --- embedding (a7) ---
        te = S.buf(B, mc)
        S.time_embed(self.ts, te, B=B, dim=mc)
        e1 = S.buf(B, ted)
        S.linear(te, self._vec(net.time_embed[0].weight), e1, M=B, K=mc, N=ted, bias=self._vec(net.time_embed[0].bias),
                 act=L.ACT_SILU, tag="time_embed.0")
        semb = S.buf(B, ted)  # SiLU(emb): every consumer applies SiLU first (pyunet.py:226)
        stage_row = self._packed(lambda: PK.copy(net.stage_emb.weight.detach()[s])) if net.num_stage > 1 else None
        S.linear(e1, self._vec(net.time_embed[2].weight), semb, M=B, K=ted, N=ted, bias=self._vec(net.time_embed[2].bias),
                 rowvec=stage_row, act=L.ACT_SILU, tag="time_embed.2")
        rbs = [m for m in net.modules() if isinstance(m, M.ResBlock)]
        self._emb_off, off = {}, 0
        for rb in rbs:
            self._emb_off[id(rb)] = off
            off += rb.out_channels
        self.emb_total = off
        self.emb_all = S.buf(B, off)
        S.linear(semb, self._packed(lambda: PK.cat_rows([rb.emb_layers[1].weight for rb in rbs])), self.emb_all,
                 M=B, K=ted, N=off, bias=self._packed(lambda: PK.cat_rows([rb.emb_layers[1].bias for rb in rbs])),
                 tag="emb_layers(all)")
        # --- split head (pyunet.py:899-914) ---
        if self.c_cond:
            pc = net.pre_input_cond_blocks[s - 1][0]
            self.h_cond = torch.empty(B, H * W, mc, dtype=torch.float32, device=self.dev)
            P.conv(Src.nchw(self.x_in, H, W, 0, self.c_cond), self._conv_w(pc), self.h_cond, B=B, Hin=H, Win=W, Hout=H,
                   Wout=W, Cout=mc, ksize=3, pad=1, bias=self._vec(pc.bias), tag="pre_input_cond")
        pi = net.pre_input_blocks[s][0]
        c_lo = self.c_cond if self.spade else 0
        h = S.buf(B, H * W, mc)
        S.conv(Src.nchw(self.x_in, H, W, c_lo, self.c_end), self._conv_w(pi), h, B=B, Hin=H, Win=W, Hout=H, Wout=W,
               Cout=mc, ksize=3, pad=1, bias=self._vec(pi.bias), tag="pre_input")
        # pre_input runs on the SIMT engine (3 input channels): one per-channel statistics pass over its output, so that
        # both GroupNorms that read it (first ResBlock; last skip concat) take the producer-statistics path
        cs_h = self._csum_new(mc)
        if cs_h is not None and mc % 4 == 0:
            S.chan_stats(h, mc, cs_h, B=B, HW=H * W, tag="pre_input.stats")
            self._csum[id(h)] = cs_h
        hs = [(h, mc, H, W)]
        c, hh, ww = mc, H, W
        for blk in net.input_blocks:
            h, c, hh, ww = self._run_seq(blk, [h], [c], hh, ww)
            hs.append((h, c, hh, ww))
        h, c, hh, ww = self._run_seq(net.middle_block, [h], [c], hh, ww)
        for blk in net.output_blocks:
            sk, sc, sh, sw = hs.pop()
            assert (sh, sw) == (hh, ww)
            h, c, hh, ww = self._run_seq(blk, [h, sk], [c, sc], hh, ww)
        # --- out head (pyunet.py:947) ---
        oh = net.out[s]
        t = self._norm(S, [h], [c], hh, ww, oh[0], 1e-5, 1, "out.norm")
        S.conv(Src.nhwc(t, hh, ww), self._conv_w(oh[2]), self.eps, B=B, Hin=hh, Win=ww, Hout=hh, Wout=ww, Cout=self.e_s,
               ksize=3, pad=1, bias=self._vec(oh[2].bias), o_sb=self.e_s * H * W, o_sp=1, o_sn=H * W, tag="out.conv")
