"""MS-VQGAN decode side — drop-in mirror of taming/models/msvqgan.py
(`VQModelInterface.decode`, :376-399) with the taming Decoder
(taming/modules/diffusionmodules/model.py:548-649) and VectorQuantizer2
(taming/modules/vqvae/quantize.py:267-308) scheduled as one libfrido_b200 program:

  per-scale rescale + VQ argmin/gather (written straight into the fine->coarse
  concatenated NHWC tensor) -> post_quant_conv -> conv_in -> mid (Res, Attn, Res)
  -> up levels (Res [+Attn]) with nearest-x2 folded into the upsample conv
  -> GroupNorm+swish -> conv_out written NCHW.

The encode side (msvqgan.py:326-374 `VQModelInterface.encode`,
model.py:435-546 MSEncoder) is a second program (`EncodePlan`): the bottom-up
encoder with one mid/out head per scale, then the coarse->fine top-down pass
(VQ of the coarser scale -> ConvTranspose2d x2 -> 1x1 conv -> concat with the
finer head -> shared decoder -> ms_quant_conv), and the assembly of the
pre-quantisation maps into the NCHW latent with the per-scale scale_factor
(frido.py:646-662 get_first_stage_encoding) fused into the last kernel.
"""
import torch
from torch import nn

from . import _lib as L
from . import modules as M
from .program import Program, Src
from .program import round_tf32_
from .unet import _pack_conv
from . import packing as PK


@torch.no_grad()
def images_to_uint8(img, mode="np"):
    """Decoded images [B,C,H,W] fp32 -> uint8 NHWC on the device, byte-identical to the reference's
    `custom_to_np` (mode "np", scripts/sample_diffusion.py:115-121) or `custom_to_pil` (mode "pil", :103-108)."""
    import ctypes as C
    if not img.is_cuda:
        raise L.FridoError("images_to_uint8 runs on a CUDA device only (no CPU path)")
    img = img.contiguous().float()
    B, Cc, H, W = img.shape
    out = torch.empty(B, H, W, Cc, dtype=torch.uint8, device=img.device)
    p = L.ToU8Params()
    p.x, p.B, p.C, p.HW, p.mode, p.out = img.data_ptr(), B, Cc, H * W, {"np": 0, "pil": 1}[mode], out.data_ptr()
    s = torch.cuda.current_stream(img.device).cuda_stream
    L.check(L.lib().frido_to_uint8(C.byref(p), C.c_void_p(s)), "to_uint8")
    return out


class VQModelInterface(nn.Module):
    def __init__(self, embed_dim, edconfig=None, ddconfig=None, lossconfig=None, n_embed=None, channel_range=[],
                 fusion="concat", ckpt_path=None, ignore_keys=[], image_key="image", colorize_nlabels=None, monitor=None,
                 remap=None, sane_index_shape=False, on_vit=[], use_aux_loss=False, unsample_type="nearest",
                 quant_beta=0.25, legacy=True, init_normal=False, **kwargs):
        super().__init__()
        if remap is not None or fusion != "concat":
            raise NotImplementedError("remap / non-concat fusion are outside the B200 hot path")
        self.embed_dim = [int(e) for e in embed_dim]
        self.n_embed = [int(n) for n in n_embed]
        assert len(self.n_embed) == len(self.embed_dim), "multiscale mode. dim of n_embed is incorrect."
        self.channel_range = channel_range
        self.image_key = image_key
        self.ddconfig = dict(ddconfig)
        self.edconfig = dict(edconfig) if edconfig is not None else None
        self.decoder = M.TDecoder(**self.ddconfig)
        self.ms_quantize = nn.ModuleList([M.VectorQuantizer(n, e, init_normal) for n, e in zip(self.n_embed, self.embed_dim)])
        self.post_quant_conv = nn.Conv2d(sum(self.embed_dim), self.ddconfig["z_channels"], 1)
        self.unsample_type = unsample_type
        if self.edconfig is not None:  # encode side (msvqgan.py:40,63-88)
            ed = self.edconfig
            assert len(self.n_embed) == ed["multiscale"], "multiscale mode. dim of n_embed is incorrect."
            self.encoder = M.MSEncoder(**ed)
            dz = 2 if ed.get("double_z", True) else 1
            self.ms_quant_conv = nn.ModuleList([nn.Conv2d(dz * ed["z_channels"][i], e, 1) for i, e in enumerate(self.embed_dim)])
            e0 = self.embed_dim[0]
            self.upsample = nn.ModuleList([nn.ConvTranspose2d(e0, e0, 4, stride=2, padding=1) for _ in self.embed_dim[1:]])
            self.shared_post_quant_conv = nn.ModuleList([nn.Conv2d(e0, ed["z_channels"][0], 1) for _ in self.embed_dim[1:]])
            self.shared_decoder = nn.ModuleList([
                M.TDecoder(ch=128, out_ch=e0, ch_mult=[1], num_res_blocks=2, attn_resolutions=[2, 4, 8, 16, 32, 64],
                           resolution=256, z_channels=sum(self.embed_dim[:i + 2])) for i in range(len(self.embed_dim) - 1)])
        # frido.py:609 reads len(first_stage_model.res_list)
        if self.edconfig is not None:
            nres = len(self.edconfig["ch_mult"])
            self.res_list = [self.edconfig["resolution"] / 2 ** (nres - i - 1) for i in range(self.edconfig["multiscale"])]
        else:
            self.res_list = [0] * len(self.embed_dim)
        self._plans = {}
        self._pack_version = 0
        if ckpt_path is not None:
            self.init_from_ckpt(ckpt_path, ignore_keys=ignore_keys)

    def init_from_ckpt(self, path, ignore_keys=list()):
        sd = torch.load(path, map_location="cpu")["state_dict"]
        for k in list(sd.keys()):
            if any(k.startswith(ik) for ik in ignore_keys):
                del sd[k]
        missing, unexpected = self.load_state_dict(sd, strict=False)
        print(f"Restored from {path} with {len(missing)} missing and {len(unexpected)} unexpected keys")
        self.invalidate()

    def invalidate(self):
        self._pack_version += 1
        for p in self._plans.values():
            p.repack()

    @torch.no_grad()
    def encode(self, x, scale_factor=None, return_code=False):
        """msvqgan.py:326-374: image [B,3,H,W] -> pre-quantisation multi-scale latent [B, sum(embed_dim), H/f, W/f]
        (coarse scale first, nearest-upsampled to the finest latent resolution).  `scale_factor` (one per scale)
        folds get_first_stage_encoding's per-group multiply (frido.py:656-661) into the assembly kernel."""
        if self.edconfig is None:
            raise L.FridoError("VQModelInterface.encode needs an `edconfig`")
        if not x.is_cuda:
            raise L.FridoError("VQModelInterface.encode runs on a CUDA device only (no CPU path)")
        if len(self.channel_range) == 2 and (self.channel_range[0] != 0 or self.channel_range[1] != sum(self.embed_dim)):
            raise NotImplementedError("partial channel_range is outside the B200 path")
        B, C, H, W = x.shape
        sf = tuple(float(s) for s in (scale_factor if scale_factor is not None else [1.0] * len(self.embed_dim)))
        key = ("enc", B, H, W, sf)
        plan = self._plans.get(key)
        if plan is None:
            plan = EncodePlan(self, B, H, W, sf)
            self._plans[key] = plan
        plan.repack_if_stale()
        plan.x.copy_(x)
        plan.prog.run()
        z = plan.latent.clone()
        if return_code:
            return z, [idx.view(B, -1).clone() for idx in plan.indices]
        return z

    def _decode_plan(self, h_in, scale_factor, u8_mode=0):
        if not h_in.is_cuda:
            raise L.FridoError("VQModelInterface.decode runs on a CUDA device only (no CPU path)")
        B, C, H, W = h_in.shape
        sf = tuple(float(s) for s in (scale_factor if scale_factor is not None else [1.0] * len(self.embed_dim)))
        key = (B, H, W, sf, u8_mode)
        plan = self._plans.get(key)
        if plan is None:
            plan = DecodePlan(self, B, H, W, sf, u8_mode)
            self._plans[key] = plan
        plan.repack_if_stale()
        plan.z.copy_(h_in)
        plan.prog.run()
        return plan

    @torch.no_grad()
    def decode_uint8(self, h_in, scale_factor=None, mode="np", out=None):
        """decode + the script's output formatting in one program: uint8 NHWC [B,H,W,3], byte-identical to
        `custom_to_np(decode(h))` (mode "np", sample_diffusion.py:115-121: what goes into the .npz) or to
        `custom_to_pil` (mode "pil", :103-108: what goes into the PNGs).  The bytes are written by conv_out's epilogue
        from the fp32 value it stores; no fp32 image leaves the device.  `out`: optional destination (e.g. a slot of the
        all-gather buffer)."""
        plan = self._decode_plan(h_in, scale_factor, {"np": 0, "pil": 1}[mode])
        if plan.image_u8 is None:
            raise L.FridoError("decode_uint8: the decoder head has more than 4 output channels")
        if out is not None:
            out.copy_(plan.image_u8)
            return out
        return plan.image_u8.clone()

    @torch.no_grad()
    def decode(self, h_in, force_not_quantize=False, return_code=False, scale_factor=None):
        """msvqgan.py:376-399.  `scale_factor` (list, one per scale) folds
        decode_first_stage's per-group 1/scale (frido.py:832-838) into the VQ kernel."""
        plan = self._decode_plan(h_in, scale_factor)
        B = h_in.shape[0]
        dec = plan.image.clone()
        if return_code:
            code = [idx.view(B, -1).tolist() for idx in plan.indices]  # msvqgan.py:390
            return dec, code
        return dec


class _VQPlan:
    """Shared emitters of the taming blocks (ResnetBlock, AttnBlock, Decoder body) over one Program."""

    def _init_plan(self, fs, B, name, n_gn):
        self.fs, self.B = fs, B
        dev = next(fs.parameters()).device
        self.dev = dev
        self.packers = []
        self.version = fs._pack_version
        self.prog = Program(dev, name)
        self._sums = torch.zeros(n_gn + 1, B, 32, 2, dtype=torch.float64, device=dev)
        self._slot = 0
        self._mma_ids = set()
        self.prog.zero(self._sums, tag="gn.zero")

    def _finish_plan(self):
        P = self.prog
        self._mma_ids = {id(t) for t in P.mma_weights} if P.R else set()
        for dst, _ in self.packers:
            if id(dst) in self._mma_ids:
                round_tf32_(dst)
        P.prepare_weights()

    def _decoder_body(self, dec, zin, H, W, out, nchw_out, tag, out_u8=None, u8_mode=0):
        """taming Decoder.forward (model.py:618-649) from its z input (NHWC [B,HW,zc]) to conv_out, written into `out`
        as NCHW (the image) or NHWC (the shared decoders of the encode side)."""
        P, B = self.prog, self.B
        c = dec.conv_in.weight.shape[0]
        zc = dec.conv_in.weight.shape[1]
        h = P.buf(B, H * W, c)
        P.conv(Src.nhwc(zin, H, W, c_total=zc), self._conv_w(dec.conv_in), h, B=B, Hin=H, Win=W, Hout=H, Wout=W, Cout=c, ksize=3,
               pad=1, bias=self._vec(dec.conv_in.bias), tag=tag + ".conv_in")
        hh, ww = H, W
        h = self._res(dec.mid.block_1, h, hh, ww)
        h = self._attn(dec.mid.attn_1, h, hh, ww)
        h = self._res(dec.mid.block_2, h, hh, ww)
        for lvl in reversed(range(dec.num_resolutions)):
            up = dec.up[lvl]
            for j, blk in enumerate(up.block):
                h = self._res(blk, h, hh, ww)
                if len(up.attn) > 0:
                    h = self._attn(up.attn[j], h, hh, ww)
            if lvl != 0:
                c = up.upsample.conv.weight.shape[0]
                o = P.buf(B, 4 * hh * ww, c)
                if P.tc_code and c % 64 == 0:
                    u2 = P.buf(B, 4 * hh * ww, c)
                    P.upsample2x(h, u2, B=B, H=hh, W=ww, Cdim=c, round_tf32=P.R)
                    P.conv(Src.nhwc(u2, 2 * hh, 2 * ww), self._conv_w(up.upsample.conv), o, B=B, Hin=2 * hh, Win=2 * ww,
                           Hout=2 * hh, Wout=2 * ww, Cout=c, ksize=3, pad=1, bias=self._vec(up.upsample.conv.bias), tag=tag + ".up")
                    P.release(u2)
                else:
                    P.conv(Src.nhwc(h, hh, ww), self._conv_w(up.upsample.conv), o, B=B, Hin=hh, Win=ww, Hout=2 * hh, Wout=2 * ww,
                           Cout=c, ksize=3, pad=1, ups=2, bias=self._vec(up.upsample.conv.bias), tag=tag + ".up")
                P.release(h)
                h, hh, ww = o, 2 * hh, 2 * ww
        c = dec.norm_out.weight.shape[0]
        t = self._gn(h, c, hh, ww, dec.norm_out, 1)
        P.release(h)
        oc = dec.conv_out.weight.shape[0]
        kw = dict(o_sb=oc * hh * ww, o_sp=1, o_sn=hh * ww) if nchw_out else {}
        P.conv(Src.nhwc(t, hh, ww), self._conv_w(dec.conv_out), out, B=B, Hin=hh, Win=ww, Hout=hh, Wout=ww, Cout=oc,
               ksize=3, pad=1, bias=self._vec(dec.conv_out.bias), out_u8=out_u8, u8_mode=u8_mode, tag=tag + ".conv_out", **kw)
        P.release(t)
        return hh, ww

    # packing -------------------------------------------------------------
    def _packed(self, fn):
        dst = fn().to(self.dev).contiguous()
        self.packers.append((dst, fn))
        return dst

    def _vec(self, p):
        return self._packed(lambda: PK.copy(p))

    def _conv_w(self, conv):
        return self._packed(lambda: _pack_conv(conv.weight))

    def repack(self):
        for dst, fn in self.packers:
            PK.place(fn().view(1, -1), dst.view(1, -1))
            if id(dst) in self._mma_ids:
                round_tf32_(dst)
        self.prog.prepare_weights()
        self.version = self.fs._pack_version

    def repack_if_stale(self):
        if self.version != self.fs._pack_version:
            self.repack()

    # blocks --------------------------------------------------------------
    def _gn(self, x, c, h, w, norm, silu):
        P, B = self.prog, self.B
        sums = self._sums[self._slot]
        self._slot += 1
        P.gn_stats(x, c, sums, B=B, HW=h * w)
        out = P.buf(B, h * w, c)
        P.norm_act(x, c, sums, self._vec(norm.weight), self._vec(norm.bias), out, B=B, HW=h * w, eps=1e-6, silu=silu,
                   round_tf32=P.R, tag="dec.norm")
        return out

    def _res(self, rb, x, h, w, tag="dec.res"):
        """taming ResnetBlock (model.py:115-137), temb is None in the decoder."""
        P, B = self.prog, self.B
        cin, cout = rb.in_channels, rb.out_channels
        t1 = self._gn(x, cin, h, w, rb.norm1, 1)
        h1 = P.buf(B, h * w, cout)
        P.conv(Src.nhwc(t1, h, w), self._conv_w(rb.conv1), h1, B=B, Hin=h, Win=w, Hout=h, Wout=w, Cout=cout, ksize=3, pad=1,
               bias=self._vec(rb.conv1.bias), tag=tag + ".conv1")
        P.release(t1)
        t2 = self._gn(h1, cout, h, w, rb.norm2, 1)
        P.release(h1)
        res, sk = x, None
        if cin != cout:
            sk = P.buf(B, h * w, cout)
            P.conv(Src.nhwc(x, h, w), self._packed(lambda: PK.copy(rb.nin_shortcut.weight.detach().view(cout, cin))), sk, B=B,
                   Hin=h, Win=w, Hout=h, Wout=w, Cout=cout, bias=self._vec(rb.nin_shortcut.bias), tag=tag + ".nin")
            res = sk
        out = P.buf(B, h * w, cout)
        P.conv(Src.nhwc(t2, h, w), self._conv_w(rb.conv2), out, B=B, Hin=h, Win=w, Hout=h, Wout=w, Cout=cout, ksize=3, pad=1,
               bias=self._vec(rb.conv2.bias), res=res, tag=tag + ".conv2")
        P.release(t2)
        if sk is not None:
            P.release(sk)
        P.release(x)
        return out

    def _attn(self, ab, x, h, w, tag="dec.attn"):
        """taming AttnBlock (model.py:166-192): single head, d = C, softmax over keys."""
        P, B = self.prog, self.B
        C = ab.q.weight.shape[0]
        N = h * w
        t = self._gn(x, C, h, w, ab.norm, 0)
        wqk = self._packed(lambda: PK.cat_rows([ab.q.weight.detach().view(C, C), ab.k.weight.detach().view(C, C)]))
        bqk = self._packed(lambda: PK.cat_rows([ab.q.bias, ab.k.bias]))
        flash = P.tc_code == 3 and P.flash_eligible(B, N, C)
        bf = dict(dtype=torch.bfloat16)
        qk = P.buf(B, N, 2 * C)
        qk_pair = (P.buf(B, N, 2 * C, **bf), P.buf(B, N, 2 * C, **bf)) if flash else None
        P.linear(t, wqk, qk, M=B * N, K=C, N=2 * C, bias=bqk, round_tf32=P.R, out_pair=qk_pair, tag=tag + ".qk")
        vT = P.buf(B, C, N)
        vT_pair = (P.buf(B, C, N, **bf), P.buf(B, C, N, **bf)) if flash else None
        P.conv(Src(t, C, N * C, 0, C, 1), self._packed(lambda: PK.copy(ab.v.weight.detach().view(C, C))), vT, B=B, Hin=1, Win=N,
               Hout=1, Wout=N, Cout=C, bias=self._vec(ab.v.bias), o_sb=C * N, o_sp=1, o_sn=N, round_tf32=P.R, out_pair=vT_pair,
               tag=tag + ".vT")
        P.release(t)
        o = P.buf(B, N, C)
        if flash:
            # scores, softmax and PV in one launch (csrc/attn_flash.cu): the [B,N,N] tensor (1 GB at 64x64, B=16) never exists
            P.flash(qk_pair, qk_pair, vT_pair, o, B=B, N=N, Cdim=C, scale=float(int(C) ** (-0.5)), vt_sb=C * N, vt_ld=N,
                    q_ld=2 * C, k_off=C, k_ld=2 * C, tag=tag + ".flash")
            for t_ in (qk, vT) + qk_pair + vT_pair:
                P.release(t_)
        else:
            sc = P.buf(B, N, N)
            P.conv(Src(qk, C, N * 2 * C, 0, 2 * C, 1), qk, sc, B=B, Hin=1, Win=N, Hout=1, Wout=N, Cout=N, w_sb=N * 2 * C,
                   w_ld=2 * C, w_off=C, tag=tag + ".qk^T")
            P.softmax(sc, rows=B * N, n=N, ld=N, scale=float(int(C) ** (-0.5)), round_tf32=P.R, tag=tag + ".softmax")
            P.conv(Src(sc, N, N * N, 0, N, 1), vT, o, B=B, Hin=1, Win=N, Hout=1, Wout=N, Cout=C, w_sb=C * N, w_ld=N,
                   round_tf32=P.R, tag=tag + ".pv")
            P.release(qk); P.release(vT); P.release(sc)
        out = P.buf(B, N, C)
        P.linear(o, self._packed(lambda: PK.copy(ab.proj_out.weight.detach().view(C, C))), out, M=B * N, K=C, N=C,
                 bias=self._vec(ab.proj_out.bias), res=x, tag=tag + ".proj_out")
        P.release(o); P.release(x)
        return out


class DecodePlan(_VQPlan):
    def __init__(self, fs: VQModelInterface, B, H, W, sf, u8_mode=0):
        self.H, self.W = H, W
        dec = fs.decoder
        self._init_plan(fs, B, "decode", sum(1 for m in dec.modules() if isinstance(m, nn.GroupNorm)))
        dev, P = self.dev, self.prog
        Ct = sum(fs.embed_dim)
        self.z = torch.zeros(B, Ct, H, W, dtype=torch.float32, device=dev)
        self.indices = [torch.zeros(B * H * W, dtype=torch.int64, device=dev) for _ in fs.embed_dim]
        # a14: VQ per scale, written fine->coarse (msvqgan.py:392-393)
        quant = P.buf(B, H * W, Ct)
        start = 0
        for i, e in enumerate(fs.embed_dim):
            coff = sum(fs.embed_dim[i + 1:])
            P.vq(self.z, self._vec(fs.ms_quantize[i].embedding.weight), quant, self.indices[i], B=B, C_total=Ct, HW=H * W,
                 c_start=start, e_dim=e, scale_factor=sf[i], out_C=Ct, out_coff=coff, tag=f"vq{i}")
            start += e
        zc = fs.post_quant_conv.weight.shape[0]
        pq = P.buf(B, H * W, zc)
        P.conv(Src.nhwc(quant, H, W), self._packed(lambda: PK.copy(fs.post_quant_conv.weight.detach().view(zc, Ct))), pq, B=B,
               Hin=H, Win=W, Hout=H, Wout=W, Cout=zc, bias=self._vec(fs.post_quant_conv.bias), tag="post_quant_conv")
        f = 2 ** (dec.num_resolutions - 1)
        oc = dec.conv_out.weight.shape[0]
        self.image = torch.zeros(B, oc, H * f, W * f, dtype=torch.float32, device=dev)
        # the sampling script's output formatting (sample_diffusion.py:103-121) rides conv_out's epilogue: uint8 NHWC
        self.image_u8 = torch.zeros(B, H * f, W * f, oc, dtype=torch.uint8, device=dev) if oc <= 4 else None
        self._decoder_body(dec, pq, H, W, self.image, True, "dec", out_u8=self.image_u8, u8_mode=u8_mode)
        self._finish_plan()


class EncodePlan(_VQPlan):
    """One program for VQModelInterface.encode (msvqgan.py:326-374) + the scale-factor multiply of
    get_first_stage_encoding (frido.py:646-662)."""

    def __init__(self, fs: VQModelInterface, B, H, W, sf):
        self.H, self.W = H, W
        enc = fs.encoder
        n_gn = sum(1 for m in list(enc.modules()) + list(fs.shared_decoder.modules()) if isinstance(m, nn.GroupNorm))
        self._init_plan(fs, B, "encode", n_gn)
        dev, P = self.dev, self.prog
        S = enc.multiscale
        f = 2 ** (enc.num_resolutions - 1)
        if H % f or W % f:
            raise L.FridoError(f"encode: image size {H}x{W} must be a multiple of {f}")
        cin = enc.conv_in.weight.shape[1]
        self.x = torch.zeros(B, cin, H, W, dtype=torch.float32, device=dev)
        # ---- bottom-up (model.py:512-528) --------------------------------
        c = enc.ch
        hh, ww = H, W
        h = P.buf(B, hh * ww, c)
        P.conv(Src.nchw(self.x, hh, ww, 0, cin), self._conv_w(enc.conv_in), h, B=B, Hin=hh, Win=ww, Hout=hh, Wout=ww, Cout=c,
               ksize=3, pad=1, bias=self._vec(enc.conv_in.bias), tag="enc.conv_in")
        taps = []  # (tensor, C, h, w) of the last block of each level
        for lvl in range(enc.num_resolutions):
            down = enc.down[lvl]
            keep = lvl >= enc.num_resolutions - S  # this level's output also feeds a multi-scale head
            for j, blk in enumerate(down.block):
                h = self._res(blk, h, hh, ww, tag="enc.res")
                if len(down.attn) > 0:
                    h = self._attn(down.attn[j], h, hh, ww, tag="enc.attn")
            c = down.block[-1].out_channels
            if keep:
                taps.append((h, c, hh, ww))
            if lvl != enc.num_resolutions - 1:
                o = P.buf(B, (hh // 2) * (ww // 2), c)
                P.conv(Src.nhwc(h, hh, ww), self._conv_w(down.downsample.conv), o, B=B, Hin=hh, Win=ww, Hout=hh // 2, Wout=ww // 2,
                       Cout=c, ksize=3, stride=2, pad=0, bias=self._vec(down.downsample.conv.bias), tag="enc.down")
                if not keep:
                    P.release(h)
                h, hh, ww = o, hh // 2, ww // 2
        # ---- per-scale heads (model.py:530-544), taps[i] <-> mid_ms[i]: finest first ----
        dz = fs.ms_quant_conv[0].weight.shape[1] // fs.edconfig["z_channels"][0]
        zc0 = fs.edconfig["z_channels"][0]
        heads = [None] * S  # indexed coarse-first (ii of msvqgan.py:332)
        for i, (t, c, th, tw) in enumerate(taps):
            ii = S - 1 - i
            zc = enc.conv_out_ms[i].weight.shape[0]
            ctot = ii * zc0 + zc  # cat((*prev_h[:ii], h_ms[ii]), dim=1): the head writes the last zc channels
            cat = P.buf(B, th * tw, ctot)
            # the tap's other reader (the downsample conv) is already enqueued: block_1 may consume and release it
            m = enc.mid_ms[i]
            y = self._res(m.block_1, t, th, tw, tag="enc.mid.res")
            y = self._attn(m.attn_1, y, th, tw, tag="enc.mid.attn")
            y = self._res(m.block_2, y, th, tw, tag="enc.mid.res")
            n = self._gn(y, c, th, tw, enc.norm_out_ms[i], 1)
            P.release(y)
            P.conv(Src.nhwc(n, th, tw), self._conv_w(enc.conv_out_ms[i]), cat, B=B, Hin=th, Win=tw, Hout=th, Wout=tw, Cout=zc,
                   ksize=3, pad=1, bias=self._vec(enc.conv_out_ms[i].bias), o_sb=th * tw * ctot, o_sp=ctot, o_sn=1,
                   out_off=ii * zc0, tag="enc.conv_out_ms")
            P.release(n)
            heads[ii] = (cat, ctot, th, tw)
        # ---- top-down, coarse -> fine (msvqgan.py:332-350) ---------------
        Ct = sum(fs.embed_dim)
        fh, fw = heads[S - 1][2], heads[S - 1][3]
        self.latent = torch.zeros(B, Ct, fh, fw, dtype=torch.float32, device=dev)
        self.indices = []
        prev = []  # prev_h of msvqgan.py:330: (tensor, pixel stride, channel offset) NHWC maps at the previous resolution
        e0 = fs.embed_dim[0]
        coff = 0
        for ii in range(S):
            cat, ctot, th, tw = heads[ii]
            e = fs.embed_dim[ii]
            if ii > 0:
                ct, pq = fs.upsample[ii - 1], fs.shared_post_quant_conv[ii - 1]
                wt, bt = self._vec(ct.weight), self._vec(ct.bias)
                wp, bp = self._packed(lambda pq=pq: PK.copy(pq.weight.detach().view(zc0, e0))), self._vec(pq.bias)
                for j in range(ii):
                    u = P.buf(B, th * tw, e0)
                    src, ld, off = prev[j]
                    P.conv_transpose2d(src, wt, bt, u, B=B, H=th // 2, W=tw // 2, Cin=e0, Cout=e0, x_ld=ld, x_off=off,
                                       tag="enc.upsample")
                    # shared_post_quant_conv writes straight into its channel slice of the concat (msvqgan.py:337-340)
                    P.conv(Src.nhwc(u, th, tw), wp, cat, B=B, Hin=th, Win=tw, Hout=th, Wout=tw, Cout=zc0, bias=bp,
                           o_sb=th * tw * ctot, o_sp=ctot, o_sn=1, out_off=j * zc0, tag="enc.shared_post_quant_conv")
                    P.release(u)
                    prev[j] = (cat, ctot, j * zc0)
                q_in = P.buf(B, th * tw, e0)
                self._decoder_body(fs.shared_decoder[ii - 1], cat, th, tw, q_in, False, "enc.shared_decoder")
                src_c = e0
            else:
                q_in, src_c = cat, ctot
            hq = P.buf(B, th * tw, e)
            mq = fs.ms_quant_conv[ii]
            P.conv(Src.nhwc(q_in, th, tw, c_total=src_c), self._packed(lambda mq=mq, e=e, k=src_c: PK.copy(mq.weight.detach().view(e, k))),
                   hq, B=B, Hin=th, Win=tw, Hout=th, Wout=tw, Cout=e, bias=self._vec(mq.bias), tag="enc.ms_quant_conv")
            sh = (fh // th).bit_length() - 1
            P.assemble_latent(hq, self.latent, B=B, H=fh, W=fw, e=e, sh=sh, scale=sf[ii], C_total=Ct, c_off=coff, tag="enc.assemble")
            coff += e
            idx = torch.zeros(B * th * tw, dtype=torch.int64, device=dev)
            self.indices.append(idx)
            q = P.buf(B, th * tw, e)
            P.vq(hq, self._vec(fs.ms_quantize[ii].embedding.weight), q, idx, B=B, C_total=e, HW=th * tw, c_start=0, e_dim=e,
                 scale_factor=1.0, out_C=e, out_coff=0, z_nhwc=1, tag=f"enc.vq{ii}")
            prev.append((q, e, 0))
        self._finish_plan()
