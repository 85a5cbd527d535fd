"""Parameter containers that mirror the reference's module tree so that the
reference's checkpoints (`state_dict` keys and tensor layouts) load unchanged
(SURVEY.md §8b).  These modules hold weights only: none of them has a torch
forward — the arithmetic lives in the CUDA library and is scheduled by
unet.py (UNetPlan) / first_stage.py (DecodePlan, EncodePlan) / cond.py.

Key layout follows frido/modules/diffusionmodules/pyunet.py:477-835,
frido/modules/attention.py:152-287, frido/modules/diffusionmodules/spade_norm.py:26-42
and taming/modules/diffusionmodules/model.py:78-192,548-616 (names only).
"""
import torch
from torch import nn


class _NoForward(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError(f"{type(self).__name__} is a weight container; compute runs in libfrido_b200.so")


def _zero(m):
    for p in m.parameters():
        p.detach().zero_()
    return m


class Seq(nn.Sequential):
    """nn.Sequential used purely for its integer child names."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("weight container only")


class Marker(_NoForward):
    """Parameter-free placeholder (SiLU / Dropout / ReLU / Identity slots keep indices aligned)."""


def gn(ch, eps):
    return nn.GroupNorm(32, ch, eps=eps, affine=True)


class SPADE(_NoForward):
    def __init__(self, norm, norm_nc, cond_nc, nhidden=128):
        super().__init__()
        self.param_free_norm = norm
        self.mlp_shared = Seq(nn.Conv2d(cond_nc, nhidden, 3, padding=1), Marker())
        self.mlp_gamma = nn.Conv2d(nhidden, norm_nc, 3, padding=1)
        self.mlp_beta = nn.Conv2d(nhidden, norm_nc, 3, padding=1)


def make_norm(ch, cond_ch, eps, spade):
    return SPADE(gn(ch, eps), ch, cond_ch) if spade else gn(ch, eps)


class ResBlock(_NoForward):
    def __init__(self, ch, cond_ch, emb_ch, out_ch, spade):
        super().__init__()
        self.channels, self.out_channels = ch, out_ch
        self.in_layers = Seq(make_norm(ch, cond_ch, 1e-5, spade), Marker(), nn.Conv2d(ch, out_ch, 3, padding=1))
        self.emb_layers = Seq(Marker(), nn.Linear(emb_ch, out_ch))
        self.out_layers = Seq(make_norm(out_ch, cond_ch, 1e-5, spade), Marker(), Marker(),
                              _zero(nn.Conv2d(out_ch, out_ch, 3, padding=1)))
        self.skip_connection = Marker() if out_ch == ch else nn.Conv2d(ch, out_ch, 1)


class Downsample(_NoForward):
    def __init__(self, ch, out_ch):
        super().__init__()
        self.op = nn.Conv2d(ch, out_ch, 3, stride=2, padding=1)


class Upsample(_NoForward):
    def __init__(self, ch, out_ch):
        super().__init__()
        self.conv = nn.Conv2d(ch, out_ch, 3, padding=1)


class CrossAttention(_NoForward):
    def __init__(self, query_dim, context_dim, inner):
        super().__init__()
        context_dim = context_dim or query_dim
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(context_dim, inner, bias=False)
        self.to_v = nn.Linear(context_dim, inner, bias=False)
        self.to_out = Seq(nn.Linear(inner, query_dim), Marker())


class GEGLU(_NoForward):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)


class FeedForward(_NoForward):
    def __init__(self, dim, mult=4):
        super().__init__()
        self.net = Seq(GEGLU(dim, dim * mult), Marker(), nn.Linear(dim * mult, dim))


class BasicTransformerBlock(_NoForward):
    def __init__(self, dim, context_dim):
        super().__init__()
        self.attn1 = CrossAttention(dim, None, dim)
        self.ff = FeedForward(dim)
        self.attn2 = CrossAttention(dim, context_dim, dim)
        self.norm1, self.norm2, self.norm3 = nn.LayerNorm(dim), nn.LayerNorm(dim), nn.LayerNorm(dim)


class SpatialTransformer(_NoForward):
    def __init__(self, ch, cond_ch, depth, context_dim, spade):
        super().__init__()
        self.in_channels = ch
        self.norm = make_norm(ch, cond_ch, 1e-6, spade)
        self.proj_in = nn.Conv2d(ch, ch, 1)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(ch, context_dim) for _ in range(depth)])
        self.proj_out = _zero(nn.Conv2d(ch, ch, 1))


# ---- taming decoder (first stage) -------------------------------------------


class TResnetBlock(_NoForward):
    def __init__(self, cin, cout):
        super().__init__()
        self.in_channels, self.out_channels = cin, cout
        self.norm1 = gn(cin, 1e-6)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = gn(cout, 1e-6)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        if cin != cout:
            self.nin_shortcut = nn.Conv2d(cin, cout, 1)


class TAttnBlock(_NoForward):
    def __init__(self, ch):
        super().__init__()
        self.norm = gn(ch, 1e-6)
        self.q, self.k, self.v = nn.Conv2d(ch, ch, 1), nn.Conv2d(ch, ch, 1), nn.Conv2d(ch, ch, 1)
        self.proj_out = nn.Conv2d(ch, ch, 1)


class TUpsample(_NoForward):
    def __init__(self, ch):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, padding=1)


class TDecoder(_NoForward):
    """taming Decoder (model.py:548-616): weights only."""

    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, resolution, z_channels,
                 **ignored):
        super().__init__()
        self.ch, self.num_resolutions, self.num_res_blocks = ch, len(ch_mult), num_res_blocks
        block_in = ch * ch_mult[-1]
        curr_res = resolution // 2 ** (self.num_resolutions - 1)
        self.conv_in = nn.Conv2d(z_channels, block_in, 3, padding=1)
        self.mid = nn.Module()
        self.mid.block_1 = TResnetBlock(block_in, block_in)
        self.mid.attn_1 = TAttnBlock(block_in)
        self.mid.block_2 = TResnetBlock(block_in, block_in)
        ups = []
        for i_level in reversed(range(self.num_resolutions)):
            block, attn = nn.ModuleList(), nn.ModuleList()
            block_out = ch * ch_mult[i_level]
            for _ in range(num_res_blocks + 1):
                block.append(TResnetBlock(block_in, block_out))
                block_in = block_out
                if curr_res in attn_resolutions:
                    attn.append(TAttnBlock(block_in))
            up = nn.Module()
            up.block, up.attn = block, attn
            if i_level != 0:
                up.upsample = TUpsample(block_in)
                curr_res *= 2
            ups.insert(0, up)
        self.up = nn.ModuleList(ups)
        self.norm_out = gn(block_in, 1e-6)
        self.conv_out = nn.Conv2d(block_in, out_ch, 3, padding=1)


class TDownsample(_NoForward):
    """taming Downsample (model.py:57-78): 3x3 stride-2 conv after a right/bottom zero pad."""

    def __init__(self, ch):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, stride=2, padding=0)


class MSEncoder(_NoForward):
    """taming MSEncoder (model.py:435-510): weights only.  One mid/norm/conv_out head per scale, finest first."""

    def __init__(self, *, ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, in_channels, resolution, z_channels,
                 double_z=True, multiscale=3, **ignored):
        super().__init__()
        self.ch, self.num_resolutions, self.num_res_blocks = ch, len(ch_mult), num_res_blocks
        self.resolution, self.in_channels, self.multiscale = resolution, in_channels, multiscale
        self.conv_in = nn.Conv2d(in_channels, ch, 3, padding=1)
        curr_res = resolution
        in_ch_mult = (1,) + tuple(ch_mult)
        self.down = nn.ModuleList()
        for i_level in range(self.num_resolutions):
            block, attn = nn.ModuleList(), nn.ModuleList()
            block_in, block_out = ch * in_ch_mult[i_level], ch * ch_mult[i_level]
            for _ in range(num_res_blocks):
                block.append(TResnetBlock(block_in, block_out))
                block_in = block_out
                if curr_res in attn_resolutions:
                    attn.append(TAttnBlock(block_in))
            down = nn.Module()
            down.block, down.attn = block, attn
            if i_level != self.num_resolutions - 1:
                down.downsample = TDownsample(block_in)
                curr_res = curr_res // 2
            self.down.append(down)
        in_ch_mult = in_ch_mult[-multiscale:]
        assert len(z_channels) == multiscale, "Error using multiscale encoder, but the z gets wrong dim."
        self.mid_ms, self.norm_out_ms, self.conv_out_ms = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        for i in range(multiscale):
            block_in = ch * in_ch_mult[i]
            mid = nn.Module()
            mid.block_1 = TResnetBlock(block_in, block_in)
            mid.attn_1 = TAttnBlock(block_in)
            mid.block_2 = TResnetBlock(block_in, block_in)
            self.mid_ms.append(mid)
            self.norm_out_ms.append(gn(block_in, 1e-6))
            self.conv_out_ms.append(nn.Conv2d(block_in, 2 * z_channels[i] if double_z else z_channels[i], 3, padding=1))


class VectorQuantizer(_NoForward):
    def __init__(self, n_e, e_dim, init_normal=False):
        super().__init__()
        self.n_e, self.e_dim = n_e, e_dim
        self.embedding = nn.Embedding(n_e, e_dim)
        if init_normal:
            self.embedding.weight.data.normal_(0.0, 1.0)  # quantize.py:225
        else:
            self.embedding.weight.data.uniform_(-1.0 / n_e, 1.0 / n_e)  # quantize.py:223
