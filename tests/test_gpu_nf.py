"""GPU parity of the normalise-on-load tcgen05 engine (csrc/conv_nf.cu, FridoConvParams.nrm_ab):
GroupNorm -> [SPADE] -> [SiLU] -> conv3x3 / conv1x1 [+ 1x1 side input, bias, timestep row, residual, channel sums] as one
launch, against fp64 PyTorch on the CPU restating pyunet.py:209-240 / spade_norm.py:44-60 / attention.py:296-298, and
against the two-launch path (frido_norm_act + conv2d_tc) it replaces.

Tolerance: BF16x3 products are good to ~2^-16 relative per operand pair and the fp32 tensor-core accumulation truncates;
for O(1) outputs |err| < 1e-4 + 1.5e-8 * K * |out|max, the bound tests/test_gpu_tc.py uses for engine 3."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def _pack(w):
    return w.permute(0, 2, 3, 1).contiguous().view(w.shape[0], -1)


@pytest.fixture(params=["auto", "force", "off", "pair"])
def sk_mode(request):
    """Schedules of the engine: stream-K by cost model / whenever legal / never, and 'pair' = the CTA-pair (cta_group::2)
    kernels whenever legal (FRIDO_NF_PAIR=2, FRIDO_TC_PAIR=2) with stream-K off."""
    import os
    old = {k: os.environ.get(k) for k in ("FRIDO_SK", "FRIDO_NF_PAIR", "FRIDO_TC_PAIR")}
    os.environ["FRIDO_SK"] = {"auto": "1", "force": "2", "off": "0", "pair": "0"}[request.param]
    if request.param == "pair":
        os.environ["FRIDO_NF_PAIR"] = "2"
        os.environ["FRIDO_TC_PAIR"] = "2"
    yield request.param
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


CASES = [
    # B, C0, C1, Cout, H, W, k, spade, silu, side
    (2, 64, 0, 64, 16, 16, 3, False, True, 0),      # one tile per image, halo = TMA zero fill on all four sides
    (2, 64, 0, 64, 16, 16, 3, True, True, 0),       # + SPADE maps
    (4, 64, 32, 128, 8, 8, 3, True, True, 96),      # skip concat (two sources), 2 images per tile, side input over both sources
    (2, 192, 0, 192, 32, 32, 3, False, True, 0),    # 8 tiles per image: interior tiles whose halo is real neighbour data
    (2, 192, 0, 192, 32, 32, 3, True, True, 64),    # the 64^2-level ResBlock conv2 + skip shape, scaled down
    (3, 960, 960, 960, 8, 8, 3, True, True, 0),     # long K loop (60 chunks x 9 taps), half-empty second tile (TB=2, B=3)
    (3, 96, 0, 256, 9, 7, 3, True, True, 0),        # ragged spatial size: partial tiles, masked rows
    (2, 384, 0, 384, 32, 32, 1, True, False, 0),    # SpatialTransformer norm -> proj_in (1x1, no SiLU)
    (16, 576, 0, 576, 16, 16, 1, False, False, 0),
    (3, 64, 0, 192, 40, 24, 3, True, True, 0),      # non-power-of-two image larger than a tile in both directions
]


def _reference(x0, x1, gn_w, gn_b, gb, silu, w, bias, k, eps, side_w, rowvec, res):
    """fp64 restatement: GroupNorm(32) over the concatenated sources, SPADE `normalized*(1+gamma)+beta`, SiLU, conv
    (zero padding of the activated tensor), + the 1x1 side conv over the RAW sources, + bias / per-image row / residual."""
    xin = torch.cat([x0, x1], 1) if x1 is not None else x0
    y = F.group_norm(xin.double(), 32, gn_w.double(), gn_b.double(), eps)
    if gb is not None:
        C = xin.shape[1]
        y = y * (1 + gb[:, :C].double()) + gb[:, C:].double()
    if silu:
        y = F.silu(y)
    out = F.conv2d(y, w.double(), bias.double(), padding=k // 2)
    if side_w is not None:
        out = out + F.conv2d(xin.double(), side_w.double()[:, :, None, None])
    if rowvec is not None:
        out = out + rowvec.double()[:, :, None, None]
    if res is not None:
        out = out + res.double()
    return out


@pytest.mark.parametrize("case", CASES)
def test_conv_nf_matches_fp64_and_the_unfused_path(dev, case, sk_mode):
    from frido_b200 import _lib as L
    from frido_b200.program import Program, Src
    B, C0, C1, Cout, H, W, k, spade, silu, side = case
    g = torch.Generator().manual_seed(abs(hash(case)) % 2**31)
    Cin = C0 + C1
    # activations with a per-channel offset and scale, so that the normalisation actually matters
    x0 = torch.randn(B, C0, H, W, generator=g) * (0.5 + torch.rand(1, C0, 1, 1, generator=g)) + torch.randn(1, C0, 1, 1, generator=g)
    x1 = (torch.randn(B, C1, H, W, generator=g) * 1.5 + 0.3) if C1 else None
    gn_w, gn_b = 1 + 0.2 * torch.randn(Cin, generator=g), 0.2 * torch.randn(Cin, generator=g)
    gb = 0.3 * torch.randn(B, 2 * Cin, H, W, generator=g) if spade else None
    w = torch.randn(Cout, Cin, k, k, generator=g) / np.sqrt(Cin * k * k)
    bias = torch.randn(Cout, generator=g)
    side_w = torch.randn(Cout, side, generator=g) / np.sqrt(side) if side else None
    rowvec = torch.randn(B, Cout, generator=g)
    res = torch.randn(B, Cout, H, W, generator=g) if not side else None
    eps = 1e-5
    xin = torch.cat([x0, x1], 1) if C1 else x0
    assert side in (0, Cin) or side <= C0
    side_src = xin[:, :side] if side else None
    ref = _reference(x0, x1, gn_w, gn_b, gb, silu, w, bias, k, eps, None, None, None)
    if side:
        ref_full = ref + F.conv2d(side_src.double(), side_w.double()[:, :, None, None]) + rowvec.double()[:, :, None, None]
    else:
        ref_full = ref + rowvec.double()[:, :, None, None] + res.double()

    P = Program(dev, "nf", engine="bf16x3")
    d0 = _nhwc(x0).to(dev).view(B, H * W, C0)
    d1 = _nhwc(x1).to(dev).view(B, H * W, C1) if C1 else None
    a0 = Src.nhwc(d0, H, W)
    a1 = Src.nhwc(d1, H, W) if C1 else None
    out_chk = torch.zeros(B, H * W, Cout, device=dev)
    assert P.nf_eligible(a0, a1, out_chk, B=B, H=H, W=W, Cout=Cout, ksize=k)
    gw, gbias = gn_w.to(dev), gn_b.to(dev)
    gbd = _nhwc(gb).to(dev).view(B, H * W, 2 * Cin) if spade else None
    # statistics: per-channel sums, the format conv epilogues produce (chan_sums)
    cs0 = torch.zeros(B, C0, 2, dtype=torch.float64, device=dev)
    P.zero(cs0)  # the program is replayed below: every accumulator is cleared by the program itself
    P.chan_stats(d0, C0, cs0, B=B, HW=H * W)
    cs1 = None
    if C1:
        cs1 = torch.zeros(B, C1, 2, dtype=torch.float64, device=dev)
        P.zero(cs1)
        P.chan_stats(d1, C1, cs1, B=B, HW=H * W)
    ab = torch.zeros(B, Cin, 2, device=dev)
    P.gn_finalize(ab, gw, gbias, B=B, HW=H * W, c0=C0, c1=C1, eps=eps, csum0=cs0, csum1=cs1)
    wd = _pack(w).to(dev)
    if side:
        wd = torch.cat([wd, side_w.to(dev)], 1).contiguous()
        if side <= C0:
            side_arg = (Src.nhwc(d0, H, W, C_=side), None)
        else:
            side_arg = (Src.nhwc(d0, H, W), Src.nhwc(d1, H, W))
    else:
        side_arg = None
    out1 = torch.zeros(B, H * W, Cout, device=dev)
    out2 = torch.zeros(B, H * W, Cout, device=dev)
    csum = torch.zeros(B, Cout, 2, dtype=torch.float64, device=dev)
    P.zero(csum)
    kw = dict(B=B, Hin=H, Win=W, Hout=H, Wout=W, Cout=Cout, ksize=k, pad=k // 2, a1=a1, bias=bias.to(dev), engine=3,
              nrm=(ab, gbd, int(silu)))
    if not side:
        P.conv(a0, wd, out1, **kw)
    ok = P.conv(a0, wd, out2, rowvec=rowvec.to(dev), rowvec_sb=Cout, side=side_arg, csum=csum,
                res=None if side else _nhwc(res).to(dev).view(B, H * W, Cout), **kw)
    # the two-launch path this replaces
    t = torch.zeros(B, H * W, Cin, device=dev)
    P.norm_act(d0, C0, None, gw, gbias, t, B=B, HW=H * W, eps=eps, a1=d1, c1=C1, gb=gbd, silu=int(silu), csum0=cs0, csum1=cs1)
    out3 = torch.zeros(B, H * W, Cout, device=dev)
    wd3 = _pack(w).to(dev)
    P.conv(Src.nhwc(t, H, W), wd3, out3, B=B, Hin=H, Win=W, Hout=H, Wout=W, Cout=Cout, ksize=k, pad=k // 2, bias=bias.to(dev), engine=3)
    # the split path: norm_act writes the engine's operand form, the conv's halo-resident path feeds it as is
    out4 = None
    if k == 3:
        t4 = torch.zeros(B, H * W, Cin, device=dev)
        P.norm_act(d0, C0, None, gw, gbias, t4, B=B, HW=H * W, eps=eps, a1=d1, c1=C1, gb=gbd, silu=int(silu), csum0=cs0, csum1=cs1,
                   out_split=1)
        out4 = torch.zeros(B, H * W, Cout, device=dev)
        P.conv(Src.nhwc(t4, H, W), wd, out4, B=B, Hin=H, Win=W, Hout=H, Wout=W, Cout=Cout, ksize=k, pad=k // 2, bias=bias.to(dev),
               engine=3, presplit=True, side=side_arg, rowvec=rowvec.to(dev), rowvec_sb=Cout,
               res=None if side else _nhwc(res).to(dev).view(B, H * W, Cout))
    P.prepare_weights()
    P.run()
    torch.cuda.synchronize(dev)

    def nchw(o):
        return o.view(B, H, W, Cout).permute(0, 3, 1, 2).double().cpu()

    K = Cin * k * k + side
    tol = 1e-4 + 1.5e-8 * K * ref_full.abs().max().item()
    e3 = (nchw(out3) - ref).abs().max().item()
    if not side:
        e1 = (nchw(out1) - ref).abs().max().item()
        d13 = (nchw(out1) - nchw(out3)).abs().max().item()
        assert e1 < tol and d13 < tol, (case, e1, e3, d13, tol)
    e2 = (nchw(out2) - ref_full).abs().max().item()
    e4 = (nchw(out4) - ref_full).abs().max().item() if out4 is not None else 0.0
    print(f"case {case}: fused err {e2:.3e}, split err {e4:.3e}, two-launch err {e3:.3e}, tol {tol:.3e}, |ref| {ref_full.abs().max():.2f}")
    assert e2 < tol and e3 < tol and e4 < tol, (case, e2, e3, e4, tol)
    if ok:  # channel sums of the stored outputs (what the NEXT GroupNorm consumes)
        o = nchw(out2)
        want = torch.stack([o.sum((2, 3)), (o * o).sum((2, 3))], -1)
        got = csum.cpu()
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-3), (case, (got - want).abs().max())
    if sk_mode != "off":  # replay: stream-K arrival counters back at zero, fixed summation order -> bit-identical
        first = out2.clone()
        P.run()
        torch.cuda.synchronize(dev)
        assert torch.equal(out2, first)


def test_gn_finalize_from_group_sums_and_channel_sums_agree(dev):
    from frido_b200.program import Program
    B, C, HW = 3, 192, 24 * 24
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(B, HW, C, generator=g) * 2 + 1).to(dev)
    gw, gb_ = (1 + 0.1 * torch.randn(C, generator=g)).to(dev), (0.1 * torch.randn(C, generator=g)).to(dev)
    P = Program(dev, "fin")
    sums = torch.zeros(B, 32, 2, dtype=torch.float64, device=dev)
    cs = torch.zeros(B, C, 2, dtype=torch.float64, device=dev)
    P.gn_stats(x, C, sums, B=B, HW=HW)
    P.chan_stats(x, C, cs, B=B, HW=HW)
    ab1, ab2 = torch.zeros(B, C, 2, device=dev), torch.zeros(B, C, 2, device=dev)
    P.gn_finalize(ab1, gw, gb_, B=B, HW=HW, c0=C, eps=1e-6, sums=sums)
    P.gn_finalize(ab2, gw, gb_, B=B, HW=HW, c0=C, eps=1e-6, csum0=cs)
    P.run()
    torch.cuda.synchronize(dev)
    xd = x.double().cpu().view(B, HW, 32, C // 32)
    mean = xd.mean((1, 3))
    var = xd.var((1, 3), unbiased=False)
    rstd = 1.0 / torch.sqrt(var + 1e-6)
    a = (rstd[:, :, None] * gw.double().cpu().view(1, 32, -1)).reshape(B, C)
    b = gb_.double().cpu()[None] - (mean[:, :, None] * rstd[:, :, None] * gw.double().cpu().view(1, 32, -1)).reshape(B, C)
    for ab in (ab1, ab2):
        assert (ab[..., 0].double().cpu() - a).abs().max() < 1e-5
        assert (ab[..., 1].double().cpu() - b).abs().max() < 1e-5
