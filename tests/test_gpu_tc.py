"""GPU parity for the tcgen05 TF32 engine (FridoConvParams.engine = 1) against fp32 PyTorch on the
CPU.  Tolerance: TF32 keeps 10 mantissa bits per operand -> relative error ~ 2^-11 per product,
accumulated in fp32; we bound |err| by 4e-3 * sqrt-ish scale of the output (outputs here are O(1))."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
TOL = 4e-3   # single-pass TF32 on un-rounded operands (hardware truncation)


def tol3(K, absmax):
    """3xTF32: products are fp32-faithful (~2^-21); what remains is the tensor core's round-toward-zero
    fp32 accumulation, a bias that grows linearly with the number of K-steps (measured 5e-9 * K * |out|)."""
    return 1e-5 + 1.5e-8 * K * absmax



@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def _pack(w):
    return w.permute(0, 2, 3, 1).contiguous().view(w.shape[0], -1)


def _run(P, dev):
    P.prepare_weights()
    P.run()
    torch.cuda.synchronize(dev)


CASES = [
    # B, C0, C1, Cout, H, W, k
    (1, 64, 0, 64, 1, 256, 1),      # plain linear, 2 m-tiles, 2 k-steps
    (2, 64, 0, 64, 16, 16, 3),      # conv3x3, halo via TMA OOB zero fill
    (4, 64, 32, 128, 8, 8, 3),      # channel concat + 2 images per tile
    (2, 192, 0, 192, 32, 32, 3),    # BN=192
    (1, 960, 960, 960, 8, 8, 3),    # long K loop (540 k-steps), half-empty tile (TB=2, B=1)
    (3, 96, 0, 256, 9, 7, 3),       # ragged spatial size: masked rows
    (16, 128, 0, 512, 4, 4, 1),     # 4x4 images, 8 per tile
    (2, 384, 0, 7680 // 4, 16, 16, 1),
]


@pytest.fixture(params=["auto", "force", "off"])
def sk_mode(request):
    """Stream-K schedule of conv2d_tc: cost model / whenever legal / never (FRIDO_SK, read at every launch)."""
    import os
    old = os.environ.get("FRIDO_SK")
    os.environ["FRIDO_SK"] = {"auto": "1", "force": "2", "off": "0"}[request.param]
    yield request.param
    if old is None:
        os.environ.pop("FRIDO_SK", None)
    else:
        os.environ["FRIDO_SK"] = old


@pytest.mark.parametrize("eng", [1, 2, 3])
@pytest.mark.parametrize("case", CASES)
def test_conv_tc_matches_torch(dev, case, eng, sk_mode):
    from frido_b200 import _lib as L
    from frido_b200.program import Program, Src
    B, C0, C1, Cout, H, W, k = case
    g = torch.Generator().manual_seed(abs(hash(case)) % 2**31)
    Cin = C0 + C1
    x0 = torch.randn(B, C0, H, W, generator=g)
    x1 = torch.randn(B, C1, H, W, generator=g) if C1 else None
    w = torch.randn(Cout, Cin, k, k, generator=g) / np.sqrt(Cin * k * k)
    bias = torch.randn(Cout, generator=g)
    rowvec = torch.randn(B, Cout, generator=g)
    res = torch.randn(B, Cout, H, W, generator=g)
    xin = torch.cat([x0, x1], 1) if C1 else x0
    ref = F.conv2d(xin, w, bias, padding=k // 2)
    ref2 = F.silu(ref + rowvec[:, :, None, None] + res)
    P = Program(dev, "tc")
    a0 = Src.nhwc(_nhwc(x0).to(dev), H, W)
    a1 = Src.nhwc(_nhwc(x1).to(dev), H, W) if C1 else None
    wd = _pack(w).to(dev)
    out = torch.zeros(B, H * W, Cout, device=dev)
    out2 = torch.zeros(B, H * W, Cout, device=dev)
    P.conv(a0, wd, out, B=B, Hin=H, Win=W, Hout=H, Wout=W, Cout=Cout, ksize=k, pad=k // 2, a1=a1, bias=bias.to(dev), engine=eng)
    P.conv(a0, wd, out2, B=B, Hin=H, Win=W, Hout=H, Wout=W, Cout=Cout, ksize=k, pad=k // 2, a1=a1, bias=bias.to(dev),
           rowvec=rowvec.to(dev), rowvec_sb=Cout, res=_nhwc(res).to(dev).view(B, H * W, Cout), act=L.ACT_SILU, engine=eng)
    _run(P, dev)
    got = out.view(B, H, W, Cout).permute(0, 3, 1, 2).cpu()
    got2 = out2.view(B, H, W, Cout).permute(0, 3, 1, 2).cpu()
    if sk_mode != "off":  # replay: the stream-K arrival counters must be back at zero, and the fixed summation order
        P.run()           # of the partial sums makes the result bit-reproducible
        torch.cuda.synchronize(dev)
        assert torch.equal(out2.view(B, H, W, Cout).permute(0, 3, 1, 2).cpu(), got2)
    e1, e2 = (got - ref).abs().max().item(), (got2 - ref2).abs().max().item()
    print(f"case {case} engine {eng}: err {e1:.3e} {e2:.3e} (ref absmax {ref.abs().max():.2f})")
    tol = TOL if eng == 1 else tol3(Cin * k * k, ref.abs().max().item()) + (1e-4 if eng == 3 else 0.0)  # bf16x3 products ~2^-16
    assert e1 < tol and e2 < tol, (case, e1, e2)


@pytest.mark.parametrize("eng", [1, 2, 3])
def test_tc_geglu_and_attention_shapes(dev, eng):
    from frido_b200 import _lib as L
    from frido_b200.program import Program, Src
    g = torch.Generator().manual_seed(3)
    P = Program(dev, "tc2")
    M_, K, inner = 256, 64, 128
    xt = torch.randn(M_, K, generator=g)
    wp = torch.randn(2 * inner, K, generator=g) / 8
    bp = torch.randn(2 * inner, generator=g)
    hh = F.linear(xt, wp, bp)
    refg = hh[:, :inner] * F.gelu(hh[:, inner:])
    wi = torch.stack([wp[:inner], wp[inner:]], 1).reshape(2 * inner, K).contiguous()
    bi = torch.stack([bp[:inner], bp[inner:]], 1).reshape(-1).contiguous()
    outg = torch.zeros(M_, inner, device=dev)
    P.linear(xt.to(dev), wi.to(dev), outg, M=M_, K=K, N=2 * inner, bias=bi.to(dev), act=L.ACT_GEGLU, engine=eng)
    # attention: QK^T with per-image weights inside a fused q|k tensor; V^T transposed store; P.V
    B, N, C = 3, 256, 64
    qk = torch.randn(B, N, 2 * C, generator=g)
    S = torch.einsum("bid,bjd->bij", qk[..., :C], qk[..., C:])
    qkd = qk.to(dev)
    sc = torch.zeros(B, N, N, device=dev)
    P.conv(Src(qkd, C, N * 2 * C, 0, 2 * C, 1), qkd, sc, B=B, Hin=1, Win=N, Hout=1, Wout=N, Cout=N, w_sb=N * 2 * C,
           w_ld=2 * C, w_off=C, engine=eng)
    x = torch.randn(B, N, C, generator=g)
    wv = torch.randn(C, C, generator=g) / 8
    vT_ref = F.linear(x, wv).transpose(1, 2).contiguous()
    vT = torch.zeros(B, C, N, device=dev)
    P.conv(Src(x.to(dev), C, N * C, 0, C, 1), wv.to(dev), vT, B=B, Hin=1, Win=N, Hout=1, Wout=N, Cout=C, o_sb=C * N,
           o_sp=1, o_sn=N, engine=eng)
    prob = torch.softmax(torch.randn(B, N, N, generator=g), -1)
    vTm = torch.randn(B, C, N, generator=g)
    o_ref = torch.einsum("bij,bdj->bid", prob, vTm)
    o = torch.zeros(B, N, C, device=dev)
    P.conv(Src(prob.to(dev), N, N * N, 0, N, 1), vTm.to(dev), o, B=B, Hin=1, Win=N, Hout=1, Wout=N, Cout=C, w_sb=C * N,
           w_ld=N, engine=eng)
    _run(P, dev)
    errs = [(outg.cpu() - refg).abs().max().item(), (sc.cpu() - S).abs().max().item() / 8,
            (vT.cpu() - vT_ref).abs().max().item(), (o.cpu() - o_ref).abs().max().item()]
    print("engine", eng, "geglu/qk/vT/pv errs", errs)
    assert max(errs) < {1: 3e-2, 2: 3e-5, 3: 2e-4}[eng], errs


@pytest.mark.parametrize("eng", [2, 3])
@pytest.mark.parametrize("B,C,Cx0,Cx1,Cout,H,W", [(2, 64, 64, 32, 128, 16, 16), (16, 96, 160, 0, 192, 8, 8), (1, 64, 32, 32, 64, 9, 7)])
def test_conv_tc_fused_skip(dev, eng, B, C, Cx0, Cx1, Cout, H, W):
    """ResBlock tail (pyunet.py:299): conv3x3(h) + skip_connection 1x1(cat(x, skip)) as ONE launch - the side input's
    channels are extra K steps of the implicit GEMM."""
    from frido_b200.program import Program, Src
    g = torch.Generator().manual_seed(B + C + Cx0 + H)
    hx = torch.randn(B, C, H, W, generator=g)
    x0 = torch.randn(B, Cx0, H, W, generator=g)
    x1 = torch.randn(B, Cx1, H, W, generator=g) if Cx1 else None
    w3 = torch.randn(Cout, C, 3, 3, generator=g) / np.sqrt(9 * C)
    w1 = torch.randn(Cout, Cx0 + Cx1, 1, 1, generator=g) / np.sqrt(Cx0 + Cx1)
    b3, b1 = torch.randn(Cout, generator=g), torch.randn(Cout, generator=g)
    xin = torch.cat([x0, x1], 1) if Cx1 else x0
    ref = F.conv2d(hx.double(), w3.double(), b3.double(), padding=1) + F.conv2d(xin.double(), w1.double(), b1.double())
    P = Program(dev, "tc_skip")
    wcat = torch.cat([_pack(w3), w1.view(Cout, -1)], 1).contiguous().to(dev)
    out = torch.zeros(B, H * W, Cout, device=dev)
    side = (Src.nhwc(_nhwc(x0).to(dev), H, W), Src.nhwc(_nhwc(x1).to(dev), H, W) if Cx1 else None)
    P.conv(Src.nhwc(_nhwc(hx).to(dev), H, W), wcat, out, B=B, Hin=H, Win=W, Hout=H, Wout=W, Cout=Cout, ksize=3, pad=1,
           bias=(b3 + b1).to(dev), side=side, engine=eng)
    _run(P, dev)
    err = (out.view(B, H, W, Cout).permute(0, 3, 1, 2).cpu().double() - ref).abs().max().item()
    print("fused skip err", err)
    assert err < tol3(9 * C + Cx0 + Cx1, ref.abs().max().item()) + (1e-4 if eng == 3 else 0.0)


@pytest.mark.parametrize("B,C,H,W", [(2, 64, 16, 16), (3, 192, 64, 64), (2, 96, 9, 7)])
def test_conv_tc_stride2(dev, B, C, H, W):
    """Downsample (pyunet.py:152-156): 3x3 stride 2 pad 1 through the TMA traversal stride."""
    from frido_b200.program import Program, Src
    g = torch.Generator().manual_seed(B * C + H)
    x = torch.randn(B, C, H, W, generator=g)
    w = torch.randn(64, C, 3, 3, generator=g) / np.sqrt(9 * C)
    b = torch.randn(64, generator=g)
    ref = F.conv2d(x, w, b, stride=2, padding=1)
    Ho, Wo = ref.shape[2:]
    P = Program(dev, "tc_s2")
    out = torch.zeros(B, Ho * Wo, 64, device=dev)
    P.conv(Src.nhwc(_nhwc(x).to(dev), H, W), _pack(w).to(dev), out, B=B, Hin=H, Win=W, Hout=Ho, Wout=Wo, Cout=64, ksize=3,
           stride=2, pad=1, bias=b.to(dev), engine=2)
    _run(P, dev)
    err = (out.view(B, Ho, Wo, 64).permute(0, 3, 1, 2).cpu() - ref).abs().max().item()
    print("stride2 err", err)
    assert err < tol3(9 * C, ref.abs().max().item())


def test_tc_rejects_unsupported_shapes(dev):
    from frido_b200 import _lib as L
    from frido_b200.program import Program, Src
    P = Program(dev, "bad")
    x = torch.zeros(1, 8, 8, 3, device=dev)
    w = torch.zeros(64, 27, device=dev)
    out = torch.zeros(1, 64, 64, device=dev)
    P.conv(Src.nhwc(x, 8, 8), w, out, B=1, Hin=8, Win=8, Hout=8, Wout=8, Cout=64, ksize=3, pad=1, engine=1)
    with pytest.raises(L.FridoError):
        P.run()


@pytest.mark.parametrize("variant", ["coupled", "epi16"])
@pytest.mark.parametrize("case", [CASES[1], CASES[3], CASES[7]])
def test_bf16x3_kernel_variants_bit_identical(dev, case, variant):
    """Engine 3 has three kernels for the same launch: the default (decoupled operand rings, warp-uniform issue), the
    stage-coupled one (FRIDO_TC_DECOUPLE=0) and the 16-epilogue-warp variant (FRIDO_TC_EPI16=2).  The K order of every output
    element is the same in all of them: results must be bit-identical."""
    import os
    from frido_b200.program import Program, Src
    B, C0, C1, Cout, H, W, k = case
    g = torch.Generator().manual_seed(5 + Cout + H)
    x = torch.randn(B, C0 + C1, H, W, generator=g)
    w = torch.randn(Cout, C0 + C1, k, k, generator=g) / np.sqrt((C0 + C1) * k * k)
    bias = torch.randn(Cout, generator=g)
    res = torch.randn(B, Cout, H, W, generator=g)

    def run(env):
        old = {k_: os.environ.get(k_) for k_ in ("FRIDO_TC_DECOUPLE", "FRIDO_TC_EPI16", "FRIDO_SK")}
        os.environ.update(env)
        os.environ["FRIDO_SK"] = "0"
        try:
            P = Program(dev, "variants")
            out = torch.zeros(B, H * W, Cout, device=dev)
            P.conv(Src.nhwc(_nhwc(x).to(dev), H, W), _pack(w).to(dev), out, B=B, Hin=H, Win=W, Hout=H, Wout=W, Cout=Cout, ksize=k,
                   pad=k // 2, bias=bias.to(dev), res=_nhwc(res).to(dev).view(B, H * W, Cout), engine=3)
            _run(P, dev)
            P.run()
            torch.cuda.synchronize(dev)
            return out.cpu()
        finally:
            for k_, v in old.items():
                if v is None:
                    os.environ.pop(k_, None)
                else:
                    os.environ[k_] = v

    base = run({"FRIDO_TC_DECOUPLE": "1", "FRIDO_TC_EPI16": "0"})
    other = run({"FRIDO_TC_DECOUPLE": "0"} if variant == "coupled" else {"FRIDO_TC_DECOUPLE": "1", "FRIDO_TC_EPI16": "2"})
    ref = (F.conv2d(x.double(), w.double(), bias.double(), padding=k // 2) + res.double())
    got = base.view(B, H, W, Cout).permute(0, 3, 1, 2).double()
    assert (got - ref).abs().max().item() < tol3((C0 + C1) * k * k, ref.abs().max().item()) + 1e-4
    assert torch.equal(base, other), variant
