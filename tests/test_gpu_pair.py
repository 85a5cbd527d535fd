"""GPU parity of the CTA-pair (tcgen05.mma.cta_group::2) variant of the BF16x3 conv engine (csrc/conv_tc2.cu) against the
single-CTA kernel it specialises and against fp64 PyTorch.  The K order of every output element is the same in both kernels,
so the results must be BIT-IDENTICAL; the fp64 bound is the engine's usual one (tests/test_gpu_tc.py)."""
import os

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


CASES = [
    # B, C0, C1, Cout, H, W, k, side channels
    (2, 64, 0, 64, 16, 16, 3, 0),      # 4 M tiles, one N tile of 64
    (2, 192, 0, 192, 32, 32, 3, 0),    # BN = 192 -> 96-row W halves
    (4, 64, 32, 128, 8, 8, 3, 0),      # two sources, 2 images per tile
    (1, 64, 0, 384, 1, 512, 1, 0),     # linear, 2 N tiles
    (2, 96, 0, 192, 16, 16, 3, 64),    # fused 1x1 side input
    (16, 192, 0, 192, 64, 64, 3, 0),   # more items than clusters: persistent loop, both accumulators in flight
]


def _run_case(dev, case, pair):
    from frido_b200 import _lib as L
    from frido_b200.program import Program, Src
    B, C0, C1, Cout, H, W, k, Cs = case
    g = torch.Generator().manual_seed(17 + C0 + Cout + H)
    Cin = C0 + C1
    x0 = torch.randn(B, C0, H, W, generator=g)
    x1 = torch.randn(B, C1, H, W, generator=g) if C1 else None
    xs = torch.randn(B, Cs, H, W, generator=g) if Cs else None
    w = torch.randn(Cout, Cin, k, k, generator=g) / np.sqrt(Cin * k * k)
    ws = torch.randn(Cout, Cs, 1, 1, generator=g) / np.sqrt(max(Cs, 1)) if Cs else None
    bias = torch.randn(Cout, generator=g)
    rowvec = torch.randn(B, Cout, generator=g)
    res = torch.randn(B, Cout, H, W, generator=g)
    xin = torch.cat([x0, x1], 1) if C1 else x0
    ref = F.conv2d(xin.double(), w.double(), bias.double(), padding=k // 2)
    if Cs:
        ref = ref + F.conv2d(xs.double(), ws.double())
    ref2 = ref + res.double()
    old = {k_: os.environ.get(k_) for k_ in ("FRIDO_TC_PAIR", "FRIDO_SK")}
    os.environ["FRIDO_TC_PAIR"] = "2" if pair else "0"
    os.environ["FRIDO_SK"] = "0"
    try:
        P = Program(dev, "pair")
        a0 = Src.nhwc(_nhwc(x0).to(dev), H, W)
        a1 = Src.nhwc(_nhwc(x1).to(dev), H, W) if C1 else None
        wd = (torch.cat([_pack(w), ws.view(Cout, -1)], 1).contiguous() if Cs else _pack(w)).to(dev)
        side = (Src.nhwc(_nhwc(xs).to(dev), H, W), None) if Cs else None
        out = torch.zeros(B, H * W, Cout, device=dev)
        out2 = torch.zeros(B, H * W, Cout, device=dev)
        cs = torch.zeros(B, Cout, 2, dtype=torch.float64, device=dev)
        kw = dict(B=B, Hin=H, Win=W, Hout=H, Wout=W, Cout=Cout, ksize=k, pad=k // 2, a1=a1, bias=bias.to(dev), side=side, engine=3)
        P.conv(a0, wd, out, **kw)                                                    # EPI_BIAS
        P.conv(a0, wd, out2, res=_nhwc(res).to(dev).view(B, H * W, Cout), csum=cs if H * W >= 32 else None, **kw)  # EPI_BIAS_RES(_CS)
        P.prepare_weights()
        P.run()
        torch.cuda.synchronize(dev)
        P.run()   # replay: barriers / tensor memory come back clean
        torch.cuda.synchronize(dev)
    finally:
        for k_, v in old.items():
            if v is None:
                os.environ.pop(k_, None)
            else:
                os.environ[k_] = v
    to = lambda t: t.view(B, H, W, Cout).permute(0, 3, 1, 2).cpu()
    return to(out), to(out2), ref, ref2


@pytest.mark.parametrize("case", CASES)
def test_pair_kernel_bit_identical_to_single_cta(dev, case):
    o_p, o2_p, ref, ref2 = _run_case(dev, case, True)
    o_s, o2_s, _, _ = _run_case(dev, case, False)
    B, C0, C1, Cout, H, W, k, Cs = case
    K = (C0 + C1) * k * k + Cs
    tol = 1e-5 + 1.5e-8 * K * ref.abs().max().item() + 1e-4
    assert (o_s.double() - ref).abs().max().item() < tol
    assert (o_p.double() - ref).abs().max().item() < tol, "pair kernel vs fp64"
    assert (o2_p.double() - ref2).abs().max().item() < tol, "pair kernel (residual epilogue) vs fp64"
    assert torch.equal(o_p, o_s), "pair kernel differs from the single-CTA kernel"
    assert torch.equal(o2_p, o2_s)
