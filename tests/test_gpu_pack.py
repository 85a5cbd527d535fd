"""GPU tests of the weight-packing entry points of the C ABI (csrc/pack.cu via frido_b200/packing.py) against the PyTorch
expressions they replace.  Permutes / concatenations are copies: bit-exact.  The attention folds are fp64 sums rounded once:
they may differ from torch's fp64 GEMM by the rounding of a near-tie, bounded here at 1 ulp (2^-23 relative)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def test_conv_weight_and_concats_bit_exact(dev):
    from frido_b200 import packing as PK
    g = torch.Generator().manual_seed(3)
    w3 = torch.randn(48, 20, 3, 3, generator=g).to(dev)
    w3b = torch.randn(16, 20, 3, 3, generator=g).to(dev)
    w1 = torch.randn(48, 12, 1, 1, generator=g).to(dev)
    ref = w3.permute(0, 2, 3, 1).contiguous().view(48, -1)
    assert torch.equal(PK.conv_weight(w3), ref)
    assert torch.equal(PK.conv_rows([w3, w3b]), torch.cat([ref, w3b.permute(0, 2, 3, 1).contiguous().view(16, -1)], 0))
    assert torch.equal(PK.conv_plus_side(w3, w1), torch.cat([ref, w1.view(48, 12)], 1))
    a, b = torch.randn(7, 33, generator=g).to(dev), torch.randn(5, 33, generator=g).to(dev)
    assert torch.equal(PK.cat_rows([a, b]), torch.cat([a, b], 0))
    assert torch.equal(PK.cat_rows([a[0], b[1]]), torch.cat([a[0], b[1]], 0))
    assert torch.equal(PK.transpose(a), a.t().contiguous())
    big = torch.randn(14, 33, generator=g).to(dev)
    assert torch.equal(PK.interleave_rows(big[:7], big[7:]), torch.stack([big[:7], big[7:]], 1).reshape(14, 33))
    assert torch.equal(PK.interleave_rows(big[:7, 0], big[7:, 0]), torch.stack([big[:7, 0], big[7:, 0]], 1).reshape(-1))
    assert torch.equal(PK.copy(big[3]), big[3])
    assert torch.equal(PK.copy(w1.view(48, 12)), w1.view(48, 12))
    assert torch.equal(PK.vec_add(a[0], b[0]), a[0] + b[0])
    dst = torch.zeros(1, 14 * 33, device=dev)
    PK.place(big.view(1, -1), dst)
    assert torch.equal(dst.view(14, 33), big)


def test_fold_self_attention_matches_fp64(dev):
    from frido_b200 import packing as PK
    g = torch.Generator().manual_seed(4)
    for Cd in (64, 200, 384):
        wq, wk, wv, wo = (torch.randn(Cd, Cd, generator=g).to(dev) / Cd ** 0.5 for _ in range(4))
        a, v = PK.fold_self_attention(wq, wk, wv, wo)
        ra = (wk.double().t() @ wq.double())
        rv = (wo.double() @ wv.double())
        for got, ref in ((a, ra), (v, rv)):
            err = (got.double() - ref).abs()
            assert (err <= 2.0 ** -23 * ref.abs() + 1e-12).all(), (Cd, err.max().item())


def test_module_folds_are_the_expressions_the_cpu_test_pins(dev):
    """frido_b200.unet.fold_self_attention / fold_cross_attention_weights on a CUDA module = the expressions that
    tests/test_oracle_golden.py::test_attention_weight_folds_match_the_oracle checks against the oracle's CrossAttention."""
    from frido_b200 import modules as M
    from frido_b200.unet import fold_self_attention, fold_cross_attention_weights
    torch.manual_seed(0)
    blk = M.BasicTransformerBlock(96, 40).to(dev)
    for prm in blk.parameters():
        prm.data.normal_(0, 0.2)
    ca = blk.attn1
    wq, wk, wv, wo = (m.weight.detach().double() for m in (ca.to_q, ca.to_k, ca.to_v, ca.to_out[0]))
    a, v = fold_self_attention(ca)
    assert (a.double() - wk.t() @ wq).abs().max().item() <= 2.0 ** -22 * (wk.t() @ wq).abs().max().item()
    assert (v.double() - wo @ wv).abs().max().item() <= 2.0 ** -22 * (wo @ wv).abs().max().item()
    wq_t, wo2 = fold_cross_attention_weights(blk.attn2)
    assert torch.equal(wq_t, blk.attn2.to_q.weight.detach().t().contiguous())
    assert torch.equal(wo2, blk.attn2.to_out[0].weight.detach())
