"""GPU parity, kernel level: every launcher of include/frido_b200.h against the
same op in plain PyTorch fp32 (on the CPU) or against oracle/torch_oracle.py.
Byte/index results must be bit-exact; float results within the stated tolerance."""
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
    from frido_b200 import _lib
    _lib.check(_lib.lib().frido_check_device(), "check_device")
    return torch.device("cuda:0")


def _prog(dev):
    from frido_b200.program import Program
    return Program(dev, "test", engine="simt")  # this file pins the fp32 SIMT engine; test_gpu_tc.py covers tcgen05


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def _pack(w):
    return w.permute(0, 2, 3, 1).contiguous().view(w.shape[0], -1)


CONV_CASES = [
    # B, Cin, Cout, H, W, k, stride, ups
    (2, 32, 64, 8, 8, 3, 1, 1),
    (1, 192, 192, 16, 16, 3, 1, 1),
    (2, 64, 96, 9, 7, 3, 1, 1),      # ragged spatial size
    (2, 64, 64, 8, 8, 3, 2, 1),      # Downsample (pyunet.py:152-156)
    (2, 64, 64, 7, 7, 3, 2, 1),      # odd size stride 2
    (2, 64, 32, 4, 4, 3, 1, 2),      # Upsample folded (pyunet.py:119-121)
    (3, 96, 40, 5, 5, 1, 1, 1),      # 1x1, Cout not a multiple of the tile
    (2, 3, 32, 8, 8, 3, 1, 1),       # Cin = 3 (pre_input_blocks)
    (2, 32, 3, 8, 8, 3, 1, 1),       # Cout = 3 (out head)
    (1, 6, 6, 8, 8, 1, 1, 1),        # post_quant_conv
    (2, 64, 3, 32, 32, 3, 1, 1),     # out head at >= 256 pixels: shared-memory tiled small-Cout kernel
    (1, 32, 3, 19, 21, 3, 1, 1),     # same, ragged (partial tiles, zero-filled halo)
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_simt_matches_torch(dev, case):
    from frido_b200.program import Src
    B, Cin, Cout, H, W, k, stride, ups = case
    g = torch.Generator().manual_seed(hash(case) % 2**31)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / np.sqrt(Cin * k * k)
    b = torch.randn(Cout, generator=g)
    xi = F.interpolate(x, scale_factor=2, mode="nearest") if ups == 2 else x
    # fp64 reference: the check must not depend on the host BLAS' own fp32 accumulation order
    ref = F.conv2d(xi.double(), w.double(), b.double(), stride=stride, padding=k // 2)
    Ho, Wo = ref.shape[2:]
    res = torch.randn(B, Cout, Ho, Wo, generator=g)
    rowvec = torch.randn(B, Cout, generator=g)
    ref2 = F.relu(ref + rowvec[:, :, None, None].double() + res.double())
    P = _prog(dev)
    xd, wd, bd = _nhwc(x).to(dev), _pack(w).to(dev), b.to(dev)
    out = torch.zeros(B, Ho * Wo, Cout, device=dev)
    out2 = torch.zeros(B, Ho * Wo, Cout, device=dev)
    P.conv(Src.nhwc(xd, H, W), wd, out, B=B, Hin=H, Win=W, Hout=Ho, Wout=Wo, Cout=Cout, ksize=k, stride=stride, pad=k // 2,
           ups=ups, bias=bd)
    from frido_b200 import _lib as L
    P.conv(Src.nhwc(xd, H, W), wd, out2, B=B, Hin=H, Win=W, Hout=Ho, Wout=Wo, Cout=Cout, ksize=k, stride=stride, pad=k // 2,
           ups=ups, bias=bd, rowvec=rowvec.to(dev), rowvec_sb=Cout, res=_nhwc(res).to(dev).view(B, Ho * Wo, Cout),
           act=L.ACT_RELU)
    P.run()
    got = out.view(B, Ho, Wo, Cout).permute(0, 3, 1, 2).cpu()
    got2 = out2.view(B, Ho, Wo, Cout).permute(0, 3, 1, 2).cpu()
    tol = 2e-6 * np.sqrt(Cin * k * k) * max(1.0, ref.abs().max().item())  # fp32 FFMA chain vs exact: ~sqrt(K) ulp
    e1, e2 = (got.double() - ref).abs().max().item(), (got2.double() - ref2).abs().max().item()
    assert e1 < tol and e2 < tol, (case, e1, e2, tol)


def test_conv_concat_nchw_in_out_and_geglu(dev):
    from frido_b200 import _lib as L
    from frido_b200.program import Src
    g = torch.Generator().manual_seed(5)
    B, H, W = 2, 6, 6
    a, b2 = torch.randn(B, 64, H, W, generator=g), torch.randn(B, 32, H, W, generator=g)
    w = torch.randn(48, 96, 3, 3, generator=g) / 30
    bias = torch.randn(48, generator=g)
    ref = F.conv2d(torch.cat([a, b2], 1), w, bias, padding=1)
    P = _prog(dev)
    out = torch.zeros(B, 48, H, W, device=dev)  # NCHW output
    ad, bd = _nhwc(a).to(dev), _nhwc(b2).to(dev)
    P.conv(Src.nhwc(ad, H, W), _pack(w).to(dev), out, B=B, Hin=H, Win=W, Hout=H, Wout=W, Cout=48, ksize=3, pad=1,
           a1=Src.nhwc(bd, H, W), bias=bias.to(dev), o_sb=48 * H * W, o_sp=1, o_sn=H * W)
    # NCHW strided input slice (latent channels 3..6 of 6)
    z = torch.randn(B, 6, H, W, generator=g)
    wz = torch.randn(32, 3, 3, 3, generator=g) / 5
    refz = F.conv2d(z[:, 3:6], wz, None, padding=1)
    zd = z.to(dev)
    outz = torch.zeros(B, H * W, 32, device=dev)
    P.conv(Src.nchw(zd, H, W, 3, 6), _pack(wz).to(dev), outz, B=B, Hin=H, Win=W, Hout=H, Wout=W, Cout=32, ksize=3, pad=1)
    # GEGLU linear (attention.py:37-44) with interleaved rows
    M_, K, inner = 50, 64, 40
    xt = torch.randn(M_, K, generator=g)
    wp = torch.randn(2 * inner, K, generator=g) / 8
    bp = torch.randn(2 * inner, generator=g)
    hh = F.linear(xt, wp, bp)
    refg = hh[:, :inner] * F.gelu(hh[:, inner:])
    wi = torch.stack([wp[:inner], wp[inner:]], 1).reshape(2 * inner, K).contiguous()
    bi = torch.stack([bp[:inner], bp[inner:]], 1).reshape(-1).contiguous()
    outg = torch.zeros(M_, inner, device=dev)
    P.linear(xt.to(dev), wi.to(dev), outg, M=M_, K=K, N=2 * inner, bias=bi.to(dev), act=L.ACT_GEGLU)
    P.run()
    assert (out.cpu() - ref).abs().max() < 2e-5
    assert (outz.view(B, H, W, 32).permute(0, 3, 1, 2).cpu() - refz).abs().max() < 2e-5
    assert (outg.cpu() - refg).abs().max() < 2e-5


def test_batched_matmul_attention_shapes(dev):
    """QK^T with per-image weights inside a fused q|k tensor, V^T transposed store, P.V with padded keys."""
    from frido_b200.program import Src
    g = torch.Generator().manual_seed(6)
    B, N, C, Lc, Lp = 2, 20, 32, 5, 32
    qk = torch.randn(B, N, 2 * C, generator=g)
    S = torch.einsum("bid,bjd->bij", qk[..., :C], qk[..., C:])
    P = _prog(dev)
    qkd = qk.to(dev)
    sc = torch.zeros(B, N, N, device=dev)
    P.conv(Src(qkd, C, N * 2 * C, 0, 2 * C, 1), qkd, sc, B=B, Hin=1, Win=N, Hout=1, Wout=N, Cout=N, w_sb=N * 2 * C,
           w_ld=2 * C, w_off=C)
    x = torch.randn(B, N, C, generator=g)
    wv = torch.randn(C, C, generator=g) / 6
    vT_ref = F.linear(x, wv).transpose(1, 2).contiguous()
    vT = torch.zeros(B, C, N, device=dev)
    P.conv(Src(x.to(dev), C, N * C, 0, C, 1), wv.to(dev), vT, B=B, Hin=1, Win=N, Hout=1, Wout=N, Cout=C, o_sb=C * N,
           o_sp=1, o_sn=N)
    prob = torch.zeros(B, N, Lp)
    prob[..., :Lc] = torch.rand(B, N, Lc, generator=g)
    vTc = torch.zeros(B, C, Lp)
    vTc[..., :Lc] = torch.randn(B, C, Lc, generator=g)
    o_ref = torch.einsum("bij,bdj->bid", prob, vTc)
    o = torch.zeros(B, N, C, device=dev)
    P.conv(Src(prob.to(dev), Lc, N * Lp, 0, Lp, 1), vTc.to(dev), o, B=B, Hin=1, Win=N, Hout=1, Wout=N, Cout=C, w_sb=C * Lp,
           w_ld=Lp)
    P.run()
    assert (sc.cpu() - S).abs().max() < 2e-5
    assert (vT.cpu() - vT_ref).abs().max() < 2e-5
    assert (o.cpu() - o_ref).abs().max() < 2e-5


@pytest.mark.parametrize("C0,C1,HW,spade,silu", [(64, 0, 64, False, True), (192, 0, 1024, True, True), (576, 384, 16, True, False),
                                                 (960, 960, 64, False, True), (1920, 0, 16, False, True), (128, 0, 65536, False, True)])
def test_groupnorm_spade_silu(dev, C0, C1, HW, spade, silu):
    g = torch.Generator().manual_seed(C0 + C1 + HW)
    B = 2
    C = C0 + C1
    x0 = torch.randn(B, HW, C0, generator=g) * 2 + 0.5
    x1 = torch.randn(B, HW, C1, generator=g) - 1.0 if C1 else None
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    gb = torch.randn(B, HW, 2 * C, generator=g) * 0.3 if spade else None
    xc = torch.cat([x0, x1], -1) if C1 else x0
    xn = xc.permute(0, 2, 1).reshape(B, C, HW, 1)
    ref = F.group_norm(xn, 32, gamma, beta, 1e-6)
    if spade:
        G = gb[..., :C].permute(0, 2, 1).reshape(B, C, HW, 1)
        Bt = gb[..., C:].permute(0, 2, 1).reshape(B, C, HW, 1)
        ref = ref * (1 + G) + Bt
    if silu:
        ref = F.silu(ref)
    ref = ref.reshape(B, C, HW).permute(0, 2, 1)
    P = _prog(dev)
    sums = torch.zeros(B, 32, 2, dtype=torch.float64, device=dev)
    out = torch.zeros(B, HW, C, device=dev)
    x0d = x0.to(dev)
    x1d = x1.to(dev) if C1 else None
    P.zero(sums)
    P.gn_stats(x0d, C0, sums, B=B, HW=HW, a1=x1d, c1=C1)
    P.norm_act(x0d, C0, sums, gamma.to(dev), beta.to(dev), out, B=B, HW=HW, eps=1e-6, a1=x1d, c1=C1,
               gb=None if gb is None else gb.to(dev), silu=int(silu))
    P.run()
    P.run()  # re-run: the zero op must make the program idempotent
    assert (out.cpu() - ref).abs().max() < 2e-5


@pytest.mark.parametrize("B,HW,C", [(2, 4096, 192), (3, 77, 32), (1, 256, 1920)])
def test_chan_stats_feeds_norm_act(dev, B, HW, C):
    """gn_stats in per-channel mode produces what a conv epilogue's chan_sums would: norm_act on top of it == GroupNorm."""
    g = torch.Generator().manual_seed(B + HW + C)
    x = torch.randn(B, HW, C, generator=g) * 1.7 + 0.4
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    ref = F.group_norm(x.permute(0, 2, 1).reshape(B, C, HW, 1), 32, gamma, beta, 1e-5).reshape(B, C, HW).permute(0, 2, 1)
    P = _prog(dev)
    cs = torch.zeros(B, C, 2, dtype=torch.float64, device=dev)
    out = torch.zeros(B, HW, C, device=dev)
    xd = x.to(dev)
    P.zero(cs)
    P.chan_stats(xd, C, cs, B=B, HW=HW)
    P.norm_act(xd, C, None, gamma.to(dev), beta.to(dev), out, B=B, HW=HW, eps=1e-5, silu=0, csum0=cs)
    P.run()
    P.run()
    xx = x.double()
    assert (cs.cpu()[..., 0] - xx.sum(1)).abs().max() < 1e-3 and (cs.cpu()[..., 1] - (xx * xx).sum(1)).abs().max() < 1e-2
    assert (out.cpu() - ref).abs().max() < 2e-5


@pytest.mark.parametrize("C", [32, 384, 576, 960])
def test_layernorm(dev, C):
    g = torch.Generator().manual_seed(C)
    x = torch.randn(37, C, generator=g) * 3 + 1
    gm, bt = torch.randn(C, generator=g), torch.randn(C, generator=g)
    out = torch.zeros(37, C, device=dev)
    P = _prog(dev)
    P.layernorm(x.to(dev), gm.to(dev), bt.to(dev), out, rows=37, Cdim=C)
    P.run()
    assert (out.cpu() - F.layer_norm(x, (C,), gm, bt, 1e-5)).abs().max() < 2e-5


@pytest.mark.parametrize("M,K,N", [(16, 768, 832), (3, 192, 768), (1, 768, 14976)])
def test_linear_small_m(dev, M, K, N):
    """time_embed / emb_layers (pyunet.py:561-565,225-231): M = batch rows through the weight-streaming kernel."""
    from frido_b200 import _lib as L
    g = torch.Generator().manual_seed(M + K + N)
    x = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / np.sqrt(K)
    b, rv = torch.randn(N, generator=g), torch.randn(N, generator=g)
    P = _prog(dev)
    out = torch.empty(M, N, device=dev)
    P.linear(x.to(dev), w.to(dev), out, M=M, K=K, N=N, bias=b.to(dev), rowvec=rv.to(dev), act=L.ACT_SILU)
    P.run()
    ref = F.silu(x.double() @ w.double().t() + b.double() + rv.double())
    assert (out.cpu().double() - ref).abs().max() < 2e-5


ATTN_CASES = [
    # B, N, Nk, C, packed qkv, fused LayerNorm + bias + residual
    (2, 1024, 26, 384, False, True),    # cross-attention sub-block, 26 layout tokens, 32x32 level (4 rows per warp)
    (16, 256, 26, 576, False, True),    # 16x16 level (2 rows per warp)
    (2, 64, 26, 960, False, True),      # 8x8 level: K/V chunk of 26 keys fills shared memory
    (2, 1024, 26, 384, False, False),
    (3, 64, 64, 960, True, False),      # self-attention of the 8x8 level: 3 key chunks, online softmax rescale
    (2, 16, 16, 960, True, False),      # 4x4 level of the f16f8 configs
    (1, 37, 5, 24, False, True),        # ragged rows, tiny C (parity fixtures)
    (2, 50, 50, 132, True, False),      # C/4 not a multiple of 32, keys not a multiple of the chunk
    (2, 50, 45, 132, False, True),
    (1, 9, 1, 768, False, False),       # a single key (CLIP pooled condition): softmax == 1
    (1, 40, 64, 384, False, True),      # 64 keys with 4 rows per warp: two chunks
]


@pytest.mark.parametrize("case", ATTN_CASES)
def test_attn_small_matches_torch(dev, case):
    """CrossAttention.forward core (attention.py:178-191): einsum * scale -> softmax -> einsum, optionally with the
    block's LayerNorm in front and the to_out bias + residual behind (attention.py:324); fp64 reference."""
    B, N, Nk, C, packed, fused = case
    g = torch.Generator().manual_seed(B * 1000 + N + Nk + C)
    scale = C ** -0.5
    P = _prog(dev)
    out = torch.empty(B, N, C, device=dev)
    kw, ln_w, ln_b, bias = {}, None, None, None
    if packed:
        assert N == Nk and not fused
        qkv = torch.randn(B, N, 3 * C, generator=g) * 2
        q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
        d = qkv.to(dev)
        P.attn_small(d, d, d, out, B=B, N=N, Nk=Nk, Cdim=C, scale=scale, q_sb=N * 3 * C, q_ld=3 * C, k_off=C, k_sb=N * 3 * C,
                     k_ld=3 * C, v_off=2 * C, v_sb=N * 3 * C, v_ld=3 * C)
    else:
        q = torch.randn(B, N, C, generator=g) * 2 + 0.3
        k = torch.randn(B, Nk, C, generator=g) * 2
        v = torch.randn(B, Nk, C, generator=g)
        qd = q.to(dev)
        if fused:
            ln_w, ln_b, bias = torch.randn(C, generator=g), torch.randn(C, generator=g), torch.randn(C, generator=g)
            ln2_w, ln2_b = torch.randn(C, generator=g), torch.randn(C, generator=g)
            out2 = torch.empty(B, N, C, device=dev)
            kw = dict(ln=(ln_w.to(dev), ln_b.to(dev)), bias=bias.to(dev), res=qd, ln2=(ln2_w.to(dev), ln2_b.to(dev)), out2=out2)
        P.attn_small(qd, k.to(dev), v.to(dev), out, B=B, N=N, Nk=Nk, Cdim=C, scale=scale, q_sb=N * C, q_ld=C,
                     k_sb=Nk * C, k_ld=C, v_sb=Nk * C, v_ld=C, **kw)
    P.run()
    torch.cuda.synchronize()
    qq = F.layer_norm(q.double(), (C,), ln_w.double(), ln_b.double(), 1e-5) if fused else q.double()
    sim = torch.einsum("bid,bjd->bij", qq, k.double()) * scale
    ref = torch.einsum("bij,bjd->bid", sim.softmax(-1), v.double())
    if fused:
        ref = ref + bias.double() + q.double()
        ref2 = F.layer_norm(ref, (C,), ln2_w.double(), ln2_b.double(), 1e-5)  # the block's next LayerNorm, second output
        assert (out2.cpu().double() - ref2).abs().max() < 5e-5
    assert (out.cpu().double() - ref).abs().max() < 3e-5


@pytest.mark.parametrize("n,ld", [(5, 32), (26, 32), (64, 64), (1024, 1024), (4096, 4096), (1500, 1504)])
def test_softmax(dev, n, ld):
    g = torch.Generator().manual_seed(n)
    rows = 19
    s = torch.zeros(rows, ld)
    s[:, :n] = torch.randn(rows, n, generator=g) * 4
    sd = s.to(dev)
    P = _prog(dev)
    P.softmax(sd, rows=rows, n=n, ld=ld, scale=0.37)
    P.run()
    ref = torch.softmax(s[:, :n] * 0.37, -1)
    got = sd.cpu()
    assert (got[:, :n] - ref).abs().max() < 1e-6
    assert (got[:, n:] == 0).all()  # pad columns untouched


def test_time_embed(dev):
    from oracle import torch_oracle as O
    t = torch.tensor([996, 1, 501, 0, 999], dtype=torch.int64)
    out = torch.zeros(5, 192, device=dev)
    P = _prog(dev)
    P.time_embed(t.to(dev), out, B=5, dim=192)
    P.run()
    assert (out.cpu() - O.timestep_embedding(t, 192)).abs().max() < 2e-6


def _sched(S, eta):
    from oracle import torch_oracle as O
    acp = O.alphas_cumprod().astype(np.float32)
    return O.ddim_schedule(S, eta, acp)


@pytest.mark.parametrize("eta,cfg", [(0.0, False), (0.7, False), (0.0, True), (1.0, True)])
def test_ddim_update_bit_exact(dev, eta, cfg):
    """a3: the fused update must reproduce ddim.py:243-268 bit for bit (fp32, same op order)."""
    from oracle import torch_oracle as O
    g = torch.Generator().manual_seed(11)
    B, H, W, start, end = 3, 8, 8, 3, 6
    sch = _sched(50, eta)
    T = len(sch["timesteps"])
    coef = np.stack([sch["a_t"], sch["a_prev"], sch["sigma"], sch["sqrt_1m"]], 1)[::-1].copy()
    x = torch.randn(B, end, H, W, generator=g)
    e_c = torch.randn(B, end - start, H, W, generator=g)
    e_u = torch.randn(B, end - start, H, W, generator=g)
    nz = torch.randn(B, end, H, W, generator=g)
    for i in (0, 7, T - 1):
        index = T - 1 - i
        e = O.cfg_combine(e_c, e_u, 1.5) if cfg else e_c
        xp_ref, p0_ref = O.ddim_update(x, e, sch, index, start, nz)
        P = _prog(dev)
        step = torch.tensor([i], dtype=torch.int32, device=dev)
        xp, p0 = torch.zeros(B, end, H, W, device=dev), torch.zeros(B, end, H, W, device=dev)
        P.update(x.to(dev), e_c.to(dev), torch.from_numpy(coef).to(dev), step, xp, B=B, c_start=start, c_end=end, HW=H * W,
                 eps_uncond=e_u.to(dev) if cfg else None, cfg_scale=1.5, noise=nz.to(dev), pred_x0=p0)
        P.run()
        assert torch.equal(xp.cpu(), xp_ref), f"x_prev differs at step {i}"
        assert torch.equal(p0.cpu(), p0_ref), f"pred_x0 differs at step {i}"
        assert int(step.item()) == i + 1


def test_plms_history_bit_exact(dev):
    """a4: Adams-Bashforth orders 1-4 through the device ring buffer (plms.py:285-299)."""
    from oracle import torch_oracle as O
    g = torch.Generator().manual_seed(12)
    B, H, W, start, end = 2, 4, 4, 0, 3
    sch = _sched(10, 0.0)
    T = len(sch["timesteps"])
    coef = torch.from_numpy(np.stack([sch["a_t"], sch["a_prev"], sch["sigma"], sch["sqrt_1m"]], 1)[::-1].copy()).to(dev)
    x = torch.randn(B, end, H, W, generator=g)
    eps_list = [torch.randn(B, end, H, W, generator=g) for _ in range(T + 1)]
    # oracle trajectory
    xr, old = x.clone(), []
    xd = x.to(dev).clone()
    step = torch.zeros(1, dtype=torch.int32, device=dev)
    hist = torch.zeros(3, B, end, H, W, device=dev)
    save = torch.zeros(B, end, H, W, device=dev)
    x_orig = torch.zeros(B, end, H, W, device=dev)
    for i in range(T):
        index = T - 1 - i
        e_t = eps_list[i]
        if i == 0:
            e_next = eps_list[T]
            e_p = O.plms_eps_prime(e_t, old, e_next)
        else:
            e_p = O.plms_eps_prime(e_t, old)
        xr, _ = O.ddim_update(xr, e_p, sch, index, start)
        old.append(e_t)
        if len(old) >= 4:
            old.pop(0)
        P = _prog(dev)
        if i == 0:
            x_orig.copy_(xd)
            P.update(x_orig, e_t.to(dev), coef, step, xd, B=B, c_start=start, c_end=end, HW=H * W, plms_order=4, plms_mode=1,
                     advance=0, hist=hist, eps_save=save)
            P.update(x_orig, eps_list[T].to(dev), coef, step, xd, B=B, c_start=start, c_end=end, HW=H * W, plms_order=4,
                     plms_mode=2, advance=1, hist=hist, eps_save=save)
        else:
            P.update(xd, e_t.to(dev), coef, step, xd, B=B, c_start=start, c_end=end, HW=H * W, plms_order=4, plms_mode=0,
                     advance=1, hist=hist, eps_save=save)
        P.run()
        assert torch.equal(xd.cpu(), xr), f"PLMS x_prev differs at step {i}"


@pytest.mark.parametrize("n", [1, 2])
def test_stage_snap_bit_exact(dev, n):
    from oracle import torch_oracle as O
    x = torch.randn(2, 9, 16, 16, generator=torch.Generator().manual_seed(n))
    ref = O.stage_snap(x, 3, 6, n)
    xd = x.to(dev)
    P = _prog(dev)
    P.snap(xd, B=2, Ctot=9, H=16, W=16, c_start=3, c_end=6, n=n)
    P.run()
    assert torch.equal(xd.cpu(), ref)


@pytest.mark.parametrize("n_e,e_dim,scale", [(4096, 3, 1.5), (8192, 4, 1.0), (64, 3, 3.0)])
def test_vq_indices_bit_exact(dev, n_e, e_dim, scale):
    """a14: argmin indices bit-exact vs the reference expression (quantize.py:276-280), incl. the
    straight-through value z + (e - z) and the fine->coarse channel placement."""
    from oracle import torch_oracle as O
    g = torch.Generator().manual_seed(n_e)
    B, H, W = 4, 32, 32
    cb = torch.randn(n_e, e_dim, generator=g)
    z = torch.randn(B, 2 * e_dim, H, W, generator=g) * scale
    sf = 0.8
    zs = z.clone()
    zs[:, e_dim:] *= 1.0 / torch.tensor(sf)
    zq_ref, idx_ref = O.vq_lookup(zs[:, e_dim:], cb)
    out = torch.zeros(B, H * W, 2 * e_dim, device=dev)
    idx = torch.zeros(B * H * W, dtype=torch.int64, device=dev)
    P = _prog(dev)
    P.vq(z.to(dev), cb.to(dev), out, idx, B=B, C_total=2 * e_dim, HW=H * W, c_start=e_dim, e_dim=e_dim, scale_factor=sf,
         out_C=2 * e_dim, out_coff=0)
    P.run()
    assert torch.equal(idx.cpu(), idx_ref)
    got = out.view(B, H, W, 2 * e_dim)[..., :e_dim].permute(0, 3, 1, 2).cpu()
    assert torch.equal(got, zq_ref)


def test_errors_are_loud(dev):
    from frido_b200 import _lib as L
    from frido_b200.program import Src
    P = _prog(dev)
    x = torch.zeros(1, 4, 4, 30, device=dev)
    sums = torch.zeros(1, 32, 2, dtype=torch.float64, device=dev)
    P.gn_stats(x, 30, sums, B=1, HW=16)  # 30 channels: not divisible into 32 groups
    with pytest.raises(L.FridoError):
        P.run()


@pytest.mark.parametrize("mode", ["np", "pil"])
def test_images_to_uint8_bit_exact(dev, mode):
    """output formatting identical to custom_to_np / custom_to_pil (scripts/sample_diffusion.py:103-121)."""
    import frido_b200 as fb
    g = torch.Generator().manual_seed(4)
    x = torch.randn(3, 3, 37, 41, generator=g) * 0.8
    x[0, 0, 0, :8] = torch.tensor([-1.0, 1.0, -1.5, 1.5, 0.0, 0.999999, -0.999999, 0.5])
    if mode == "np":
        ref = ((x + 1) * 127.5).clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()
    else:
        y = (torch.clamp(x, -1.0, 1.0) + 1.0) / 2.0
        ref = torch.from_numpy((255 * y.permute(0, 2, 3, 1).numpy()).astype(np.uint8))
    got = fb.images_to_uint8(x.to(dev), mode).cpu()
    assert torch.equal(got, ref)
