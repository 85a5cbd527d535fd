"""Pins oracle/torch_oracle.py (the CPU restatement) against outputs of the
UNMODIFIED reference, generated in the build container by oracle/make_golden.py
(the reference itself has no tests or golden vectors — SURVEY.md §4)."""
import os

import numpy as np
import pytest
import torch

from oracle import synth
from oracle import torch_oracle as O

torch.set_grad_enabled(False)
TOL = 2e-5  # same algorithm, same fp32 CPU kernels; only op-order differences


def _load(golden_dir, name):
    p = os.path.join(golden_dir, name)
    if not os.path.exists(p):
        pytest.skip(f"{name} not generated")
    return torch.load(p, weights_only=False)


def test_schedule_tables(golden_dir):
    g = _load(golden_dir, "sched.pt")
    acp = O.alphas_cumprod()
    assert np.array_equal(acp.astype(np.float32), g["alphas_cumprod"].numpy())
    for S in (200, 250, 100, 50, 4):
        for eta in (0.0, 1.0):
            ref = g[f"S{S}_eta{eta}"]
            sch = O.ddim_schedule(S, eta, acp.astype(np.float32))
            assert np.array_equal(sch["timesteps"], ref["timesteps"].numpy())
            tab = np.stack([sch["a_t"], sch["a_prev"], sch["sigma"], sch["sqrt_1m"]], 1)
            # bit-exact: these scalars feed every step
            assert np.array_equal(tab, ref["table"].numpy()), (S, eta)


@pytest.mark.parametrize("tag", ["tiny2", "tiny3"])
def test_tiny_model_vs_reference(golden_dir, tag):
    g = _load(golden_dir, f"{tag}.pt")
    sd = synth.synth_state_dict(g["manifest"], g["seed"])
    split, B = g["split"], g["B"]
    ns = len(split)
    C = sum(split)
    ctx = synth.synth_input("ctx", (B, 5, 24), 1)
    uc = synth.synth_input("uc", (B, 5, 24), 2)
    for s in range(ns):
        x = synth.synth_input(f"x{s}", (B, 3 * (s + 1), 8, 8), 4)
        for t in (996, 1):
            e = O.unet_forward(sd, x, torch.full((B,), t, dtype=torch.long), ctx, s, split)
            assert (e - g[f"eps_s{s}_t{t}"]).abs().max() < TOL
    tr = []
    out = O.sample(sd, split, ctx, g["ddim4_xinit"], 4, trace=tr)
    assert (out - g["ddim4_out"]).abs().max() < 1e-4
    assert (tr[0][2] - g["ddim4_xinter1"]).abs().max() < 1e-4
    assert (tr[0][3] - g["ddim4_predx0_1"]).abs().max() < 2e-3  # /sqrt(a_t): ~x30 amplification at t=751
    noises = [synth.synth_input(f"nz{k}", (B, C, 8, 8), 5) for k in range(4 * ns)]
    noises_s = []
    for s in range(ns):
        for i in range(4):
            noises_s.append(noises[s * 4 + i][:, : sum(split[: s + 1])])
    out = O.sample(sd, split, ctx, g["ddim4e_xinit"], 4, eta=0.7, noises=noises_s)
    assert (out - g["ddim4e_out"]).abs().max() < 1e-4
    out = O.sample(sd, split, ctx, g["plms5_xinit"], 5, sampler="plms")
    assert (out - g["plms5_out"]).abs().max() < 1e-4
    out = O.sample(sd, split, ctx, g["cfg2_xinit"], 2, uc=uc, cfg_scale=1.5)
    assert (out - g["cfg2_out"]).abs().max() < 1e-4
    out = O.sample(sd, split, ctx, g["plmscfg4_xinit"], 4, sampler="plms", uc=uc, cfg_scale=1.5)
    assert (out - g["plmscfg4_out"]).abs().max() < 1e-4
    # decode: indices bit-exact, image close
    for zk, ik, ck in (("ddim4_out", "dec_img", "dec_codes"), ("dec2_z", "dec2_img", "dec2_codes")):
        img, codes = O.decode_first_stage(sd, g[zk], split, g["scale_factor"].tolist())
        for a, b in zip(codes, g[ck]):
            assert torch.equal(a, b)
        assert (img - g[ik]).abs().max() < 1e-4


def test_mask_x0_branch_vs_reference(golden_dir):
    """8f.3: inpainting blend (ddim.py:158-161, plms.py:162-165) with x_T given (stage 0 skipped), injected q_sample noise."""
    g = _load(golden_dir, "mask.pt")
    t2 = _load(golden_dir, "tiny2.pt")
    sd = synth.synth_state_dict(t2["manifest"], t2["seed"])
    B = 2
    ctx = synth.synth_input("ctx", (B, 5, 24), 1)
    uc = synth.synth_input("uc", (B, 5, 24), 2)
    for tag, S, kw in (("ddim4", 4, {}), ("plms4", 4, dict(sampler="plms")), ("ddimcfg2", 2, dict(uc=uc, cfg_scale=1.5))):
        tr = []
        out = O.sample(sd, [3, 3], ctx, g["x_T"], S, mask=g["mask"], x0=g["x0"], mask_noises=g[tag + "_noises"], x_T_skip=True,
                       trace=tr, **kw)
        assert (out - g[tag + "_out"]).abs().max() < 1e-4, tag
        assert (tr[0][2] - g[tag + "_xinter1"]).abs().max() < 1e-4, tag
    # the blend really acts: without it the result differs
    plain = O.sample(sd, [3, 3], ctx, g["x_T"], 4, x_T_skip=True)
    assert (plain - g["ddim4_out"]).abs().max() > 1e-2


@pytest.mark.slow
def test_full_size_l2i_unet_step_and_decoder(golden_dir):
    """BASELINE config 1: single DDIM step on the full-size 511 M-parameter
    UNet at a 32x32 latent, plus the full-size f8f4 decoder."""
    g = _load(golden_dir, "l2i32.pt")
    sd = synth.synth_state_dict(g["manifest"], g["seed"])
    ctx = synth.synth_input("ctx", (1, 26, 640), 1)
    acp = O.alphas_cumprod().astype(np.float32)
    sch = O.ddim_schedule(200, 0.0, acp)
    for s in (0, 1):
        x = synth.synth_input(f"x{s}", (1, 3 * (s + 1), 32, 32), 2)
        for t in (996, 1):
            e = O.unet_forward(sd, x, torch.full((1,), t, dtype=torch.long), ctx, s, [3, 3])
            assert (e - g[f"eps_s{s}_t{t}"]).abs().max() < 5e-5
            if t == 996:
                xp, p0 = O.ddim_update(x, e, sch, 199, 3 * s)
                assert (xp - g[f"step_s{s}_xprev"]).abs().max() < 1e-5
                assert (p0 - g[f"step_s{s}_predx0"]).abs().max() < 2e-3
    del sd
    sd = synth.synth_state_dict(g["dec_manifest"], g["seed"])
    img, codes = O.decode_first_stage(sd, g["dec_z"], [3, 3], g["scale_factor"].tolist())
    for a, b in zip(codes, g["dec_codes"]):
        assert torch.equal(a, b)
    assert (img - g["dec_img"]).abs().max() < 1e-4


@pytest.mark.parametrize("tag", ["small", "full"])
def test_bert_embedder_vs_reference(golden_dir, tag):
    """the condition encoder restatement against the reference's BERTEmbedder (x-transformer encoder)."""
    g = _load(golden_dir, "bert.pt")[tag]
    sd = synth.synth_state_dict(g["manifest"], g["seed"])
    z = O.bert_embedder(sd, g["tokens"])
    assert (z - g["z"]).abs().max() < 5e-5


@pytest.mark.parametrize("tag", ["tiny2", "tiny3", "l2i"])
def test_encode_first_stage_vs_reference(golden_dir, tag):
    """MSEncoder + VQModelInterface.encode + get_first_stage_encoding against the reference's outputs
    (pre-quantisation latent, scale-factor multiply, and the code indices every scale's quantiser picked)."""
    g = _load(golden_dir, "enc.pt")[tag]
    if tag == "l2i":
        man, seed, sf, ed = g["manifest"], g["seed"], g["scale_factor"], g["fs_params"]["embed_dim"]
    else:
        t = _load(golden_dir, f"{tag}.pt")
        man, seed, sf, ed = t["manifest"], t["seed"], t["scale_factor"], t["split"]
    sd = synth.synth_state_dict([(n, s) for n, s in man if n.startswith("first_stage_model.")], seed)
    h, codes = O.encode_first_stage(sd, g["x"], list(ed))
    for a, b in zip(codes, g["codes"]):
        assert torch.equal(a.reshape(-1), b.reshape(-1))
    assert (h - g["h"]).abs().max() < TOL
    z = O.first_stage_encoding(h, list(ed), sf)
    assert (z - g["z"]).abs().max() < TOL


def test_attention_weight_folds_match_the_oracle():
    """The UNet plan re-associates the attention matmuls (frido_b200/unet.py: fold_self_attention,
    fold_cross_attention_weights) so that the key projection and to_out leave the per-step work.  Exact algebra: the
    folded form must reproduce the oracle's CrossAttention (attention.py:170-193) to fp32 rounding, for the
    self-attention and for the cross-attention against a short condition.  The folds themselves are native launches
    (csrc/pack.cu) and are compared with these same expressions on the GPU (tests/test_gpu_pack.py)."""
    import torch.nn.functional as F
    from frido_b200 import modules as M
    from oracle import torch_oracle as O

    def fold_self_attention(ca):   # (Wk^T Wq, Wo Wv), fp64 products rounded once
        wq, wk, wv, wo = (m.weight.detach().double() for m in (ca.to_q, ca.to_k, ca.to_v, ca.to_out[0]))
        return (wk.t() @ wq).float(), (wo @ wv).float()

    def fold_cross_attention_weights(ca):   # (Wq^T, Wo)
        return ca.to_q.weight.detach().t().contiguous(), ca.to_out[0].weight.detach()

    torch.manual_seed(0)
    C, D, N, L, B = 96, 40, 50, 7, 2
    blk = M.BasicTransformerBlock(C, D)
    for prm in blk.parameters():
        prm.data.normal_(0, 0.2)
    sd = {"b." + k: v.detach() for k, v in blk.state_dict().items()}
    x = torch.randn(B, N, C)
    ctx = torch.randn(B, L, D)
    scale = C ** -0.5
    # self-attention sub-block: h1 = x + to_out(attn1(LN(x)))
    ln = F.layer_norm(x, (C,), sd["b.norm1.weight"], sd["b.norm1.bias"], 1e-5)
    ref1 = O.cross_attention(ln, None, sd, "b.attn1") + x
    w_a, w_v = fold_self_attention(blk.attn1)
    t = F.linear(ln, w_a)                                    # x (Wk^T Wq)^T
    p = (torch.einsum("bid,bjd->bij", t, ln) * scale).softmax(-1)   # keys are the LayerNorm output itself
    got1 = torch.einsum("bij,bjd->bid", p, F.linear(ln, w_v)) + sd["b.attn1.to_out.0.bias"] + x
    assert (got1 - ref1).abs().max() < 2e-5
    # cross-attention sub-block: h2 = h1 + to_out(attn2(LN(h1), ctx))
    ln2 = F.layer_norm(ref1, (C,), sd["b.norm2.weight"], sd["b.norm2.bias"], 1e-5)
    ref2 = O.cross_attention(ln2, ctx, sd, "b.attn2") + ref1
    wq_t, wo = fold_cross_attention_weights(blk.attn2)
    kf = F.linear(F.linear(ctx, sd["b.attn2.to_k.weight"]), wq_t)   # K' = (ctx Wk^T) Wq
    vf = F.linear(F.linear(ctx, sd["b.attn2.to_v.weight"]), wo)     # V' = (ctx Wv^T) Wo^T
    p2 = (torch.einsum("bid,bjd->bij", ln2, kf) * scale).softmax(-1)
    got2 = torch.einsum("bij,bjd->bid", p2, vf) + sd["b.attn2.to_out.0.bias"] + ref1
    assert (got2 - ref2).abs().max() < 2e-5
