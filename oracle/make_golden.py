"""TEST INFRASTRUCTURE ONLY — mint golden vectors by running the UNMODIFIED
reference (/root/reference, imported via oracle/ref_loader.py) on CPU.

Run in the build container only:   python oracle/make_golden.py
Writes small fixtures to tests/golden/*.pt.  Weights are not stored: each
fixture carries the parameter manifest [(name, shape)] and a seed, and
oracle/synth.py regenerates the same values on the checker side.
"""
import copy
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader, synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
torch.set_grad_enabled(False)


def tiny_cfg(base, num_stage=2):
    m = copy.deepcopy(base["model"])
    p = m["params"]
    split = [3] * num_stage
    p["image_size"] = 8
    p["channels"] = 3 * num_stage
    p["stage_loss_ratio"] = [1.0 / num_stage] * num_stage
    u = p["unet_config"]["params"]
    u.update(image_size=8, in_channels=3 * num_stage, out_channels=3 * num_stage, model_channels=32,
             attention_resolutions=[2], num_res_blocks=1, channel_mult=[1, 2], context_dim=24,
             num_stage=num_stage, split_embed_dim_list=split)
    f = p["first_stage_config"]["params"]
    f["embed_dim"] = split
    f["n_embed"] = [64] * num_stage
    f["edconfig"].update(multiscale=num_stage, z_channels=split, ch=32, ch_mult=[1, 1, 2] if num_stage == 2 else [1, 1, 1, 2],
                         resolution=32, attn_resolutions=[8])
    f["ddconfig"].update(z_channels=3 * num_stage, ch=32, ch_mult=[1, 2], resolution=16, attn_resolutions=[8])
    return m


def prep(model, seed):
    synth.fill_module_(model, seed)
    man = synth.manifest_of(model)
    return man


def run_sampler(kind, model, S, B, shape, ctx, num_stage, seed, eta=0.0, noises=None, uc=None, scale=1.0):
    DDIM, PLMS = ref_loader.activate()
    import frido.models.diffusion.ddim as ddim_mod
    import frido.models.diffusion.plms as plms_mod

    sampler = (DDIM if kind == "ddim" else PLMS)(model)
    torch.manual_seed(seed)
    x_init = torch.randn((B,) + tuple(shape))
    torch.manual_seed(seed)
    mod = ddim_mod if kind == "ddim" else plms_mod
    orig = mod.noise_like
    if noises is not None:
        it = iter(noises)
        mod.noise_like = lambda shp, dev, rep=False: next(it)[:, : shp[1]].clone()
    try:
        out, inter = sampler.sample(S, B, shape, conditioning=ctx, num_stage=num_stage, eta=eta, verbose=False,
                                    log_every_t=1, unconditional_guidance_scale=scale,
                                    unconditional_conditioning=uc)
    finally:
        mod.noise_like = orig
    return x_init, out, inter


def case_tiny(base, num_stage, tag):
    cfg = tiny_cfg(base, num_stage)
    model = ref_loader.build_model(cfg)
    man = prep(model, seed=3)
    model.scale_factor.copy_(torch.tensor([0.8, 1.3, 1.1][:num_stage]))
    split = [3] * num_stage
    C = 3 * num_stage
    B = 2
    ctx = synth.synth_input("ctx", (B, 5, 24), 1)
    uc = synth.synth_input("uc", (B, 5, 24), 2)
    g = dict(cfg=cfg, manifest=man, seed=3, split=split, scale_factor=model.scale_factor.clone(), B=B)
    unet = model.model.diffusion_model
    # single evals
    for s in range(num_stage):
        x = synth.synth_input(f"x{s}", (B, 3 * (s + 1), 8, 8), 4)
        for t in (996, 1):
            ts = torch.full((B,), t, dtype=torch.long)
            g[f"eps_s{s}_t{t}"] = model.apply_model(x, ts, ctx, stage=s).clone()
    # samplers
    xi, out, inter = run_sampler("ddim", model, 4, B, (C, 8, 8), ctx, num_stage, seed=11)
    g["ddim4_xinit"], g["ddim4_out"] = xi, out
    g["ddim4_xinter1"], g["ddim4_predx0_1"] = inter["x_inter"][1].clone(), inter["pred_x0"][1].clone()
    noises = [synth.synth_input(f"nz{k}", (B, C, 8, 8), 5) for k in range(4 * num_stage)]
    xi, out, _ = run_sampler("ddim", model, 4, B, (C, 8, 8), ctx, num_stage, seed=12, eta=0.7, noises=noises)
    g["ddim4e_xinit"], g["ddim4e_out"] = xi, out
    xi, out, _ = run_sampler("plms", model, 5, B, (C, 8, 8), ctx, num_stage, seed=13)
    g["plms5_xinit"], g["plms5_out"] = xi, out
    xi, out, _ = run_sampler("ddim", model, 2, B, (C, 8, 8), ctx, num_stage, seed=14, uc=uc, scale=1.5)
    g["cfg2_xinit"], g["cfg2_out"] = xi, out
    xi, out, _ = run_sampler("plms", model, 4, B, (C, 8, 8), ctx, num_stage, seed=15, uc=uc, scale=1.5)
    g["plmscfg4_xinit"], g["plmscfg4_out"] = xi, out
    # decode (a12-a15) on the DDIM output and on a raw latent
    z = g["ddim4_out"]
    img, code = model.decode_first_stage(z, return_code=True)
    g["dec_img"] = img.clone()
    g["dec_codes"] = [torch.tensor(c, dtype=torch.int64) for c in code]
    z2 = synth.synth_input("zdec", (B, C, 4, 4), 6, scale=1.5)
    img, code = model.decode_first_stage(z2, return_code=True)
    g["dec2_z"], g["dec2_img"] = z2, img.clone()
    g["dec2_codes"] = [torch.tensor(c, dtype=torch.int64) for c in code]
    torch.save(g, os.path.join(OUT, f"{tag}.pt"))
    print(tag, "done", {k: tuple(v.shape) for k, v in g.items() if torch.is_tensor(v)})


def case_unet_l2i32(base):
    m = copy.deepcopy(base["model"])
    m["params"]["image_size"] = 32
    m["params"]["unet_config"]["params"]["image_size"] = 32
    model = ref_loader.build_model(m)
    unet = model.model.diffusion_model
    man = synth.manifest_of(unet, "model.diffusion_model.")
    synth.fill_module_(unet, 0, "model.diffusion_model.")
    ctx = synth.synth_input("ctx", (1, 26, 640), 1)
    g = dict(manifest=man, seed=0, split=[3, 3], unet_cfg=m["params"]["unet_config"]["params"])
    for s in (0, 1):
        x = synth.synth_input(f"x{s}", (1, 3 * (s + 1), 32, 32), 2)
        for t in (996, 1):
            ts = torch.full((1,), t, dtype=torch.long)
            g[f"eps_s{s}_t{t}"] = model.apply_model(x, ts, ctx, stage=s).clone()
    # one full config-1 step: DDIM-200, index 199 (t=996), eta 0
    DDIM, _ = ref_loader.activate()
    smp = DDIM(model)
    smp.make_schedule(200, ddim_eta=0.0, verbose=False)
    smp.num_stage = 2
    for s in (0, 1):
        x = synth.synth_input(f"x{s}", (1, 3 * (s + 1), 32, 32), 2)
        ts = torch.full((1,), 996, dtype=torch.long)
        xp, p0 = smp.p_sample_ddim(x, ctx, ts, s, index=199)
        g[f"step_s{s}_xprev"], g[f"step_s{s}_predx0"] = xp.clone(), p0.clone()
    # full-size decoder (f8f4) on a 16x16 latent
    fs = model.first_stage_model
    dman = [(n, s) for n, s in synth.manifest_of(fs, "first_stage_model.")
            if ".encoder." not in n and ".shared_" not in n and ".upsample." not in n.split("decoder")[0]
            and ".ms_quant_conv." not in n and ".loss." not in n]
    synth.fill_module_(fs, 0, "first_stage_model.")
    model.scale_factor.copy_(torch.tensor([0.8, 1.3]))
    z = synth.synth_input("zdec", (1, 6, 16, 16), 3, scale=1.5)
    img, code = model.decode_first_stage(z, return_code=True)
    g["dec_manifest"] = dman
    g["dec_z"], g["dec_img"] = z, img.clone()
    g["dec_codes"] = [torch.tensor(c, dtype=torch.int64) for c in code]
    g["scale_factor"] = model.scale_factor.clone()
    torch.save(g, os.path.join(OUT, "l2i32.pt"))
    print("l2i32 done")


def case_sched(base):
    m = copy.deepcopy(base["model"])
    DDIM, _ = ref_loader.activate()

    class M:  # just the attributes make_schedule reads (ddim.py:28-34)
        pass

    from frido.modules.diffusionmodules.util import make_beta_schedule

    betas = make_beta_schedule("linear", 1000, linear_start=0.0015, linear_end=0.0155)
    acp = np.cumprod(1.0 - betas, axis=0)
    mm = M()
    mm.num_timesteps = 1000
    mm.device = torch.device("cpu")
    mm.betas = torch.tensor(betas, dtype=torch.float32)
    mm.alphas_cumprod = torch.tensor(acp, dtype=torch.float32)
    mm.alphas_cumprod_prev = torch.tensor(np.append(1.0, acp[:-1]), dtype=torch.float32)
    g = dict(alphas_cumprod=mm.alphas_cumprod.clone())
    for S in (200, 250, 100, 50, 4):
        for eta in (0.0, 1.0):
            smp = DDIM(mm)
            smp.make_schedule(S, ddim_eta=eta, verbose=False)
            b = 1
            rows = []
            for index in range(len(smp.ddim_timesteps)):
                # exactly the scalars p_sample_ddim materialises (ddim.py:237-240)
                rows.append([
                    float(torch.full((b,), smp.ddim_alphas[index])[0]),
                    float(torch.full((b,), smp.ddim_alphas_prev[index])[0]),
                    float(torch.full((b,), smp.ddim_sigmas[index])[0]),
                    float(torch.full((b,), smp.ddim_sqrt_one_minus_alphas[index])[0]),
                ])
            g[f"S{S}_eta{eta}"] = dict(timesteps=torch.tensor(np.asarray(smp.ddim_timesteps)),
                                       table=torch.tensor(rows, dtype=torch.float32))
    torch.save(g, os.path.join(OUT, "sched.pt"))
    print("sched done")


def _encode_ref(model, x):
    """Reference encode_first_stage -> get_first_stage_encoding, plus the code indices each scale's quantiser picked."""
    fs = model.first_stage_model
    picked, hooks = [], []
    for q in fs.ms_quantize:
        hooks.append(q.register_forward_hook(lambda m, i, o: picked.append(o[2][2].reshape(-1).clone())))
    try:
        h = model.encode_first_stage(x)
    finally:
        for hk in hooks:
            hk.remove()
    n = len(fs.ms_quantize)
    z = model.get_first_stage_encoding(h.clone())
    return h.clone(), z.clone(), picked[-n:]  # encode_first_stage runs encode twice (frido.py:1000-1006); keep the last pass


def case_encode(base):
    """8f.3: MS-VQGAN encode side.  tiny2 / tiny3 (same configs, seeds and weights as tiny2.pt / tiny3.pt) on a
    [2,3,32,32] image, and the full-size f8f4 first stage of config 2 on a [1,3,64,64] image."""
    g = {}
    for ns, tag in ((2, "tiny2"), (3, "tiny3")):
        cfg = tiny_cfg(base, ns)
        model = ref_loader.build_model(cfg)
        prep(model, seed=3)
        model.scale_factor.copy_(torch.tensor([0.8, 1.3, 1.1][:ns]))
        x = synth.synth_input("img", (2, 3, 32, 32), 7)
        h, z, codes = _encode_ref(model, x)
        g[tag] = dict(x=x, h=h, z=z, codes=codes)
        # get_input on a 'b h w c' batch dict (frido.py:766-817), cond_stage "__is_unconditional__" + crossattn key
        print(tag, "encode", tuple(h.shape), [tuple(c.shape) for c in codes])
    m = copy.deepcopy(base["model"])
    model = ref_loader.build_model(m)
    fs = model.first_stage_model
    man = [(n, s) for n, s in synth.manifest_of(fs, "first_stage_model.") if ".loss." not in n]
    synth.fill_module_(fs, 0, "first_stage_model.")
    model.scale_factor.copy_(torch.tensor([0.8, 1.3]))
    x = synth.synth_input("img", (1, 3, 64, 64), 8)
    h, z, codes = _encode_ref(model, x)
    g["l2i"] = dict(x=x, h=h, z=z, codes=codes, manifest=man, seed=0, scale_factor=model.scale_factor.clone(),
                    fs_params={k: v for k, v in m["params"]["first_stage_config"]["params"].items() if k != "lossconfig"})
    print("l2i encode", tuple(h.shape))
    torch.save(g, os.path.join(OUT, "enc.pt"))


def case_bert(base):
    """BERTEmbedder of the layout2img config (32 layers, 640-d, vocab 30522, max_seq_len 96), tokens [2, 26]."""
    ref_loader.activate()
    from frido.modules.encoders.modules import BERTEmbedder
    g = {}
    for tag, kw, Lseq in (("full", dict(n_embed=640, n_layer=32, max_seq_len=96, use_tokenizer=False, device="cpu"), 26),
                          ("small", dict(n_embed=64, n_layer=2, vocab_size=100, max_seq_len=16, use_tokenizer=False, device="cpu"), 7)):
        enc = BERTEmbedder(**kw).eval()
        man = synth.manifest_of(enc, "cond_stage_model.")
        synth.fill_module_(enc, 7, "cond_stage_model.")
        vocab = kw.get("vocab_size", 30522)
        tokens = torch.from_numpy(np.random.default_rng(3).integers(0, min(vocab, 1024), size=(2, Lseq)))
        z = enc.encode(tokens)
        g[tag] = dict(kwargs=kw, manifest=man, seed=7, tokens=tokens, z=z.clone())
    torch.save(g, os.path.join(OUT, "bert.pt"))
    print("bert done", {k: tuple(v["z"].shape) for k, v in g.items()})


def case_mask(base):
    """8f.3: the samplers' mask / x0 branch (ddim.py:158-161, plms.py:162-165) on the tiny 2-scale model of tiny2.pt
    (same config, seed and weights).  With split heads the blend only type-checks when `img` carries all channels, i.e.
    with `x_T` given (stage 0 is then skipped, ddim.py:150-152) - that is the case minted here.  q_sample's
    torch.randn_like draws are replaced by a pre-drawn sequence (injected on the checker side as well)."""
    cfg = tiny_cfg(base, 2)
    model = ref_loader.build_model(cfg)
    prep(model, seed=3)
    model.scale_factor.copy_(torch.tensor([0.8, 1.3]))
    DDIM, PLMS = ref_loader.activate()
    B, C = 2, 6
    ctx = synth.synth_input("ctx", (B, 5, 24), 1)
    uc = synth.synth_input("uc", (B, 5, 24), 2)
    x_T = synth.synth_input("mask_xT", (B, C, 8, 8), 21)
    x0 = synth.synth_input("mask_x0", (B, C, 8, 8), 22)
    mask = (synth.synth_input("mask_m", (B, 1, 8, 8), 23) > 0).float()
    g = dict(x_T=x_T, x0=x0, mask=mask)
    for tag, cls, S, kw in (("ddim4", DDIM, 4, {}), ("plms4", PLMS, 4, {}),
                            ("ddimcfg2", DDIM, 2, dict(unconditional_guidance_scale=1.5, unconditional_conditioning=uc))):
        nzs = [synth.synth_input(f"mask_nz_{tag}{k}", (B, C, 8, 8), 24) for k in range(S)]
        it = iter(nzs)
        orig = torch.randn_like
        torch.randn_like = lambda t, *a, **k: next(it).clone()
        try:
            out, inter = cls(model).sample(S, B, (C, 8, 8), conditioning=ctx, num_stage=2, eta=0.0, verbose=False, log_every_t=1,
                                           x_T=x_T, mask=mask, x0=x0, **kw)
        finally:
            torch.randn_like = orig
        g[tag + "_out"] = out.clone()
        g[tag + "_noises"] = nzs
        g[tag + "_xinter1"] = inter["x_inter"][1].clone()
    torch.save(g, os.path.join(OUT, "mask.pt"))
    print("mask done", {k: tuple(v.shape) for k, v in g.items() if torch.is_tensor(v)})


def main():
    os.makedirs(OUT, exist_ok=True)
    base = ref_loader.load_config("configs/frido/layout2i/frido_f8f4_coco_seg.yaml")
    which = sys.argv[1:] or ["sched", "tiny2", "tiny3", "l2i32", "bert", "enc", "mask"]
    if "sched" in which:
        case_sched(base)
    if "tiny2" in which:
        case_tiny(base, 2, "tiny2")
    if "tiny3" in which:
        case_tiny(base, 3, "tiny3")
    if "l2i32" in which:
        case_unet_l2i32(base)
    if "bert" in which:
        case_bert(base)
    if "enc" in which:
        case_encode(base)
    if "mask" in which:
        case_mask(base)


if __name__ == "__main__":
    main()
