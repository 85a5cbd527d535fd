import sys, os, copy, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
torch.set_grad_enabled(False)
from oracle import synth
dev = torch.device('cuda:0')
def build(g, eng):
    os.environ['FRIDO_ENGINE'] = eng
    import frido_b200 as fb
    p = copy.deepcopy(g["cfg"]["params"]); p["cond_stage_config"] = "__is_unconditional__"; p["use_ema"] = False
    p["first_stage_config"]["params"]["ckpt_path"] = None
    m = fb.FridoDiffusion(**p); synth.fill_module_(m, g["seed"]); m.scale_factor.copy_(g["scale_factor"]); m = m.to(dev)
    return m, fb
for eng in ('bf16x3', 'tc3'):
    for tag in ('tiny2', 'tiny3'):
        g = torch.load(f'tests/golden/{tag}.pt', weights_only=False)
        m, fb = build(g, eng)
        split, B = g["split"], g["B"]; ns, C = len(split), sum(split)
        ctx = synth.synth_input("ctx", (B, 5, 24), 1).to(dev); uc = synth.synth_input("uc", (B, 5, 24), 2).to(dev)
        r = {}
        for s in range(ns):
            x = synth.synth_input(f"x{s}", (B, 3 * (s + 1), 8, 8), 4).to(dev)
            for t in (996, 1):
                e = m.apply_model(x, torch.full((B,), t, dtype=torch.long, device=dev), ctx, stage=s)
                r[f'eps{s}_{t}'] = (e.cpu() - g[f"eps_s{s}_t{t}"]).abs().max().item()
        smp = fb.DDIMSampler(m)
        out, inter = smp.sample(4, B, (C, 8, 8), conditioning=ctx, num_stage=ns, eta=0.0, verbose=False, log_every_t=1, init_noise=g["ddim4_xinit"].to(dev))
        r['ddim4'] = (out.cpu() - g["ddim4_out"]).abs().max().item()
        r['xinter1'] = (inter["x_inter"][1].cpu() - g["ddim4_xinter1"]).abs().max().item()
        out, _ = smp.sample(2, B, (C, 8, 8), conditioning=ctx, num_stage=ns, eta=0.0, verbose=False, init_noise=g["cfg2_xinit"].to(dev), unconditional_guidance_scale=1.5, unconditional_conditioning=uc)
        r['cfg2'] = (out.cpu() - g["cfg2_out"]).abs().max().item()
        pl = fb.PLMSSampler(m)
        out, _ = pl.sample(5, B, (C, 8, 8), conditioning=ctx, num_stage=ns, eta=0.0, verbose=False, init_noise=g["plms5_xinit"].to(dev))
        r['plms5'] = (out.cpu() - g["plms5_out"]).abs().max().item()
        img, codes = m.decode_first_stage(g["dec2_z"].to(dev), return_code=True)
        r['dec2'] = (img.cpu() - g["dec2_img"]).abs().max().item()
        r['codes'] = all(torch.equal(torch.tensor(a), b) for a, b in zip(codes, g["dec2_codes"]))
        r['img_absmax'] = g["dec2_img"].abs().max().item()
        print(eng, tag, {k: (f'{v:.2e}' if isinstance(v, float) else v) for k, v in r.items()})
    g = torch.load('tests/golden/l2i32.pt', weights_only=False)
    os.environ['FRIDO_ENGINE'] = eng
    import frido_b200 as fb
    unet = fb.PyUNetModel(**g["unet_cfg"]); synth.fill_module_(unet, g["seed"], "model.diffusion_model."); unet = unet.to(dev)
    ctx = synth.synth_input("ctx", (1, 26, 640), 1).to(dev)
    r = {}
    for s in (0, 1):
        x = synth.synth_input(f"x{s}", (1, 3 * (s + 1), 32, 32), 2).to(dev)
        for t in (996, 1):
            e = unet(x, torch.full((1,), t, dtype=torch.long, device=dev), context=ctx, stage=s)
            d = (e.cpu() - g[f"eps_s{s}_t{t}"])
            r[f'eps{s}_{t}'] = f'{d.abs().max().item():.2e}/rms{d.pow(2).mean().sqrt().item():.1e}/std{g[f"eps_s{s}_t{t}"].std().item():.2f}'
    print(eng, 'l2i32', r)
    del unet
    dd = dict(double_z=False, z_channels=6, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2, 4], num_res_blocks=2, attn_resolutions=[64], dropout=0.0)
    fs = fb.VQModelInterface(embed_dim=[3, 3], n_embed=[4096, 4096], ddconfig=dd, edconfig=None, init_normal=True)
    synth.fill_module_(fs, g["seed"], "first_stage_model."); fs = fs.to(dev)
    img, codes = fs.decode(g["dec_z"].to(dev), return_code=True, scale_factor=g["scale_factor"].tolist())
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(3): fs.decode(g["dec_z"].to(dev), scale_factor=g["scale_factor"].tolist())
    torch.cuda.synchronize()
    d = img.cpu() - g["dec_img"]
    print(eng, 'decoder', f'{d.abs().max().item():.2e} rms {d.pow(2).mean().sqrt().item():.1e} img absmax {g["dec_img"].abs().max().item():.2f} std {g["dec_img"].std().item():.2f}', 'time/dec(1 img 16x16 latent)', (time.time()-t0)/3)
