"""Where a 16-image batch goes: whole sampler.sample() + decode vs. the pure graph replays (bench config)."""
import sys, os, time
sys.path.insert(0, '/root/repo')
import torch
torch.set_grad_enabled(False)
import frido_b200 as fb
from frido_b200 import configs
dev = torch.device('cuda:0')
model, cfg = configs.build('l2i_coco', dev)
B = cfg['batch']; C, H, W = cfg['latent']; Lc, D = cfg['ctx']; S = cfg['steps']
sampler = fb.DDIMSampler(model)
ctx = torch.randn(B, Lc, D, device=dev); x0 = torch.randn(B, C, H, W, device=dev)
def sample():
    z, _ = sampler.sample(S, B, (C, H, W), conditioning=ctx, num_stage=2, eta=0.0, verbose=False, log_every_t=10**9, init_noise=x0)
    return z
def timed(fn, n=1):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        r = fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3 / n, r
sample(); z = sample()
t_sample, z = timed(sample, 2)
t_dec, _ = timed(lambda: model.decode_first_stage(z), 3)
unet = model.model.diffusion_model
t_fp, _ = timed(unet._weights_fingerprint, 3)
def force_repack():
    unet.invalidate()
    for p in unet._plans.values():
        p.repack_if_stale()
t_repack, _ = timed(force_repack, 1)
reps = {}
for key, st in sampler._stage_cache.items():
    st["step"].zero_()
    reps[key[0]], _ = timed(st["main"].replay, 50)   # 50 of the 200 steps of a stage: the step counter stays in range
    st["step"].zero_()
t_pro = {k[0]: timed(p.prologue.run, 3)[0] for k, p in unet._plans.items()}
print(f"sample() {t_sample:.1f} ms | decode {t_dec:.1f} ms | fingerprint {t_fp:.2f} ms | full re-pack of both stage plans {t_repack:.1f} ms")
print("UNet step graph replay per stage (ms):", {k: round(v, 3) for k, v in reps.items()}, "| prologue per stage (ms):", {k: round(v, 2) for k, v in t_pro.items()})
est = S * sum(reps.values()) + sum(t_pro.values())
print(f"200 x step graphs + prologues = {est:.1f} ms; host/other overhead inside sample() = {t_sample - est:.1f} ms")
