import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from oracle import torch_oracle as O
from frido_b200.program import Program
dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(12)
B, H, W, start, end = 2, 4, 4, 0, 3
acp = O.alphas_cumprod().astype(np.float32)
sch = O.ddim_schedule(10, 0.0, acp)
T = len(sch["timesteps"])
coef = torch.from_numpy(np.stack([sch["a_t"], sch["a_prev"], sch["sigma"], sch["sqrt_1m"]], 1)[::-1].copy()).to(dev)
x = torch.randn(B, end, H, W, generator=g)
e_t = torch.randn(B, end, H, W, generator=g); e_n = torch.randn(B, end, H, W, generator=g)
index = T-1
e_p = (e_t + e_n)/2
xr, p0r = O.ddim_update(x, e_p, sch, index, start)
xa, _ = O.ddim_update(x, e_t, sch, index, start)
step = torch.zeros(1, dtype=torch.int32, device=dev)
hist = torch.zeros(3, B, end, H, W, device=dev); save = torch.zeros(B, end, H, W, device=dev)
xd = x.to(dev).clone(); x_orig = xd.clone(); p0 = torch.zeros_like(xd)
P = Program(dev, 'a')
P.update(x_orig, e_t.to(dev), coef, step, xd, B=B, c_start=start, c_end=end, HW=H*W, plms_order=4, plms_mode=1, advance=0, hist=hist, eps_save=save)
P.run(); torch.cuda.synchronize()
print('mode1 diff', (xd.cpu()-xa).abs().max().item(), 'save ok', torch.equal(save.cpu(), e_t), 'step', step.item())
P2 = Program(dev, 'b')
P2.update(x_orig, e_n.to(dev), coef, step, xd, B=B, c_start=start, c_end=end, HW=H*W, plms_order=4, plms_mode=2, advance=1, hist=hist, eps_save=save, pred_x0=p0)
P2.run(); torch.cuda.synchronize()
print('mode2 diff', (xd.cpu()-xr).abs().max().item(), 'p0 diff', (p0.cpu()-p0r).abs().max().item(), 'hist ok', torch.equal(hist[0].cpu(), e_t), 'step', step.item())
