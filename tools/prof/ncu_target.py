"""Profiling target: 2 eager UNet steps (stage 1, B=16, L2I 64x64) + one decode; used under ncu only."""
import sys, os
sys.path.insert(0, '/root/repo')
import torch
torch.set_grad_enabled(False)
from frido_b200 import configs
dev = torch.device('cuda:0')
model, cfg = configs.build('l2i_coco', dev)
B = int(os.environ.get('PB', '16'))
unet = model.model.diffusion_model
plan = unet.plan(1, B, 64, 64, 26)
plan.x_in.normal_(); plan.ctx.normal_(); plan.ts.fill_(501)
plan.prologue.run()
for _ in range(int(os.environ.get('PSTEPS', '2'))):
    plan.step.run()
if os.environ.get('PDEC', '1') == '1':
    model.decode_first_stage(torch.randn(B, 6, 64, 64, device=dev))
torch.cuda.synchronize()
print('done')
