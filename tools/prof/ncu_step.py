"""ncu target: one eager stage-1 UNet step (after a warm-up step); dumps the op list so kernel rows can be matched by order."""
import sys, os, json
sys.path.insert(0, '/root/repo')
import torch
torch.set_grad_enabled(False)
from frido_b200 import configs, _lib as L
dev = torch.device('cuda:0')
model, cfg = configs.build('l2i_coco', dev)
B = int(os.environ.get('PB', '16'))
plan = model.model.diffusion_model.plan(int(os.environ.get('PSTAGE', '1')), B, 64, 64, 26)
plan.x_in.normal_(); plan.ctx.normal_(); plan.ts.fill_(501)
plan.prologue.run(); plan.step.run(); torch.cuda.synchronize()
ops = []
for op, tag in zip(plan.step.ops, plan.step.tags):
    if op.kind == L.OP_ZERO:
        continue
    d = dict(tag=tag, kind=op.kind)
    if op.kind == L.OP_CONV:
        c = op.u.conv
        d.update(engine=c.engine, flops=2 * c.B * c.Hout * c.Wout * c.Cout * (c.ksize * c.ksize * (c.c0 + c.c1) + c.cx0 + c.cx1),
                 shape=f"B{c.B} {c.Hout}x{c.Wout} cin{c.c0}+{c.c1} cout{c.Cout} k{c.ksize} s{c.stride}")
    elif op.kind == L.OP_FLASH:
        a = op.u.flash
        d.update(engine=3, flops=4 * a.B * a.N * a.N * a.C, shape=f"B{a.B} N{a.N} C{a.C}")
    ops.append(d)
json.dump(ops, open('/root/repo/gpurun_out/step_ops.json', 'w'))
torch.cuda.profiler.start()
plan.step.run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('done', len(ops))
