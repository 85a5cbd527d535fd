"""Times MS-VQGAN decode_first_stage on the bench config (batch 16) and lists its ops by time (eager CUDA events)."""
import sys, os, ctypes as C, collections
sys.path.insert(0, '/root/repo')
import torch
torch.set_grad_enabled(False)
import frido_b200 as fb
from frido_b200 import configs, _lib as L
dev = torch.device('cuda:0')
model, cfg = configs.build('l2i_coco', dev)
B = int(os.environ.get('PB', '16'))
z = torch.randn(B, 6, 64, 64, device=dev)
for _ in range(2):
    img = model.decode_first_stage(z)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    img = model.decode_first_stage(z)
b.record(); torch.cuda.synchronize()
print(f"decode_first_stage B{B}: {a.elapsed_time(b)/3:.2f} ms per batch")
plan = next(iter(model.first_stage_model._plans.values()))
ops, tags = plan.prog.ops, plan.prog.tags
lib = L.lib(); stream = torch.cuda.current_stream(); sptr = C.c_void_p(stream.cuda_stream)
best = {}
for rep in range(2):
    evs = []
    for i, op in enumerate(ops):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); L.check(lib.frido_run_program(C.byref(op), 1, sptr), 'op'); e1.record(stream)
        evs.append((e0, e1))
    torch.cuda.synchronize()
    for i, (e0, e1) in enumerate(evs):
        best[i] = min(best.get(i, 1e9), e0.elapsed_time(e1))
rows = []
for i, op in enumerate(ops):
    t = best[i]
    if op.kind == L.OP_CONV:
        c = op.u.conv
        fl = 2 * c.B * c.Hout * c.Wout * c.Cout * (c.ksize * c.ksize * (c.c0 + c.c1) + c.cx0 + c.cx1)
        rows.append((t, f"{tags[i]:18s} eng{c.engine} B{c.B} {c.Hout}x{c.Wout} cin{c.c0}+{c.c1} cout{c.Cout} k{c.ksize} wsb{int(c.w_sb != 0)} {t*1e3:8.1f} us {fl/t/1e9:8.1f} TF/s"))
    else:
        rows.append((t, f"{tags[i]:18s} kind{op.kind} {t*1e3:8.1f} us"))
print(f"eager per-op sum {sum(best.values()):.2f} ms over {len(ops)} ops")
for t, r in sorted(rows, reverse=True)[:45]:
    print(r)
