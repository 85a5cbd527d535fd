import sys, os, ctypes as C, json, collections
sys.path.insert(0, '/root/repo')
import torch
torch.set_grad_enabled(False)
import frido_b200 as fb
from frido_b200 import configs, _lib as L
dev = torch.device('cuda:0')
model, cfg = configs.build(os.environ.get('PCONFIG', 'l2i_coco'), dev)
B = int(os.environ.get('PB', str(cfg['batch'])))
unet = model.model.diffusion_model
stage = int(os.environ.get('PSTAGE', '1'))
plan = unet.plan(stage, B, cfg['latent'][1], cfg['latent'][2], cfg['ctx'][0])
plan.prologue.run(); plan.step.run(); torch.cuda.synchronize()
lib = L.lib()
ops, tags = plan.step.ops, plan.step.tags
stream = torch.cuda.current_stream(); sptr = C.c_void_p(stream.cuda_stream)
best = {}
for rep in range(3):
    evs = []
    for i, op in enumerate(ops):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); L.check(lib.frido_run_program(C.byref(op), 1, sptr), 'op'); b.record(stream)
        evs.append((a, b))
    torch.cuda.synchronize()
    for i, (a, b) in enumerate(evs):
        t = a.elapsed_time(b)
        best[i] = min(best.get(i, 1e9), t)
rows = []
agg = collections.defaultdict(lambda: [0.0, 0.0, 0])
for i, op in enumerate(ops):
    t = best[i]
    if op.kind == L.OP_CONV:
        c = op.u.conv
        fl = 2 * c.B * c.Hout * c.Wout * c.Cout * (c.ksize * c.ksize * (c.c0 + c.c1) + c.cx0 + c.cx1)
        key = f"{tags[i]} e{c.engine}"
        rows.append((t, f"{tags[i]:16s} eng{c.engine} B{c.B} {c.Hout}x{c.Wout} cin{c.c0}+{c.c1} cout{c.Cout} k{c.ksize} s{c.stride}  {t*1e3:8.1f} us  {fl/t/1e9:8.1f} TF/s"))
    elif op.kind == L.OP_ATTN:
        a = op.u.attn
        fl = 4 * a.B * a.N * a.Nk * a.C; key = tags[i]
        rows.append((t, f"{tags[i]:16s} attn B{a.B} N{a.N} Nk{a.Nk} C{a.C} ln{int(bool(a.ln_gamma))}  {t*1e3:8.1f} us  {fl/t/1e9:8.1f} TF/s"))
    elif op.kind == L.OP_FLASH:
        a = op.u.flash
        fl = 4 * a.B * a.N * a.N * a.C; key = tags[i]
        rows.append((t, f"{tags[i]:16s} flash B{a.B} N{a.N} C{a.C}  {t*1e3:8.1f} us  {fl/t/1e9:8.1f} TF/s"))
    elif op.kind == L.OP_NORM_ACT:
        a = op.u.norm_act
        by = 4 * a.B * a.HW * (a.c0 + a.c1) * (4 if a.gb else 2)
        fl = 0; key = tags[i]
        rows.append((t, f"{tags[i]:16s} norm B{a.B} HW{a.HW} C{a.c0}+{a.c1} spade{int(bool(a.gb))}  {t*1e3:8.1f} us  {by/t/1e6:8.1f} GB/s"))
    else:
        fl = 0; key = tags[i] if op.kind != L.OP_GN_STATS else 'gn_stats'
        rows.append((t, f"{key:16s} kind{op.kind}  {t*1e3:8.1f} us"))
    agg[key][0] += t; agg[key][1] += fl; agg[key][2] += 1
tot = sum(best.values())
print(f"stage {stage} B {B}: total eager per-op sum {tot:.2f} ms over {len(ops)} ops")
for k, (t, fl, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:24s} n={n:3d} {t:8.3f} ms {100*t/tot:5.1f}%  {fl/t/1e9 if t else 0:8.1f} TF/s")
print('--- slowest ops')
for t, r in sorted(rows, reverse=True)[:(10**6 if os.environ.get('PALL') else 60)]:
    print(r)
print('--- least efficient big convs')
rows_eff = [(fl_t, r) for fl_t, r in ((float(r.split()[-2]), r) for _, r in rows) if True]
# --- the same step as one captured CUDA graph (what the sampler replays): ms per replay
plan.step.capture()
torch.cuda.synchronize()
reps = int(os.environ.get('PREPS', '30'))
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    plan.step.replay()
a.record(stream)
for _ in range(reps):
    plan.step.replay()
b.record(stream)
torch.cuda.synchronize()
print(f"GRAPH stage {stage} B {B}: {a.elapsed_time(b) / reps:.3f} ms per step replay ({len(ops)} ops; env "
      f"FRIDO_SK={os.environ.get('FRIDO_SK', '1')} FRIDO_ATTN_SMALL={os.environ.get('FRIDO_ATTN_SMALL', '1')} "
      f"FRIDO_PDL={os.environ.get('FRIDO_PDL', '-')})")
