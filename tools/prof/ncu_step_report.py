"""Turns the ncu launch list of one UNet step (tools/prof/ncu_step.py under
   ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv)
   into markdown tables + profiles/traffic.json.   python tools/prof/ncu_step_report.py <launches.csv> <step_ops.json>"""
import csv, json, collections, sys, re
ops = json.load(open(sys.argv[2]))
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
recs = collections.OrderedDict()
for r in csv.DictReader(lines):
    d = recs.setdefault(r['ID'], dict(name=r['Kernel Name']))
    v = float(r['Metric Value'].replace(',', '')); u = r['Metric Unit']
    if r['Metric Name'] == 'gpu__time_duration.sum':
        d['us'] = v / 1e3 if u == 'ns' else (v if u == 'us' else v * 1e3)
    elif r['Metric Name'] in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
        scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
        d[r['Metric Name'].split('_')[-1].split('.')[0]] = v * scale
rows = [d for d in recs.values() if 'Memset' not in d['name']][-len(ops):]
assert len(rows) == len(ops), (len(rows), len(ops))
def short(n):
    n = n.split('(')[0].replace('void ', '').replace('frido::', '')
    return re.sub(r'\(int\)|\(bool\)', '', n)
kern = collections.defaultdict(lambda: [0.0, 0.0, 0, 0.0])
byop = collections.defaultdict(lambda: [0.0, 0.0, 0])
detail = []
for o, r in zip(ops, rows):
    k = short(r['name']); fam = k.split('<')[0]
    t = r['us']; fl = o.get('flops', 0); by = r.get('read', 0) + r.get('write', 0)
    kern[fam][0] += t; kern[fam][1] += fl; kern[fam][2] += 1; kern[fam][3] += by
    byop[o['tag']][0] += t; byop[o['tag']][1] += fl; byop[o['tag']][2] += 1
    if fl: detail.append((t, o, by))
tot = sum(v[0] for v in kern.values())
print(f"| kernel | launches | total us | share | algorithmic TFLOP/s | DRAM MB / launch |\n|---|---:|---:|---:|---:|---:|")
for k, (t, fl, n, by) in sorted(kern.items(), key=lambda kv: -kv[1][0]):
    print(f"| `{k}` | {n} | {t:.1f} | {100*t/tot:.1f}% | {fl/t/1e6 if t else 0:.1f} | {by/n/1e6:.2f} |")
print(f"| total | {len(ops)} | {tot:.1f} | | {sum(v[1] for v in kern.values())/tot/1e6:.1f} | |\n")
print("| op | launches | total us | share | TFLOP/s |\n|---|---:|---:|---:|---:|")
for k, (t, fl, n) in sorted(byop.items(), key=lambda kv: -kv[1][0])[:26]:
    print(f"| {k} | {n} | {t:.1f} | {100*t/tot:.1f}% | {fl/t/1e6 if t else 0:.1f} |")
print("\n| op | shape | us | algorithmic TFLOP/s | DRAM MB |\n|---|---|---:|---:|---:|")
for t, o, by in sorted(detail, key=lambda x: -x[0])[:14]:
    print(f"| {o['tag']} | {o['shape']} | {t:.1f} | {o['flops']/t/1e6:.1f} | {by/1e6:.1f} |")
tc = [kern['conv_tc_kernel'][i] + kern['conv_tc_bf_kernel'][i] + kern['conv_tc_pair_kernel'][i] + kern['conv_nf_kernel'][i] for i in range(4)]
json.dump({"kernel": "frido::conv_tc_bf_kernel + conv_nf_kernel (+ conv_tc_kernel / conv_tc_pair_kernel where selected): BF16x3 tcgen05 implicit-GEMM convs", "bytes_per_launch": round(tc[3] / tc[2]),
           "launches": tc[2], "share_of_step_under_ncu": round(tc[0] / tot, 4),
           "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over one eager stage-1 UNet step (batch 16), averaged over the tcgen05 conv launches"},
          open('profiles/traffic.json', 'w'), indent=1)
print(f"\nconv_tc share under ncu: {100*tc[0]/tot:.1f}%  avg DRAM bytes per launch {tc[3]/tc[2]/1e6:.2f} MB", file=sys.stderr)
