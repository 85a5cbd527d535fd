import csv, json, collections, sys
ops = json.load(open('gpurun_out/step_ops.json'))
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
rows = [r for r in csv.DictReader(lines) if r.get('Metric Name') == 'gpu__time_duration.sum']
def ns(r):
    v = float(r['Metric Value'].replace(',', '')); u = r['Metric Unit']
    return v * 1e3 if u == 'us' else (v * 1e6 if u == 'ms' else v)
rows = [r for r in rows if 'Memset' not in r['Kernel Name']]
rows = rows[-len(ops):]
assert len(rows) == len(ops), (len(rows), len(ops))
agg = collections.defaultdict(lambda: [0.0, 0.0, 0])
detail = []
for o, r in zip(ops, rows):
    t = ns(r) / 1e3  # us
    key = o['tag'] + (f" e{o['engine']}" if 'engine' in o else '')
    if o['kind'] == 2: key = 'gn_stats'
    agg[key][0] += t; agg[key][1] += o.get('flops', 0); agg[key][2] += 1
    if 'flops' in o: detail.append((t, o))
tot = sum(v[0] for v in agg.values())
print(f'total {tot/1e3:.2f} ms over {len(ops)} kernels (ncu, cold cache, serialised)')
for k, (t, fl, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f'{k:24s} n={n:3d} {t/1e3:8.3f} ms {100*t/tot:5.1f}%  {fl/t/1e6 if t else 0:8.1f} TF/s')
print('--- convs by time')
for t, o in sorted(detail, key=lambda x: -x[0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{o['tag']:14s} e{o['engine']} {o['shape']:44s} {t:8.1f} us {o['flops']/t/1e6:7.1f} TF/s")
