"""Micro-benchmark of linear-shaped conv2d_tc launches (CUDA events): separates per-k-step cost from per-tile (epilogue) cost.
   python tools/prof/lin_bench.py      env: LB_BNS="0,64,128,192" LB_FLUSH=0"""
import sys, os
sys.path.insert(0, '/root/repo')
import torch
torch.set_grad_enabled(False)
from frido_b200 import _lib as L
from frido_b200.program import Program
dev = torch.device('cuda:0')
SHAPES = [  # M, K, N, act, res
    (16384, 384, 3072, 'none', 0), (16384, 384, 3072, 'geglu', 0), (16384, 1536, 3072, 'none', 0),
    (16384, 1536, 384, 'none', 1), (16384, 384, 384, 'none', 1), (16384, 384, 384, 'none', 0), (16384, 384, 768, 'none', 0),
    (4096, 576, 4608, 'geglu', 0), (4096, 2304, 576, 'none', 1), (1024, 960, 7680, 'geglu', 0), (1024, 3840, 960, 'none', 1),
    (1024, 960, 960, 'none', 1),
]
if os.environ.get('LB_SEL'):
    SHAPES = [SHAPES[int(i)] for i in os.environ['LB_SEL'].split(',')]
bns = [int(b) for b in os.environ.get('LB_BNS', '0').split(',')]
flush = os.environ.get('LB_FLUSH', '0') == '1'
junk = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
stream = torch.cuda.current_stream()
for (M, K, N, act, res) in SHAPES:
    x = torch.randn(M, K, device=dev)
    w = torch.randn(N, K, device=dev) * 0.02
    n_out = N // 2 if act == 'geglu' else N
    out = torch.empty(M, n_out, device=dev)
    r = torch.randn(M, n_out, device=dev) if res else None
    bias = torch.randn(N, device=dev)
    fl = 2 * M * K * N
    for bn in bns:
        if bn:
            if N % bn:
                continue
            os.environ['FRIDO_TC_FORCE_BN'] = str(bn)
        else:
            os.environ.pop('FRIDO_TC_FORCE_BN', None)
        P = Program(dev, 'lb')
        P.linear(x, w, out, M=M, K=K, N=N, bias=bias, res=r, act=L.ACT_GEGLU if act == 'geglu' else L.ACT_NONE)
        P.prepare_weights()
        for _ in range(3):
            P.run()
        ts = []
        for _ in range(8):
            if flush:
                junk.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); P.run(); b.record(stream)
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        ts.sort()
        t = ts[len(ts) // 2]
        tiles = ((M + 127) // 128) * (N // (bn or 192)) if (bn or N % 192 == 0) else 0
        print(f"M{M} K{K} N{N} {act:5s} res{res} BN={bn or 'auto':>4}  {t:8.1f} us  {fl / t / 1e6:7.1f} TF/s  (min {ts[0]:.1f})", flush=True)
