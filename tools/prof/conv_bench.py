"""Micro-benchmark of single conv2d_tc launches (CUDA events), optionally with a forced tile width.
   python tools/prof/conv_bench.py [shape-index ...]      env: CB_BNS="0,64,192" CB_FLUSH=1"""
import sys, os, ctypes as C
sys.path.insert(0, '/root/repo')
import torch
torch.set_grad_enabled(False)
from frido_b200 import _lib as L
from frido_b200.program import Program, Src
dev = torch.device('cuda:0')
SHAPES = [  # B, H, W, cin, cout, k
    (16, 8, 8, 1920, 960, 3), (16, 8, 8, 960, 960, 3), (16, 8, 8, 960, 960, 1), (16, 8, 8, 960, 7680, 1),
    (16, 16, 16, 1536, 576, 3), (16, 16, 16, 576, 576, 3), (16, 16, 16, 576, 576, 1),
    (16, 32, 32, 384, 384, 3), (16, 64, 64, 192, 192, 3), (16, 64, 64, 576, 192, 3),
]
sel = [int(a) for a in sys.argv[1:]] or range(len(SHAPES))
bns = [int(b) for b in os.environ.get('CB_BNS', '0').split(',')]
flush = os.environ.get('CB_FLUSH', '1') == '1'
junk = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
stream = torch.cuda.current_stream()
for si in sel:
    B, H, W, cin, cout, k = SHAPES[si]
    x = torch.randn(B, H * W, cin, device=dev)
    w = torch.randn(cout, k * k * cin, device=dev) * 0.02
    out = torch.empty(B, H * W, cout, device=dev)
    fl = 2 * B * H * W * cout * k * k * cin
    for bn in bns:
        if bn:
            if cout % bn:
                continue
            os.environ['FRIDO_TC_FORCE_BN'] = str(bn)
        else:
            os.environ.pop('FRIDO_TC_FORCE_BN', None)
        P = Program(dev, 'cb')
        P.conv(Src.nhwc(x, H, W), w, out, B=B, Hin=H, Win=W, Hout=H, Wout=W, Cout=cout, ksize=k, pad=k // 2)
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
        print(f"pair={os.environ.get('FRIDO_TC_PAIR', '-')} B{B} {H}x{W} cin{cin} cout{cout} k{k} BN={bn or 'auto':>4}  {t:8.1f} us  {fl / t / 1e6:7.1f} TF/s  (min {ts[0]:.1f})", flush=True)
