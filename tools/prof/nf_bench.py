"""Micro-benchmark of the normalise-on-load conv (csrc/conv_nf.cu) against the two-launch path it replaces
(frido_norm_act + conv2d_tc), CUDA events, L2 flushed between repetitions.
   python tools/prof/nf_bench.py [shape-index ...]      env: NB_SPADE=0/1  NB_ONLY=nf|base"""
import sys, os
sys.path.insert(0, '/root/repo')
import torch
torch.set_grad_enabled(False)
from frido_b200 import _lib as L
from frido_b200.program import Program, Src
dev = torch.device('cuda:0')
SHAPES = [  # B, H, W, c0, c1, cout, side
    (16, 64, 64, 384, 192, 192, 0), (16, 64, 64, 192, 0, 192, 0), (16, 64, 64, 192, 0, 192, 192), (16, 32, 32, 384, 384, 384, 0),
    (16, 32, 32, 384, 0, 384, 384), (16, 16, 16, 576, 576, 576, 0), (16, 8, 8, 960, 960, 960, 0), (16, 64, 64, 192, 0, 192, 576),
]
sel = [int(a) for a in sys.argv[1:]] or range(len(SHAPES))
spade = os.environ.get('NB_SPADE', '0') == '1'
only = os.environ.get('NB_ONLY', '')
junk = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
stream = torch.cuda.current_stream()


def timeit(P, n=6):
    for _ in range(2):
        P.run()
    ts = []
    for _ in range(n):
        junk.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); P.run(); b.record(stream)
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


for si in sel:
    B, H, W, c0, c1, cout, side = SHAPES[si]
    C = c0 + c1
    x0 = torch.randn(B, H * W, c0, device=dev)
    x1 = torch.randn(B, H * W, c1, device=dev) if c1 else None
    w = torch.randn(cout, 9 * C + side, device=dev) * 0.02
    gw, gb_ = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    gbm = torch.randn(B, H * W, 2 * C, device=dev) * 0.1 if spade else None
    out = torch.empty(B, H * W, cout, device=dev)
    cs0 = torch.zeros(B, c0, 2, dtype=torch.float64, device=dev)
    cs1 = torch.zeros(B, c1, 2, dtype=torch.float64, device=dev) if c1 else None
    S = Program(dev, 'stats')
    S.chan_stats(x0, c0, cs0, B=B, HW=H * W)
    if c1:
        S.chan_stats(x1, c1, cs1, B=B, HW=H * W)
    S.run()
    ab = torch.zeros(B, C, 2, device=dev)
    a0 = Src.nhwc(x0, H, W)
    a1 = Src.nhwc(x1, H, W) if c1 else None
    sx = torch.randn(B, H * W, side, device=dev) if side else None
    sd = (Src.nhwc(sx, H, W), None) if side else None
    fl = 2 * B * H * W * cout * (9 * C + side)
    kw = dict(B=B, Hin=H, Win=W, Hout=H, Wout=W, Cout=cout, ksize=3, pad=1, engine=3)
    if only != 'base':
        P = Program(dev, 'nf')
        P.gn_finalize(ab, gw, gb_, B=B, HW=H * W, c0=c0, c1=c1, eps=1e-5, csum0=cs0, csum1=cs1)
        P.conv(a0, w, out, a1=a1, side=sd, nrm=(ab, gbm, 1), **kw)
        P.prepare_weights()
        t, tmin = timeit(P)
        print(f"NF    B{B} {H}x{W} c{c0}+{c1} cout{cout} side{side} spade{int(spade)}  {t:8.1f} us  {fl / t / 1e6:7.1f} TF/s  (min {tmin:.1f})  dbg={os.environ.get('FRIDO_NF_DBG', '0')}", flush=True)
    if only != 'nf':
        P = Program(dev, 'base')
        tt = torch.empty(B, H * W, C, device=dev)
        P.norm_act(x0, c0, None, gw, gb_, tt, B=B, HW=H * W, eps=1e-5, a1=x1, c1=c1, gb=gbm, silu=1, csum0=cs0, csum1=cs1)
        P.conv(Src.nhwc(tt, H, W), w, out, side=sd, **kw)
        P.prepare_weights()
        t, tmin = timeit(P)
        print(f"BASE  B{B} {H}x{W} c{c0}+{c1} cout{cout} side{side} spade{int(spade)}  {t:8.1f} us  {fl / t / 1e6:7.1f} TF/s  (min {tmin:.1f})", flush=True)
