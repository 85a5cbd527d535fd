"""GPU box: event-timed launches of the streaming-softmax attention kernel at the UNet's shapes (also the ncu target:
ncu -k regex:attn_flash ...).  Prints us per launch and algorithmic TFLOP/s (4 B N^2 C)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

torch.set_grad_enabled(False)
from frido_b200 import _lib as L

dev = torch.device("cuda:0")
lib = L.lib()
stream = torch.cuda.current_stream()
sptr = C.c_void_p(stream.cuda_stream)


def pair(t):
    hi = torch.empty(t.shape, dtype=torch.bfloat16, device=dev)
    lo = torch.empty_like(hi)
    L.check(lib.frido_split_bf16(t.data_ptr(), hi.data_ptr(), lo.data_ptr(), t.numel(), stream.cuda_stream), "split")
    return hi, lo


shapes = [(16, 1024, 384), (16, 256, 576), (8, 256, 384), (16, 4096, 384), (16, 1024, 576), (16, 256, 960)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
reps = int(os.environ.get("PREPS", "20"))
for B, N, Cd in shapes:
    g = torch.Generator().manual_seed(0)
    q = torch.randn(B, N, Cd, generator=g).to(dev)
    k = torch.randn(B, N, Cd, generator=g).to(dev)
    vt = torch.randn(Cd, B, N, generator=g).to(dev)
    res = torch.randn(B, N, Cd, generator=g).to(dev)
    bias = torch.randn(Cd, generator=g).to(dev)
    out = torch.empty(B, N, Cd, device=dev)
    qp, kp, vp = pair(q), pair(k), pair(vt)
    p = L.FlashParams()
    p.q_hi, p.q_lo, p.q_sb, p.q_ld = qp[0].data_ptr(), qp[1].data_ptr(), N * Cd, Cd
    p.k_hi, p.k_lo, p.k_sb, p.k_ld = kp[0].data_ptr(), kp[1].data_ptr(), N * Cd, Cd
    p.vt_hi, p.vt_lo, p.vt_sb, p.vt_ld = vp[0].data_ptr(), vp[1].data_ptr(), N, B * N
    p.B, p.N, p.C, p.scale = B, N, Cd, Cd ** -0.5
    p.bias, p.res, p.r_sb, p.r_ld = bias.data_ptr(), res.data_ptr(), N * Cd, Cd
    p.out, p.o_sb, p.o_ld = out.data_ptr(), N * Cd, Cd
    for _ in range(3):
        L.check(lib.frido_attn_flash(C.byref(p), sptr), "flash")
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps):
        L.check(lib.frido_attn_flash(C.byref(p), sptr), "flash")
    b.record(stream)
    torch.cuda.synchronize()
    us = a.elapsed_time(b) / reps * 1e3
    print(f"pair={os.environ.get('FRIDO_FLASH_PAIR', '-')} flash B{B} N{N} C{Cd}: {us:8.1f} us  {4 * B * N * N * Cd / us / 1e6:7.1f} TF/s algorithmic", flush=True)
