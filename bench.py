#!/usr/bin/env python
"""bench.py — Frido multi-scale denoising sampling, BASELINE.json metric.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config l2i_coco]

A "step" = one pass of the hot path over one batch: DDIM-200 over both stages
(400 UNet evals) + MS-VQGAN decode of a 16-image batch of the COCO-stuff
layout-to-image 256x256 config (BASELINE config 2), synthetic weights/inputs.

  value : images/sec, whole job, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e   : same metric through the public API (DDIMSampler.sample + decode_first_stage) with HOST
          (pinned) context/noise copied in and the decoded images copied out every step
  roofline     : the dominant kernel (tcgen05 implicit-GEMM conv), per-launch CUDA-event timing
  cpu_baseline : oracle port (oracle/torch_oracle.py = the reference algorithm in CPU PyTorch) on the
                 host cores, bounded sample, extrapolated linearly (steps are shape-homogeneous)
`--impl reference` times that same CPU implementation as the reference arm (rank 0 only).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

torch.set_grad_enabled(False)


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d.get("hbm_gbs", 6650.0), bf16=d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0)),
                    bf16_burst=d.get("bf16_tflops", 1590.0), source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, bf16=1400.0, bf16_burst=1590.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        time.sleep(0.05)
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))


# ----------------------------------------------------------------------------------------------
def _host_cores():
    """Threads the CPU arms use: every core this process may run on (torchrun exports OMP_NUM_THREADS=1, which would
    silently turn the baseline into a single-core run)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    torch.set_num_threads(n)
    return torch.get_num_threads()


def _reference_model(model, cfg, device):
    """The reference's OWN FridoDiffusion (unmodified modules from /root/reference or the archive packed by
    oracle/vendor_ref.py) carrying the same synthetic weights as `model`; None if the reference is not available."""
    try:
        from oracle import ref_loader
        if not ref_loader.available():
            return None
        on_cpu = torch.device(device).type == "cpu"
        ref_loader.activate(cpu=on_cpu)
        ref = ref_loader.build_model(cfg["model"], cpu=on_cpu)
        sd = {k: v.detach().float().cpu() for k, v in model.state_dict().items()}
        missing, unexpected = ref.load_state_dict(sd, strict=False)
        bad = [k for k in missing if ".loss." not in k and "model_ema" not in k]
        if bad:
            raise RuntimeError(f"reference model misses {len(bad)} keys, e.g. {bad[:3]}")
        return ref.to(device).eval()
    except Exception as e:  # the baseline failing must not hide the GPU number
        print(f"[bench] reference modules unavailable: {e!r}", file=sys.stderr)
        return None


def cpu_sample(model, cfg, batch, evals=3, ref=None, budget_s=40.0):
    """Bounded sample of the workload on the host cores: after one warm-up evaluation, `evals` UNet
    evaluations per stage at batch `batch` and one decode of `batch` images, extrapolated linearly to the config's
    step count (every step has the same shapes).  `ref` = the reference's own model on CPU (kind "reference"), else
    the oracle port (kind "port": oracle/torch_oracle.py, pinned against the reference's outputs in tests/)."""
    from oracle import torch_oracle as O

    cores = _host_cores()
    split = list(model.split_embed_dim_list)
    C, H, W = cfg["latent"]
    Lc, D = cfg["ctx"]
    g = torch.Generator().manual_seed(5)
    ctx = torch.randn(batch, Lc, D, generator=g)
    ts = torch.full((batch,), 501, dtype=torch.long)
    sf = [float(v) for v in model.scale_factor.cpu().tolist()]
    if ref is None:
        sd = {k: v.detach().float().cpu() for k, v in model.state_dict().items()}
        unet = lambda x, s: O.unet_forward(sd, x, ts[: x.shape[0]], ctx[: x.shape[0]], s, split)
        dec = lambda z: O.decode_first_stage(sd, z, split, sf)
    else:
        unet = lambda x, s: ref.apply_model(x, ts[: x.shape[0]], ctx[: x.shape[0]], stage=s)
        dec = lambda z: ref.decode_first_stage(z)
    t_stage, n_done = [], []
    t_begin = time.perf_counter()
    for s in range(len(split)):
        x = torch.randn(batch, sum(split[: s + 1]), H, W, generator=g)
        if s == 0:
            unet(x[:1], s)  # warm-up (allocator, oneDNN primitives, page-in of the weights)
            t_begin = time.perf_counter()
        ts_ = []
        for _ in range(evals):
            t0 = time.perf_counter()
            unet(x, s)
            ts_.append(time.perf_counter() - t0)
            if time.perf_counter() - t_begin > budget_s * (s + 1) / (len(split) + 1) and len(ts_) >= 1:
                break
        t_stage.append(statistics.median(ts_) / batch)
        n_done.append(len(ts_))
    z = torch.randn(batch, C, H, W, generator=g)
    t0 = time.perf_counter()
    dec(z)
    t_dec = (time.perf_counter() - t0) / batch
    n_evals = cfg["steps"] + (1 if cfg["sampler"] == "plms" else 0)
    sec_per_img = n_evals * sum(t_stage) + t_dec
    return dict(value=1.0 / sec_per_img, unit="images/sec", cores=cores, kind="reference" if ref is not None else "port",
                sample=f"{'/'.join(str(n) for n in n_done)} UNet evals per stage at batch {batch} (median "
                       f"{'/'.join(f'{t:.3f}' for t in t_stage)} s per image) + 1 decode of {batch} images ({t_dec:.3f} s per image), "
                       f"extrapolated to {n_evals} evals per stage; torch {torch.__version__} CPU fp32, {cores} threads; "
                       + ("the reference's own modules (oracle/_ref)" if ref is not None else "oracle port"))


def gpu_eager_reference(model, cfg, dev, steps_small=4):
    """Secondary software baseline: the reference's own modules in eager PyTorch on this
    B200 through ITS public API (DDIMSampler.sample + decode_first_stage), same weights and shapes; `steps_small`
    sampler steps per stage timed after a warm-up and extrapolated linearly to the config's step count."""
    ref = _reference_model(model, cfg, dev)
    if ref is None:
        return None
    from oracle import ref_loader
    DDIM, PLMS = ref_loader.activate(cpu=False)
    B = cfg["batch"]
    C, H, W = cfg["latent"]
    Lc, D = cfg["ctx"]
    g = torch.Generator().manual_seed(1)
    ctx = torch.randn(B, Lc, D, generator=g).to(dev)
    smp = (DDIM if cfg["sampler"] == "ddim" else PLMS)(ref)
    ns = len(model.split_embed_dim_list)
    S = 1000 // (1000 // steps_small)  # the reference's schedule needs 1000 // S to tile [0, 1000)
    import contextlib
    import io

    def run():
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            z, _ = smp.sample(S, B, (C, H, W), conditioning=ctx, num_stage=ns, eta=0.0, verbose=False)
        return z

    z = run()  # warm-up: cuDNN autotune, allocator
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    z = run()
    torch.cuda.synchronize()
    t_step = (time.perf_counter() - t0) / S  # both stages, one sampler step each
    ref.decode_first_stage(z)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ref.decode_first_stage(z)
    torch.cuda.synchronize()
    t_dec = time.perf_counter() - t0
    n_steps = cfg["steps"] + (1 if cfg["sampler"] == "plms" else 0)
    sec = n_steps * t_step + t_dec
    out = dict(value=round(B / sec, 4), unit="images/sec", kind="reference modules, eager PyTorch on this GPU",
               ms_per_sampler_step_both_stages=round(1e3 * t_step, 2), ms_decode=round(1e3 * t_dec, 2),
               tf32=dict(matmul=torch.backends.cuda.matmul.allow_tf32, cudnn=torch.backends.cudnn.allow_tf32),
               sample=f"{S} {cfg['sampler'].upper()} steps x {ns} stages at batch {B} through the reference's sampler + 1 decode, "
                      f"extrapolated to {n_steps} steps; torch {torch.__version__} defaults")
    del ref, smp
    torch.cuda.empty_cache()
    return out


def time_tc_launches(plan_step, reps=2):
    """Per-launch CUDA-event timing of the tcgen05 conv kernel inside one UNet step, ops run in program order
    (so cache state is the natural one).  Returns (sum_ms, sum_flops, n_launches, step_ms_eager)."""
    import ctypes as C

    from frido_b200 import _lib as L

    lib = L.lib()
    ops = plan_step.ops
    stream = torch.cuda.current_stream()
    sptr = C.c_void_p(stream.cuda_stream)
    tc = [i for i, op in enumerate(ops) if op.kind == L.OP_CONV and op.u.conv.engine in (1, 2, 3)]
    best = None
    for _ in range(reps + 1):
        evs = {i: (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for i in tc}
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i, op in enumerate(ops):
            if i in evs:
                evs[i][0].record(stream)
            L.check(lib.frido_run_program(C.byref(op), 1, sptr), "op")
            if i in evs:
                evs[i][1].record(stream)
        e1.record(stream)
        torch.cuda.synchronize()
        tot = sum(a.elapsed_time(b) for a, b in evs.values())
        if best is None or tot < best[0]:
            best = (tot, e0.elapsed_time(e1))
    flops = 0
    for i in tc:
        c = ops[i].u.conv
        flops += 2 * c.B * c.Hout * c.Wout * c.Cout * (c.ksize * c.ksize * (c.c0 + c.c1) + c.cx0 + c.cx1)
    return best[0], flops, len(tc), best[1]


def _traffic():
    """DRAM bytes per launch of the dominant kernel (dram__bytes_read.sum + dram__bytes_write.sum averaged over the
    tcgen05 conv launches of one UNet step), taken from the committed ncu pass of this round (profiles/traffic.json);
    bench.py itself never runs under a profiler."""
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "traffic.json")) as f:
            return json.load(f)["bytes_per_launch"]
    except Exception:
        return None


def _parity():
    """Free-running drift of the full sampler against the CPU oracle, measured by tools/prof/drift.py on the GPU box and
    committed under profiles/ (bench.py itself does not spend a minute of host time on the oracle)."""
    try:
        with open(os.path.join(ROOT, "profiles", "drift.json")) as f:
            return json.load(f)
    except Exception:
        return None


def run_ours(args):
    import frido_b200 as fb
    from frido_b200 import configs
    from frido_b200.program import default_engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    fb.lib()
    model, cfg = configs.build(args.config, dev)
    B = cfg["batch"]
    C, H, W = cfg["latent"]
    Lc, D = cfg["ctx"]
    S = cfg["steps"]
    ns = len(model.split_embed_dim_list)
    Sampler = fb.DDIMSampler if cfg["sampler"] == "ddim" else fb.PLMSSampler
    sampler = Sampler(model)
    # global synthetic inputs, sliced per rank so results do not depend on the GPU count (weak scaling: B per GPU fixed)
    g = torch.Generator().manual_seed(1)
    ctx_h = torch.randn(B * world, Lc, D, generator=g)[rank * B:(rank + 1) * B].contiguous().pin_memory()
    g = torch.Generator().manual_seed(2)
    x0_h = torch.randn(B * world, C, H, W, generator=g)[rank * B:(rank + 1) * B].contiguous().pin_memory()
    ctx_d, x0_d = ctx_h.to(dev), x0_h.to(dev)
    gathered = None
    fdec = 2 ** (model.first_stage_model.decoder.num_resolutions - 1)

    def one_batch(ctx, x0):
        """sample -> decode with the script's output formatting fused into the decoder head (uint8 NHWC, what
        custom_to_np produces, sample_diffusion.py:115-121) -> one all-gather of the uint8 images."""
        z, _ = sampler.sample(S, B, (C, H, W), conditioning=ctx, num_stage=ns, eta=0.0, verbose=False, log_every_t=10**9,
                              init_noise=x0)
        img = model.decode_first_stage_uint8(z, mode="np")
        if world > 1:  # the path's only collective: one all-gather of the finished images over NVLink
            nonlocal gathered
            if gathered is None:
                gathered = torch.empty((world,) + tuple(img.shape), dtype=img.dtype, device=dev)
            dist.all_gather_into_tensor(gathered, img)
        return img

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    launches_per_step = None
    for _ in range(args.warmup):
        img = one_batch(ctx_d, x0_d)
        launches_per_step = sampler.launches + len(next(iter(model.first_stage_model._plans.values())).prog)
    sync()
    clocks = ClockSampler(local)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    e0.record()
    for _ in range(args.steps):
        img = one_batch(ctx_d, x0_d)
    e1.record()
    sync()
    ms = e0.elapsed_time(e1)
    # ---- e2e: public API, host buffers in, host images out, every step
    sync()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    out_h = torch.empty(B, H * fdec, W * fdec, 3, dtype=torch.uint8).pin_memory()
    for _ in range(args.steps):
        c = ctx_h.to(dev, non_blocking=True)
        x = x0_h.to(dev, non_blocking=True)
        out_h.copy_(one_batch(c, x), non_blocking=True)  # result read: this rank's finished images, uint8, into pinned memory
        torch.cuda.current_stream().synchronize()
    f1.record()
    sync()
    ms_e2e = f0.elapsed_time(f1)
    clk = clocks.stop()
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
    total_imgs = B * world * args.steps
    value = total_imgs / (ms / 1e3)
    e2e_val = total_imgs / (ms_e2e / 1e3)
    if rank == 0:
        peaks = _peaks()
        unet = model.model.diffusion_model
        plan1 = unet.plan(ns - 1, B, H, W, Lc)
        plan0 = unet.plan(0, B, H, W, Lc)
        tc_ms, tc_flops, tc_n, step_ms = time_tc_launches(plan1.step)
        eng = default_engine()
        achieved = tc_flops / (tc_ms / 1e3) / 1e12
        flops_per_img = (S * sum(unet.plan(s, B, H, W, Lc).step.flops for s in range(ns))
                         + sum(unet.plan(s, B, H, W, Lc).prologue.flops for s in range(ns))) / B
        dec_plan = next(iter(model.first_stage_model._plans.values()))
        flops_per_img += dec_plan.prog.flops / B
        line = {
            "metric": _metric_name(args.config, cfg),
            "value": round(value, 4), "unit": "images/sec", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"bf16x3": "bf16x3 (error-compensated: bf16 hi/lo operand split, 3 MMAs per product, fp32 accumulate, for every "
                                "conv / linear / attention matmul on the tensor cores; short-sequence attention and norms in fp32 "
                                "SIMT) - fp32-faithful to ~1e-4 on eps",
                      "tc3": "tf32x3 (error-compensated 3xTF32 operands, fp32 accumulate: fp32-faithful)",
                      "tc": "tf32 (fp32 accumulate)", "simt": "f32"}[eng],
            "data": "synthetic (random-init weights with zero_module tensors re-drawn, N(0,1) context and start noise, eta=0)",
            "config": {"workload": _workload_name(args.config, cfg, ns),
                       "batch_per_gpu": B, "sampler_steps": S, "stages": ns, "engine": eng,
                       "l2": "working set of one step (weights 2 GB + activations) >> 126 MB L2; no explicit flush",
                       "parallelism": f"batch-sharded x{world}, one all-gather of the uint8 images" if world > 1 else "single GPU"},
            "e2e": {"value": round(e2e_val, 4), "unit": "images/sec",
                    "h2d_bytes_per_step": int(ctx_h.numel() * 4 + x0_h.numel() * 4), "d2h_bytes_per_step": int(out_h.numel()),
                    "result": "uint8 NHWC images (custom_to_np format) of this rank's batch"},
            "gpu_launches": int(launches_per_step * args.steps),
            "clocks": clk,
            "roofline": {"bound": "tensor", "achieved": round(achieved, 2), "peak": peaks["bf16"], "unit": "TFLOP/s",
                         "frac": round(achieved / peaks["bf16"], 4), "traffic": _traffic(),
                         "kernel": {"bf16x3": "frido::conv_tc_bf_kernel + conv_nf_kernel (BF16x3 tcgen05 implicit-GEMM convs)",
                                    "tc3": "frido::conv_tc_kernel<1> (3xTF32)", "tc": "frido::conv_tc_kernel<0> (TF32)"}.get(eng, "conv"),
                         "note": f"algorithmic FLOPs (1x) of the {tc_n} tcgen05 conv launches of one stage-{ns - 1} UNet step at batch {B} "
                                 f"/ their summed CUDA-event time ({tc_ms:.2f} ms of a {step_ms:.2f} ms eager step); peak = {peaks['source']} "
                                 "dense bf16 sustained; error-compensated modes issue 3 MMAs per product and TF32 issues at half the bf16 "
                                 "rate, so the ceiling of this kernel in algorithmic TFLOP/s is peak/3 (bf16x3), peak/6 (tc3), peak/2 (tc)",
                         "issued_tflops": round(achieved * {"bf16x3": 3, "tc3": 3}.get(eng, 1), 1),
                         "share_of_step": round(tc_ms / step_ms, 3)},
            "tflop_per_image": round(flops_per_img / 1e12, 3),
            "achieved_tflops_whole_job": round(flops_per_img * value / 1e12 / world, 2),
        }
        line["parity"] = _parity()
        if world == 1:
            # the GPU-eager run of the reference's modules goes first: the CPU arm below patches .cuda() to a no-op
            if os.environ.get("FRIDO_BENCH_GPU_EAGER", "1") == "1":
                try:
                    line["gpu_eager_reference"] = gpu_eager_reference(model, cfg, dev)
                    if line["gpu_eager_reference"]:
                        line["gpu_eager_reference"]["speedup_e2e"] = round(e2e_val / line["gpu_eager_reference"]["value"], 2)
                except Exception as e:
                    import traceback
                    line["gpu_eager_reference"] = {"error": repr(e), "where": traceback.format_exc().strip().splitlines()[-7:]}
            try:
                line["cpu_baseline"] = cpu_sample(model, cfg, min(B, 4), evals=2, ref=_reference_model(model, cfg, "cpu"), budget_s=30.0)
            except Exception as e:  # the checker failing must not hide the GPU number
                line["cpu_baseline"] = {"error": repr(e)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path on the box's host cores, all of them
    (`oracle/_ref` = the reference's unmodified modules when the archive is present, else the pinned oracle port).
    Each step = a bounded sample of the workload at the config's batch: 3 UNet evaluations per stage + 1 decode,
    extrapolated linearly to the config's step count.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from frido_b200 import configs

    cores = _host_cores()
    model, cfg = configs.build(args.config, "cpu")
    B = cfg["batch"]
    ref = _reference_model(model, cfg, "cpu")
    vals = []
    t_all0 = time.perf_counter()
    for i in range(max(1, args.steps)):
        if i >= 1 and time.perf_counter() - t_all0 > 150:  # keep the whole run within a few minutes
            break
        vals.append(cpu_sample(model, cfg, B, evals=3, ref=ref, budget_s=100.0))
    v = statistics.median([x["value"] for x in vals])
    C, H, W = cfg["latent"]
    ns = len(model.split_embed_dim_list)
    line = {"impl": "reference", "metric": _metric_name(args.config, cfg), "value": v,
            "unit": "images/sec", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": len(vals), "warmup": 1,
            "ms_per_step": round(1e3 * B / v, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic (same weights / shapes as the GPU arm)",
            "config": {"workload": _workload_name(args.config, cfg, ns), "batch_per_gpu": B, "sampler_steps": cfg["steps"], "stages": ns,
                       "engine": "the reference's own modules on the host CPU" if ref is not None else "reference algorithm on host CPU (oracle port)",
                       "sample": "each step = 3 UNet evals per stage + 1 decode at the config's batch, extrapolated linearly to the full step count",
                       "cores": cores},
            "cpu_baseline": dict(vals[-1], value=v),
            "e2e": {"value": v, "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def _metric_name(name, cfg):
    if name == "l2i_coco":
        return "images/sec DDIM-200 256x256 layout-to-image (sampling + decode)"
    return f"images/sec {cfg['sampler'].upper()}-{cfg['steps']} ({name}, sampling + decode)"


def _workload_name(name, cfg, ns):
    C, H, W = cfg["latent"]
    which = {"l2i_coco": "BASELINE configs[1]", "t2i_clip": "BASELINE configs[2]", "sg2i_vg": "BASELINE configs[3]",
             "l2i_512": "BASELINE configs[4]"}.get(name, "")
    return (f"{name}: latent {C}x{H}x{W}, context {cfg['ctx'][0]}x{cfg['ctx'][1]}, {cfg['sampler'].upper()}-{cfg['steps']} x {ns} stages "
            f"+ MS-VQGAN decode, batch {cfg['batch']} per GPU ({which})")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="l2i_coco")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
