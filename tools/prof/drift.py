"""Free-running drift report (SURVEY.md §7 "hard parts"): the full DDIM-S x 2-stage sampler on the full-size layout2img
UNet (32x32 latent, B = 1), GPU vs the CPU oracle from the same start noise, no teacher forcing, for the product engine
(bf16x3) and the fp32 SIMT engine as the yardstick of what two fp32 implementations differ by.  Writes the JSON that
bench.py reports as `parity` (profiles/drift.json).

  python tools/prof/drift.py [--steps 200] [--out profiles/drift.json]
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def one_engine(steps):
    import torch

    torch.set_grad_enabled(False)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_gpu_benched as T

    out, ref = T._drift_case(torch.device("cuda:0"), steps)
    d = (out - ref).abs()
    return dict(absmax=float(d.max()), rms=float((out - ref).pow(2).mean().sqrt()), latent_std=float(ref.std()),
                latent_absmax=float(ref.abs().max()))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "drift.json"))
    ap.add_argument("--engine", default=None)
    a = ap.parse_args()
    if a.engine:  # child: one engine per process (FRIDO_ENGINE is read at plan build)
        print("DRIFT " + json.dumps(one_engine(a.steps)))
        sys.exit(0)
    res = {}
    for eng in ("bf16x3", "simt"):
        t0 = time.time()
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--steps", str(a.steps), "--engine", eng],
                           env=dict(os.environ, FRIDO_ENGINE=eng), capture_output=True, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("DRIFT ")]
        if r.returncode != 0 or not line:
            res[eng] = dict(error=(r.stderr or r.stdout)[-400:])
            continue
        res[eng] = json.loads(line[0][6:])
        res[eng]["seconds"] = round(time.time() - t0, 1)
    out = dict(free_running_absmax=res.get("bf16x3", {}).get("absmax"), free_running_rms=res.get("bf16x3", {}).get("rms"),
               case=f"DDIM-{a.steps} x 2 stages, eta 0, full-size layout2img UNet (511 M parameters), latent 6x32x32, B=1, "
                    "final latent GPU vs CPU oracle (oracle/torch_oracle.py), same start noise, no teacher forcing",
               engines=res, tolerance="north star: |delta| < 1e-3 fp32")
    with open(a.out, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))
