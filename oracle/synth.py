"""TEST INFRASTRUCTURE ONLY — deterministic synthetic weights and inputs.

The reference ships no checkpoints that are reachable here, so
every parity case runs on synthetic weights.  To keep golden fixtures tiny the
weights are *not* stored: they are regenerated from (parameter name, shape,
seed) with numpy's PCG64 (stable across platforms and numpy versions), on both
the reference side (oracle/make_golden.py, run in the build container) and the
checker side (tests/, smoke(), bench cpu_baseline).

Un-zeroes the reference's `zero_module` tensors (pyunet.py:236-238,801,
attention.py:276) — an untouched random-init UNet outputs exactly 0 and would
make every parity test vacuous — and draws codebooks from N(0,1)
(quantize.py:223-226 `init_normal=True`).
"""
import zlib

import numpy as np
import torch


def _rng(name: str, seed: int):
    return np.random.default_rng((zlib.crc32(name.encode()) << 8) ^ seed)


def synth_tensor(name: str, shape, seed: int = 0) -> torch.Tensor:
    shape = tuple(int(s) for s in shape)
    r = _rng(name, seed)
    n = r.standard_normal(shape, dtype=np.float32)
    if name.endswith("embedding.weight"):  # VQ codebook
        v = n
    elif len(shape) == 1:
        if name.endswith("weight"):  # norm scales (GN/LN): around 1
            v = 1.0 + 0.1 * n
        else:  # biases
            v = 0.05 * n
    else:
        fan_in = int(np.prod(shape[1:]))
        v = n * np.float32(1.0 / np.sqrt(fan_in))
    return torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))


def synth_state_dict(manifest, seed: int = 0):
    """manifest: iterable of (name, shape) -> {name: fp32 tensor}."""
    return {name: synth_tensor(name, shape, seed) for name, shape in manifest}


def fill_module_(module: torch.nn.Module, seed: int = 0, prefix: str = ""):
    """Overwrite every *parameter* of `module` in place (buffers — schedules,
    scale_factor, EMA shadows — are left alone)."""
    with torch.no_grad():
        for name, p in module.named_parameters():
            p.copy_(synth_tensor(prefix + name, p.shape, seed).to(p.device, p.dtype))
    return module


def manifest_of(module: torch.nn.Module, prefix: str = ""):
    return [(prefix + n, list(p.shape)) for n, p in module.named_parameters()]


def synth_input(tag: str, shape, seed: int = 0, scale: float = 1.0) -> torch.Tensor:
    n = _rng("input:" + tag, seed).standard_normal(tuple(shape), dtype=np.float32)
    return torch.from_numpy(n * np.float32(scale))
