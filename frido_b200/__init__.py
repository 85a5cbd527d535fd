"""frido_b200 — B200-native (sm_100a) implementation of Frido's multi-scale
denoising sampling hot path behind the reference's Python class surface.

    from frido_b200 import FridoDiffusion, DDIMSampler, PLMSSampler

All device arithmetic is in libfrido_b200.so (hand-written CUDA, C ABI in
include/frido_b200.h); there is no CPU or PyTorch compute fallback.
"""
from ._lib import FridoError, lib  # noqa: F401
from .diffusion import DiffusionWrapper, FridoDiffusion, LitEma, instantiate_from_config  # noqa: F401
from .first_stage import VQModelInterface  # noqa: F401
from .samplers import DDIMSampler, PLMSSampler  # noqa: F401
from .unet import PyUNetModel  # noqa: F401

__all__ = ["FridoDiffusion", "DiffusionWrapper", "LitEma", "PyUNetModel", "VQModelInterface", "DDIMSampler",
           "PLMSSampler", "FridoError", "instantiate_from_config", "lib"]
