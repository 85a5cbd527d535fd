"""frido_b200 — B200-native (sm_100a) implementation of Frido's multi-scale
denoising sampling hot path behind the reference's Python class surface.

    from frido_b200 import FridoDiffusion, DDIMSampler, PLMSSampler

All device arithmetic is in libfrido_b200.so (hand-written CUDA, C ABI in
include/frido_b200.h); there is no CPU or PyTorch compute fallback.
"""
from ._lib import FridoError, lib  # noqa: F401
from .cond import BERTEmbedder  # noqa: F401
from .diffusion import DiffusionWrapper, FridoDiffusion, LitEma, instantiate_from_config  # noqa: F401
from .first_stage import VQModelInterface, images_to_uint8  # noqa: F401
from .output import SampleWriter  # noqa: F401
from .samplers import DDIMSampler, PLMSSampler  # noqa: F401
from .unet import PyUNetModel  # noqa: F401

__all__ = ["FridoDiffusion", "DiffusionWrapper", "LitEma", "PyUNetModel", "VQModelInterface", "DDIMSampler",
           "PLMSSampler", "FridoError", "instantiate_from_config", "lib", "SampleWriter", "images_to_uint8"]


def install_aliases():
    """Make the reference's import paths (`frido.models.diffusion.ddim`, `taming.models.msvqgan`, ...) and YAML
    `target:` strings resolve to this package (see frido_b200/compat/README.md)."""
    import os
    import sys

    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "compat")
    if p not in sys.path:
        sys.path.insert(0, p)
    for k in [k for k in sys.modules if k.split(".")[0] in ("frido", "taming")]:
        del sys.modules[k]
