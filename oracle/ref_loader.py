"""TEST INFRASTRUCTURE ONLY — import the *unmodified* reference from
/root/reference on CPU (build container only; the GPU box has no /root/reference).

Recipe = SURVEY.md Appendix D: three stub packages on sys.path
(oracle/shims), identity `.cuda()` on CPU, samplers' register_buffer → setattr,
stale `ldm.*` targets rewritten to `frido.*`.  Nothing is copied out of the
reference; it is imported where it lies.
"""
import copy
import os
import sys

import torch
import yaml

_HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("FRIDO_REFERENCE", "/root/reference")
if not os.path.isdir(os.path.join(REF, "frido")):
    # GPU box: the archive packed by oracle/vendor_ref.py (git-ignored build artefact, travels with the snapshot)
    sys.path.insert(0, os.path.dirname(_HERE))
    from oracle import vendor_ref as _vr
    REF = _vr.unpack() or REF

_TARGET_REWRITE = {
    "ldm.models.diffusion.msldm.MSLatentDiffusion": "frido.models.diffusion.frido.FridoDiffusion",
    "ldm.modules.diffusionmodules.openaimodel.UNetModel": "frido.modules.diffusionmodules.pyunet.PyUNetModel",
    "ldm.modules.diffusionmodules.pyunet.PyUNetModel": "frido.modules.diffusionmodules.pyunet.PyUNetModel",
}


_ORIG_CUDA = (torch.Tensor.cuda, torch.nn.Module.cuda)


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "frido"))


def activate(cpu: bool = True):
    """Make `import frido`, `import taming` resolve to the reference."""
    shims = os.path.join(_HERE, "shims")
    for p in (REF, shims):
        if p not in sys.path:
            sys.path.insert(0, p)
    # the repo's own drop-in mirror uses the same top-level names; make sure the
    # reference wins inside this process
    for k in [k for k in sys.modules if k == "frido" or k.startswith("frido.") or k == "taming" or k.startswith("taming.")]:
        mod = sys.modules[k]
        f = getattr(mod, "__file__", None) or ""
        paths = [str(p) for p in getattr(mod, "__path__", [])]  # namespace packages have no __file__
        if not (f.startswith(REF) or any(p.startswith(REF) for p in paths)):
            del sys.modules[k]
    if cpu:
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    else:  # a previous CPU activation in this process must not leak into a GPU run
        torch.Tensor.cuda, torch.nn.Module.cuda = _ORIG_CUDA
    from frido.models.diffusion.ddim import DDIMSampler
    from frido.models.diffusion.plms import PLMSSampler

    def _reg(self, name, attr):
        setattr(self, name, attr)

    for cls in (DDIMSampler, PLMSSampler):
        if not hasattr(cls, "_frido_orig_register_buffer"):
            cls._frido_orig_register_buffer = cls.register_buffer
        cls.register_buffer = _reg if cpu else cls._frido_orig_register_buffer
    return DDIMSampler, PLMSSampler


def _rewrite(node):
    if isinstance(node, dict):
        out = {}
        for k, v in node.items():
            if k == "target" and isinstance(v, str):
                v = _TARGET_REWRITE.get(v, v)
                if v.startswith("ldm."):
                    v = "frido." + v[4:]
            out[k] = _rewrite(v)
        return out
    if isinstance(node, list):
        return [_rewrite(v) for v in node]
    return node


def load_config(rel_path: str):
    with open(os.path.join(REF, rel_path)) as f:
        return _rewrite(yaml.safe_load(f))


class AttrDict(dict):
    """frido.py:538 reads `first_stage_config.params.ddconfig.ch_mult` by attribute."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return AttrDict(v) if isinstance(v, dict) and not isinstance(v, AttrDict) else v


def build_model(cfg_model: dict, cond_stage: bool = False, cpu: bool = True):
    """instantiate_from_config(cfg['model']) on CPU with ckpt_path=None and no EMA.
    cond_stage=False replaces the (out-of-scope, 32-layer) condition encoder
    with '__is_unconditional__' *after* fixing conditioning_key, so the UNet
    still takes crossattn context."""
    activate(cpu=cpu)
    from frido.util import instantiate_from_config

    cfg = copy.deepcopy(cfg_model)
    p = cfg["params"]
    p["first_stage_config"]["params"]["ckpt_path"] = None
    p["first_stage_config"]["params"]["init_normal"] = True
    p["use_ema"] = False
    ckey = p.get("conditioning_key", "crossattn")
    if not cond_stage:
        p["cond_stage_config"] = "__is_unconditional__"
        p["cond_stage_trainable"] = False
    p["first_stage_config"] = AttrDict(p["first_stage_config"])
    model = instantiate_from_config(cfg).eval()
    if not cond_stage:
        model.model.conditioning_key = ckey
    for q in model.parameters():
        q.requires_grad_(False)
    return model
