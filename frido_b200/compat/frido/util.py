"""`frido.util` as seen through the drop-in aliases.

scripts/sample_diffusion.py:16-19 imports `log_txt_as_img, exists, default, ismap, isimage, mean_flat, count_params`
and `instantiate_from_config_main` from here.  When a reference checkout follows this directory on sys.path its own
frido/util.py is executed by path and its whole namespace re-exported; only the two `instantiate_from_config*`
entry points (frido/util.py:74-95) are overridden so that YAML `target:` strings resolve to the B200-native classes.
Without a reference checkout the small helpers below stand in (same names, same results).
"""
import importlib.util as _ilu
import os as _os
import sys as _sys


def _reference_util():
    here = _os.path.abspath(_os.path.dirname(__file__))
    for p in _sys.path:
        cand = _os.path.join(p or ".", "frido", "util.py")
        if _os.path.isfile(cand) and _os.path.abspath(_os.path.dirname(cand)) != here:
            spec = _ilu.spec_from_file_location("frido._reference_util", cand)
            mod = _ilu.module_from_spec(spec)
            spec.loader.exec_module(mod)
            return mod
    return None


_ref = _reference_util()
if _ref is not None:
    globals().update({k: v for k, v in vars(_ref).items() if not k.startswith("__")})
else:
    import numpy as _np
    import torch as _torch

    def exists(x):
        return x is not None

    def default(val, d):
        if val is not None:
            return val
        return d() if callable(d) and not isinstance(d, type) else d

    def ismap(x):
        return isinstance(x, _torch.Tensor) and x.dim() == 4 and x.shape[1] > 3

    def isimage(x):
        return isinstance(x, _torch.Tensor) and x.dim() == 4 and x.shape[1] in (1, 3)

    def mean_flat(tensor):
        return tensor.mean(dim=list(range(1, tensor.dim())))

    def count_params(model, verbose=False):
        n = sum(p.numel() for p in model.parameters())
        if verbose:
            print(f"{model.__class__.__name__} has {n * 1.e-6:.2f} M params.")
        return n

    def log_txt_as_img(wh, xc, size=10):
        """Captions rendered as white [B,3,H,W] images in [-1,1] (frido/util.py:10-34)."""
        from PIL import Image, ImageDraw, ImageFont

        try:
            font = ImageFont.truetype("data/DejaVuSans.ttf", size=size)
        except OSError:
            font = ImageFont.load_default()
        per_line = int(40 * (wh[0] / 256))
        rows = []
        for cap in xc:
            cap = " ".join(str(c) for c in cap) if isinstance(cap, list) else str(cap)
            canvas = Image.new("RGB", wh, color="white")
            text = "\n".join(cap[i:i + per_line] for i in range(0, len(cap), per_line))
            try:
                ImageDraw.Draw(canvas).text((0, 0), text, fill="black", font=font)
            except UnicodeEncodeError:
                print("Cant encode string for logging. Skipping.")
            rows.append(_np.asarray(canvas).transpose(2, 0, 1) / 127.5 - 1.0)
        return _torch.tensor(_np.stack(rows))


from frido_b200.diffusion import get_obj_from_str, instantiate_from_config  # noqa: E402,F401  (override: native targets)


def instantiate_from_config_main(config, *args, **kwargs):
    """frido/util.py:84-88."""
    if "target" not in config:
        raise KeyError("Expected key `target` to instantiate.")
    return get_obj_from_str(config["target"])(*args, **dict(config.get("params", dict())), **kwargs)
