"""frido/util.py surface used by scripts/sample_diffusion.py: instantiate_from_config(+_main)."""
from frido_b200.diffusion import get_obj_from_str, instantiate_from_config  # noqa: F401


def instantiate_from_config_main(config, *args, **kwargs):
    if "target" not in config:
        raise KeyError("Expected key `target` to instantiate.")
    return get_obj_from_str(config["target"])(*args, **dict(config.get("params", dict())), **kwargs)
