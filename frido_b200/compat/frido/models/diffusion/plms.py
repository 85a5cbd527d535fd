from frido_b200.samplers import PLMSSampler  # noqa: F401  (frido/models/diffusion/plms.py)
