from frido_b200.samplers import DDIMSampler  # noqa: F401  (frido/models/diffusion/ddim.py)
