from frido_b200.diffusion import DiffusionWrapper, FridoDiffusion, LitEma  # noqa: F401  (frido/models/diffusion/frido.py)
