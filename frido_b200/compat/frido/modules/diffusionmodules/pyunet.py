from frido_b200.unet import PyUNetModel  # noqa: F401  (frido/modules/diffusionmodules/pyunet.py)
