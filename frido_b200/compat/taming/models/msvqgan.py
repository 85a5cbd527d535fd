from frido_b200.first_stage import VQModelInterface  # noqa: F401  (taming/models/msvqgan.py)
