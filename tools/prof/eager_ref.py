"""GPU box: the reference's own modules in eager PyTorch on the B200 (bench.py's secondary baseline), standalone with a
full traceback.  PCONFIG selects the config (default l2i_coco)."""
import json
import os
import sys
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

torch.set_grad_enabled(False)
import bench
from frido_b200 import configs

dev = torch.device("cuda", 0)
model, cfg = configs.build(os.environ.get("PCONFIG", "l2i_coco"), dev)
try:
    print(json.dumps(bench.gpu_eager_reference(model, cfg, dev)))
except Exception:
    traceback.print_exc()
