"""Import shim (TEST INFRASTRUCTURE ONLY): lets /root/reference import without
pytorch-lightning 1.0.8.  Only the names the reference touches at import /
construction time exist; nothing here is product code."""
import torch
from torch import nn


class LightningModule(nn.Module):
    @property
    def device(self):
        for p in self.parameters():
            return p.device
        for b in self.buffers():
            return b.device
        return torch.device("cpu")

    def log(self, *a, **k):
        pass

    def log_dict(self, *a, **k):
        pass


class LightningDataModule:
    pass


class Callback:
    pass


class Trainer:
    pass


def seed_everything(seed):
    torch.manual_seed(seed)
