"""TEST INFRASTRUCTURE ONLY: stand-in so that the reference's LightningModule subclasses import without Lightning."""
import torch


class LightningModule(torch.nn.Module):
    def log(self, *a, **k):
        pass

    def log_dict(self, *a, **k):
        pass


class Callback:
    pass
