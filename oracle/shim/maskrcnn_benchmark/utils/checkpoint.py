import torch
from .model_serialization import load_state_dict


class DetectronCheckpointer(object):
    def __init__(self, cfg, model, *a, **k):
        self.model = model

    def load(self, f=None, use_latest=False):
        ckpt = torch.load(f, map_location="cpu")
        load_state_dict(self.model, ckpt.get("model", ckpt))
        return ckpt
