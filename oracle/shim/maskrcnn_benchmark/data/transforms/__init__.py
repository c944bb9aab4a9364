"""Stand-in for upstream ``maskrcnn_benchmark.data.transforms`` (un-vendored): the one class the reference's test transform
takes from it (build_augmentation.py:48-50).  TEST INFRASTRUCTURE; see oracle/shim/maskrcnn_benchmark/__init__.py."""
import torch
from torchvision.transforms import functional as F


class Normalize(object):
    """upstream transforms.Normalize: optional RGB->BGR*255, then (x - mean) / std; passes the target through."""

    def __init__(self, mean, std, to_bgr255=True):
        self.mean, self.std, self.to_bgr255 = mean, std, to_bgr255

    def __call__(self, image, target=None):
        if self.to_bgr255:
            image = image[[2, 1, 0]] * 255
        image = F.normalize(image, mean=self.mean, std=self.std)
        if target is None:
            return image
        return image, target
