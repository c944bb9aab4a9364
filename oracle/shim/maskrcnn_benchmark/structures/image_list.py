import torch


class ImageList(object):
    def __init__(self, tensors, image_sizes):
        self.tensors = tensors
        self.image_sizes = image_sizes

    def to(self, *a, **k):
        return ImageList(self.tensors.to(*a, **k), self.image_sizes)


def to_image_list(tensors, size_divisible=0):
    if isinstance(tensors, ImageList):
        return tensors
    if isinstance(tensors, torch.Tensor):
        if tensors.dim() == 3:
            tensors = tensors[None]
        assert tensors.dim() == 4
        return ImageList(tensors, [t.shape[-2:] for t in tensors])
    raise TypeError("only tensors supported by this stand-in (size_divisible=0 path)")
