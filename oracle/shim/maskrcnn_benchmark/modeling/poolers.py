import torch
from torch import nn

from oracle import prims
from maskrcnn_benchmark.layers import ROIAlign
from .utils import cat


class LevelMapper(object):
    def __init__(self, k_min, k_max, canonical_scale=224, canonical_level=4, eps=1e-6):
        self.k_min, self.k_max = k_min, k_max
        self.s0, self.lvl0, self.eps = canonical_scale, canonical_level, eps

    def __call__(self, boxlists):
        boxes = cat([b.convert("xyxy").bbox for b in boxlists])
        return prims.map_levels(boxes, self.k_min, self.k_max, self.s0, self.lvl0, self.eps)


class Pooler(nn.Module):
    def __init__(self, output_size, scales, sampling_ratio):
        super().__init__()
        self.poolers = nn.ModuleList([ROIAlign(output_size, spatial_scale=s, sampling_ratio=sampling_ratio)
                                      for s in scales])
        self.output_size = output_size
        lvl_min = -torch.log2(torch.tensor(scales[0], dtype=torch.float32)).item()
        lvl_max = -torch.log2(torch.tensor(scales[-1], dtype=torch.float32)).item()
        self.map_levels = LevelMapper(lvl_min, lvl_max)

    def convert_to_roi_format(self, boxes):
        concat_boxes = cat([b.bbox for b in boxes], dim=0)
        ids = cat([torch.full((len(b), 1), i, dtype=concat_boxes.dtype) for i, b in enumerate(boxes)], dim=0)
        return torch.cat([ids, concat_boxes], dim=1)

    def forward(self, x, boxes):
        rois = self.convert_to_roi_format(boxes)
        if len(self.poolers) == 1:
            return self.poolers[0](x[0], rois)
        levels = self.map_levels(boxes)
        res = self.output_size[0]
        result = torch.zeros((len(rois), x[0].shape[1], res, res), dtype=x[0].dtype)
        for level, (feat, pooler) in enumerate(zip(x, self.poolers)):
            idx = torch.nonzero(levels == level).squeeze(1)
            result[idx] = pooler(feat, rois[idx]).to(x[0].dtype)
        return result
