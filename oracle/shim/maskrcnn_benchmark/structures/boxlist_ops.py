import torch

from oracle import prims
from .bounding_box import BoxList


def boxlist_nms(boxlist, nms_thresh, max_proposals=-1, score_field="scores"):
    if nms_thresh <= 0:
        return boxlist
    mode = boxlist.mode
    boxlist = boxlist.convert("xyxy")
    keep = prims.nms_legacy(boxlist.bbox, boxlist.get_field(score_field), nms_thresh)
    if max_proposals > 0:
        keep = keep[:max_proposals]
    return boxlist[keep].convert(mode)


def remove_small_boxes(boxlist, min_size):
    keep = prims.remove_small_mask(boxlist.convert("xyxy").bbox, min_size).nonzero().squeeze(1)
    return boxlist[keep]


def boxlist_iou(a, b):
    raise NotImplementedError("training-only")


def _cat(tensors, dim=0):
    if len(tensors) == 1:
        return tensors[0]
    return torch.cat(tensors, dim)


def cat_boxlist(bboxes):
    size = bboxes[0].size
    assert all(tuple(b.size) == tuple(size) for b in bboxes)
    mode = bboxes[0].mode
    assert all(b.mode == mode for b in bboxes)
    fields = set(bboxes[0].fields())
    assert all(set(b.fields()) == fields for b in bboxes)
    out = BoxList(_cat([b.bbox for b in bboxes], dim=0), size, mode)
    for f in fields:
        out.add_field(f, _cat([b.get_field(f) for b in bboxes], dim=0))
    return out
