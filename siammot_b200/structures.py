"""BoxList and box-list operations with the reference's (maskrcnn_benchmark) interface and the legacy
"+1" pixel convention: ``bbox`` (N,4) fp32, ``size`` = (W, H), ``mode`` in
{"xyxy","xywh"}, named per-box ``extra_fields``.  Used at the API boundary of the engine
(``SiamMOT.forward`` returns ``[BoxList]`` like /root/reference/siammot/modelling/rcnn.py:68)."""
import torch

TO_REMOVE = 1


class BoxList(object):
    def __init__(self, bbox, image_size, mode="xyxy"):
        device = bbox.device if isinstance(bbox, torch.Tensor) else torch.device("cpu")
        bbox = torch.as_tensor(bbox, dtype=torch.float32, device=device)
        if bbox.ndimension() != 2:
            raise ValueError("bbox should have 2 dimensions, got {}".format(bbox.ndimension()))
        if bbox.size(-1) != 4:
            raise ValueError("last dimension of bbox should have a size of 4, got {}".format(bbox.size(-1)))
        if mode not in ("xyxy", "xywh"):
            raise ValueError("mode should be 'xyxy' or 'xywh'")
        self.bbox = bbox
        self.size = image_size  # (image_width, image_height)
        self.mode = mode
        self.extra_fields = {}

    # ---- fields
    def add_field(self, field, field_data):
        self.extra_fields[field] = field_data

    def get_field(self, field):
        return self.extra_fields[field]

    def has_field(self, field):
        return field in self.extra_fields

    def fields(self):
        return list(self.extra_fields.keys())

    def _copy_extra_fields(self, other):
        for k, v in other.extra_fields.items():
            self.extra_fields[k] = v

    # ---- geometry
    def _split_into_xyxy(self):
        if self.mode == "xyxy":
            return self.bbox.split(1, dim=-1)
        xmin, ymin, w, h = self.bbox.split(1, dim=-1)
        return xmin, ymin, xmin + (w - TO_REMOVE).clamp(min=0), ymin + (h - TO_REMOVE).clamp(min=0)

    def convert(self, mode):
        if mode not in ("xyxy", "xywh"):
            raise ValueError("mode should be 'xyxy' or 'xywh'")
        if mode == self.mode:
            return self
        xmin, ymin, xmax, ymax = self._split_into_xyxy()
        if mode == "xyxy":
            bbox = BoxList(torch.cat((xmin, ymin, xmax, ymax), dim=-1), self.size, mode=mode)
        else:
            bbox = BoxList(torch.cat((xmin, ymin, xmax - xmin + TO_REMOVE, ymax - ymin + TO_REMOVE), dim=-1),
                           self.size, mode=mode)
        bbox._copy_extra_fields(self)
        return bbox

    def resize(self, size, *args, **kwargs):
        ratios = tuple(float(s) / float(s_orig) for s, s_orig in zip(size, self.size))
        if ratios[0] == ratios[1]:
            bbox = BoxList(self.bbox * ratios[0], size, mode=self.mode)
            for k, v in self.extra_fields.items():
                if not isinstance(v, torch.Tensor):
                    v = v.resize(size, *args, **kwargs)
                bbox.add_field(k, v)
            return bbox
        ratio_width, ratio_height = ratios
        xmin, ymin, xmax, ymax = self._split_into_xyxy()
        scaled = torch.cat((xmin * ratio_width, ymin * ratio_height, xmax * ratio_width, ymax * ratio_height), dim=-1)
        bbox = BoxList(scaled, size, mode="xyxy")
        for k, v in self.extra_fields.items():
            if not isinstance(v, torch.Tensor):
                v = v.resize(size, *args, **kwargs)
            bbox.add_field(k, v)
        return bbox.convert(self.mode)

    def clip_to_image(self, remove_empty=True):
        self.bbox[:, 0].clamp_(min=0, max=self.size[0] - TO_REMOVE)
        self.bbox[:, 1].clamp_(min=0, max=self.size[1] - TO_REMOVE)
        self.bbox[:, 2].clamp_(min=0, max=self.size[0] - TO_REMOVE)
        self.bbox[:, 3].clamp_(min=0, max=self.size[1] - TO_REMOVE)
        if remove_empty:
            box = self.bbox
            keep = (box[:, 3] > box[:, 1]) & (box[:, 2] > box[:, 0])
            return self[keep]
        return self

    def area(self):
        box = self.bbox
        if self.mode == "xyxy":
            return (box[:, 2] - box[:, 0] + TO_REMOVE) * (box[:, 3] - box[:, 1] + TO_REMOVE)
        return box[:, 2] * box[:, 3]

    # ---- container behaviour
    def to(self, device):
        bbox = BoxList(self.bbox.to(device), self.size, self.mode)
        for k, v in self.extra_fields.items():
            if hasattr(v, "to"):
                v = v.to(device)
            bbox.add_field(k, v)
        return bbox

    def __getitem__(self, item):
        bbox = BoxList(self.bbox[item], self.size, self.mode)
        for k, v in self.extra_fields.items():
            bbox.add_field(k, v[item])
        return bbox

    def __len__(self):
        return self.bbox.shape[0]

    def copy_with_fields(self, fields, skip_missing=False):
        bbox = BoxList(self.bbox, self.size, self.mode)
        if not isinstance(fields, (list, tuple)):
            fields = [fields]
        for field in fields:
            if self.has_field(field):
                bbox.add_field(field, self.get_field(field))
            elif not skip_missing:
                raise KeyError("Field '{}' not found in {}".format(field, self))
        return bbox

    def __repr__(self):
        return "{}(num_boxes={}, image_width={}, image_height={}, mode={})".format(
            self.__class__.__name__, len(self), self.size[0], self.size[1], self.mode)


def _cat(tensors, dim=0):
    if len(tensors) == 1:
        return tensors[0]
    return torch.cat(tensors, dim)


def cat_boxlist(bboxes):
    size = bboxes[0].size
    if not all(tuple(b.size) == tuple(size) for b in bboxes):
        raise AssertionError("cat_boxlist: all boxlists must share one image size")
    mode = bboxes[0].mode
    assert all(b.mode == mode for b in bboxes)
    fields = set(bboxes[0].fields())
    assert all(set(b.fields()) == fields for b in bboxes)
    out = BoxList(_cat([b.bbox for b in bboxes], dim=0), size, mode)
    for f in fields:
        out.add_field(f, _cat([b.get_field(f) for b in bboxes], dim=0))
    return out


def boxlist_nms(boxlist, nms_thresh, max_proposals=-1, score_field="scores"):
    """Same contract as upstream boxlist_nms -> _C.nms, computed by libsmot's device sort+NMS kernel
    (score-descending, ties by index, IoU(+1) > thresh).  Needs CUDA: tensors are moved to the
    current device if they are not there (no CPU implementation exists in the product)."""
    from . import ops
    if nms_thresh <= 0:
        return boxlist
    mode = boxlist.mode
    boxlist = boxlist.convert("xyxy")
    n = len(boxlist)
    if n == 0:
        return boxlist.convert(mode)
    if n > 4096:
        raise ValueError("boxlist_nms: at most 4096 boxes per call (got %d)" % n)
    dev = boxlist.bbox.device if boxlist.bbox.is_cuda else torch.device("cuda")
    boxes = boxlist.bbox.to(dev).contiguous()
    scores = boxlist.get_field(score_field).to(dev, torch.float32).contiguous()
    count = torch.zeros(1, dtype=torch.int32, device=dev)
    idx = torch.empty(n, dtype=torch.int32, device=dev)
    ops.sort_nms(boxes, scores, count, thresh=float(nms_thresh), max_keep=max_proposals if max_proposals > 0 else n,
                 out_index=idx)
    keep = idx[: int(count.item())].to(torch.int64).to(boxlist.bbox.device)
    return boxlist[keep].convert(mode)


def remove_small_boxes(boxlist, min_size):
    xywh = boxlist.convert("xywh").bbox
    keep = ((xywh[:, 2] >= min_size) & (xywh[:, 3] >= min_size)).nonzero().squeeze(1)
    return boxlist[keep]
