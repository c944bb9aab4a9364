import torch

from oracle import prims

TO_REMOVE = 1


class BoxList(object):
    """Boxes (N,4) f32 + image size (W,H) + mode + per-box fields, legacy +1 convention."""

    def __init__(self, bbox, image_size, mode="xyxy"):
        device = bbox.device if isinstance(bbox, torch.Tensor) else torch.device("cpu")
        bbox = torch.as_tensor(bbox, dtype=torch.float32, device=device)
        if bbox.ndimension() != 2 or bbox.size(-1) != 4:
            raise ValueError("bbox should be (N,4), got {}".format(tuple(bbox.shape)))
        if mode not in ("xyxy", "xywh"):
            raise ValueError("mode should be 'xyxy' or 'xywh'")
        self.bbox = bbox
        self.size = image_size
        self.mode = mode
        self.extra_fields = {}

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

    def _xyxy(self):
        if self.mode == "xyxy":
            return self.bbox.split(1, dim=-1)
        x, y, w, h = self.bbox.split(1, dim=-1)
        return x, y, x + (w - TO_REMOVE).clamp(min=0), y + (h - TO_REMOVE).clamp(min=0)

    def convert(self, mode):
        if mode == self.mode:
            return self
        x1, y1, x2, y2 = self._xyxy()
        if mode == "xyxy":
            out = BoxList(torch.cat((x1, y1, x2, y2), dim=-1), self.size, mode)
        else:
            out = BoxList(torch.cat((x1, y1, x2 - x1 + TO_REMOVE, y2 - y1 + TO_REMOVE), dim=-1), self.size, mode)
        out._copy_extra_fields(self)
        return out

    def resize(self, size, *args, **kwargs):
        ratios = tuple(float(s) / float(o) for s, o in zip(size, self.size))
        if ratios[0] == ratios[1]:
            out = BoxList(self.bbox * ratios[0], size, self.mode)
        else:
            rw, rh = ratios
            x1, y1, x2, y2 = self._xyxy()
            out = BoxList(torch.cat((x1 * rw, y1 * rh, x2 * rw, y2 * rh), dim=-1), size, "xyxy")
        for k, v in self.extra_fields.items():
            if not isinstance(v, torch.Tensor):
                v = v.resize(size, *args, **kwargs)
            out.add_field(k, v)
        return out if ratios[0] == ratios[1] else out.convert(self.mode)

    def to(self, device):
        out = BoxList(self.bbox.to(device), self.size, self.mode)
        for k, v in self.extra_fields.items():
            out.add_field(k, v.to(device) if hasattr(v, "to") else v)
        return out

    def __getitem__(self, item):
        out = BoxList(self.bbox[item], self.size, self.mode)
        for k, v in self.extra_fields.items():
            out.add_field(k, v[item])
        return out

    def __len__(self):
        return self.bbox.shape[0]

    def clip_to_image(self, remove_empty=True):
        self.bbox[:, 0].clamp_(min=0, max=self.size[0] - TO_REMOVE)
        self.bbox[:, 1].clamp_(min=0, max=self.size[1] - TO_REMOVE)
        self.bbox[:, 2].clamp_(min=0, max=self.size[0] - TO_REMOVE)
        self.bbox[:, 3].clamp_(min=0, max=self.size[1] - TO_REMOVE)
        if remove_empty:
            return self[prims.nonempty_mask(self.bbox)]
        return self

    def area(self):
        if self.mode == "xyxy":
            return prims.box_area(self.bbox)
        return self.bbox[:, 2] * self.bbox[:, 3]

    def copy_with_fields(self, fields, skip_missing=False):
        out = BoxList(self.bbox, self.size, self.mode)
        if not isinstance(fields, (list, tuple)):
            fields = [fields]
        for f in fields:
            if self.has_field(f):
                out.add_field(f, self.get_field(f))
            elif not skip_missing:
                raise KeyError(f)
        return out

    def __repr__(self):
        return "BoxList(num_boxes={}, image_width={}, image_height={}, mode={})".format(
            len(self), self.size[0], self.size[1], self.mode)
