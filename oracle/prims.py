"""CPU restatement of the upstream (un-vendored) maskrcnn_benchmark primitives.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker.

The SiamMOT hot path calls a handful of primitives that live in
``facebookresearch/maskrcnn-benchmark`` (no version pinned by the reference:
``/root/reference/readme/INSTALL.md:89-105``; the source is NOT under
``/root/reference``).  They are restated here from the published algorithm,
and every restatement names the reference call site that depends on it.
**Parity for these is unpinned by the reference** (it ships no tests / golden
vectors); they are cross-checked against independent implementations in
``tests/test_oracle_prims.py`` (torchvision ``roi_align(aligned=False)``,
brute-force NMS, hand-computed anchor/box-coder cases).

All arithmetic is fp32 on CPU tensors, like the reference's default
``cfg.DTYPE = "float32"``.
"""
import math

import numpy as np
import torch

TO_REMOVE = 1.0  # legacy "+1" pixel convention of maskrcnn_benchmark BoxList
BBOX_XFORM_CLIP = math.log(1000.0 / 16)


# ----------------------------------------------------------------------------
# boxes
# ----------------------------------------------------------------------------
def box_area(b):
    """BoxList.area() for xyxy boxes (used by LevelMapper; sr_pool.py:74)."""
    return (b[:, 2] - b[:, 0] + TO_REMOVE) * (b[:, 3] - b[:, 1] + TO_REMOVE)


def clip_boxes(b, w, h):
    """BoxList.clip_to_image (rpn_patch.py:51, inference.py:110, track_core.py:178):
    clamp to [0, W-1] x [0, H-1]. Returns a new tensor."""
    b = b.clone()
    b[:, 0].clamp_(min=0, max=w - TO_REMOVE)
    b[:, 1].clamp_(min=0, max=h - TO_REMOVE)
    b[:, 2].clamp_(min=0, max=w - TO_REMOVE)
    b[:, 3].clamp_(min=0, max=h - TO_REMOVE)
    return b


def nonempty_mask(b):
    """remove_empty=True rule of clip_to_image."""
    return (b[:, 3] > b[:, 1]) & (b[:, 2] > b[:, 0])


def box_decode(rel_codes, boxes, weights):
    """BoxCoder.decode (inference.py:69 weights (10,10,5,5); rpn_patch.py:38 weights 1)."""
    boxes = boxes.to(rel_codes.dtype)
    widths = boxes[:, 2] - boxes[:, 0] + TO_REMOVE
    heights = boxes[:, 3] - boxes[:, 1] + TO_REMOVE
    ctr_x = boxes[:, 0] + 0.5 * widths
    ctr_y = boxes[:, 1] + 0.5 * heights
    wx, wy, ww, wh = weights
    dx = rel_codes[:, 0::4] / wx
    dy = rel_codes[:, 1::4] / wy
    dw = rel_codes[:, 2::4] / ww
    dh = rel_codes[:, 3::4] / wh
    dw = torch.clamp(dw, max=BBOX_XFORM_CLIP)
    dh = torch.clamp(dh, max=BBOX_XFORM_CLIP)
    pcx = dx * widths[:, None] + ctr_x[:, None]
    pcy = dy * heights[:, None] + ctr_y[:, None]
    pw = torch.exp(dw) * widths[:, None]
    ph = torch.exp(dh) * heights[:, None]
    out = torch.zeros_like(rel_codes)
    out[:, 0::4] = pcx - 0.5 * pw
    out[:, 1::4] = pcy - 0.5 * ph
    out[:, 2::4] = pcx + 0.5 * pw - 1
    out[:, 3::4] = pcy + 0.5 * ph - 1
    return out


def nms_legacy(boxes, scores, thresh):
    """``_C.nms`` CUDA semantics (used by boxlist_nms: rpn_patch.py:53,
    inference.py:174, track_solver.py:22): sort by score descending (stable, so
    ties keep index order), IoU with +1 widths, suppress when IoU > thresh.
    Returns kept indices into the input, in descending-score order (int64)."""
    n = boxes.shape[0]
    if n == 0:
        return torch.zeros((0,), dtype=torch.int64)
    order = torch.sort(scores, descending=True, stable=True)[1]
    b = boxes[order].to(torch.float32).numpy()
    x1, y1, x2, y2 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    area = (x2 - x1 + np.float32(1)) * (y2 - y1 + np.float32(1))
    dead = np.zeros(n, dtype=bool)
    keep = []
    thr = np.float32(thresh)
    for i in range(n):
        if dead[i]:
            continue
        keep.append(i)
        if i + 1 == n:
            break
        xx1 = np.maximum(x1[i], x1[i + 1:])
        yy1 = np.maximum(y1[i], y1[i + 1:])
        xx2 = np.minimum(x2[i], x2[i + 1:])
        yy2 = np.minimum(y2[i], y2[i + 1:])
        w = np.maximum(xx2 - xx1 + np.float32(1), np.float32(0))
        h = np.maximum(yy2 - yy1 + np.float32(1), np.float32(0))
        inter = w * h
        iou = inter / (area[i] + area[i + 1:] - inter)
        dead[i + 1:] |= iou > thr
    return order[torch.as_tensor(keep, dtype=torch.int64)]


def remove_small_mask(boxes, min_size):
    """remove_small_boxes (rpn_patch.py:52): (w+1)>=min and (h+1)>=min."""
    ws = boxes[:, 2] - boxes[:, 0] + TO_REMOVE
    hs = boxes[:, 3] - boxes[:, 1] + TO_REMOVE
    return (ws >= min_size) & (hs >= min_size)


# ----------------------------------------------------------------------------
# anchors (Detectron legacy rounding; upstream rpn/anchor_generator.py)
# ----------------------------------------------------------------------------
def _whctrs(a):
    w = a[2] - a[0] + 1
    h = a[3] - a[1] + 1
    return w, h, a[0] + 0.5 * (w - 1), a[1] + 0.5 * (h - 1)


def _mk(ws, hs, xc, yc):
    ws = ws[:, None]
    hs = hs[:, None]
    return np.hstack((xc - 0.5 * (ws - 1), yc - 0.5 * (hs - 1),
                      xc + 0.5 * (ws - 1), yc + 0.5 * (hs - 1)))


def cell_anchors(stride, sizes, aspect_ratios):
    """generate_anchors(stride, sizes, ratios): float64 numpy maths then .float()."""
    scales = np.array(sizes, dtype=np.float64) / stride
    ratios = np.array(aspect_ratios, dtype=np.float64)
    base = np.array([1, 1, stride, stride], dtype=np.float64) - 1
    w, h, xc, yc = _whctrs(base)
    size = w * h
    ws = np.round(np.sqrt(size / ratios))
    hs = np.round(ws * ratios)
    ratio_anchors = _mk(ws, hs, xc, yc)
    out = []
    for i in range(ratio_anchors.shape[0]):
        w, h, xc, yc = _whctrs(ratio_anchors[i])
        out.append(_mk(w * scales, h * scales, xc, yc))
    return torch.from_numpy(np.vstack(out)).float()


def grid_anchors(cell, stride, gh, gw):
    """AnchorGenerator.grid_anchors: index = (y*gw + x)*A + a."""
    sx = torch.arange(0, gw * stride, step=stride, dtype=torch.float32)
    sy = torch.arange(0, gh * stride, step=stride, dtype=torch.float32)
    yy, xx = torch.meshgrid(sy, sx, indexing="ij")
    xx = xx.reshape(-1)
    yy = yy.reshape(-1)
    shifts = torch.stack((xx, yy, xx, yy), dim=1)
    return (shifts.view(-1, 1, 4) + cell.view(1, -1, 4)).reshape(-1, 4)


# ----------------------------------------------------------------------------
# FPN level mapping and legacy ROIAlign
# ----------------------------------------------------------------------------
def map_levels(boxes, k_min, k_max, canonical_scale=224, canonical_level=4, eps=1e-6):
    """LevelMapper.__call__ (sr_pool.py:38,74; upstream poolers.py)."""
    s = torch.sqrt(box_area(boxes))
    lvl = torch.floor(canonical_level + torch.log2(s / canonical_scale + eps))
    lvl = torch.clamp(lvl, min=k_min, max=k_max)
    return lvl.to(torch.int64) - int(k_min)


def roi_align_legacy(feat, rois, spatial_scale, ph, pw, sampling_ratio):
    """``_C.roi_align_forward`` (csrc/cuda/ROIAlign_cuda.cu semantics; sr_pool.py:28,89).
    feat (N,C,H,W) f32, rois (K,5) [batch, x1,y1,x2,y2].  Same function as
    torchvision.ops.roi_align(aligned=False), which is what runs here; an
    independent scalar restatement lives in ``roi_align_scalar`` below."""
    from torchvision.ops import roi_align
    if rois.shape[0] == 0:
        return feat.new_zeros((0, feat.shape[1], ph, pw))
    return roi_align(feat, rois.to(feat.dtype), (ph, pw), spatial_scale=spatial_scale,
                     sampling_ratio=sampling_ratio, aligned=False)


def _bilinear(fm, y, x):
    C, H, W = fm.shape
    if y < -1.0 or y > H or x < -1.0 or x > W:
        return np.zeros(C, dtype=np.float32)
    y = max(y, np.float32(0))
    x = max(x, np.float32(0))
    yl = int(y)
    xl = int(x)
    if yl >= H - 1:
        yh = yl = H - 1
        y = np.float32(yl)
    else:
        yh = yl + 1
    if xl >= W - 1:
        xh = xl = W - 1
        x = np.float32(xl)
    else:
        xh = xl + 1
    ly = np.float32(y - yl)
    lx = np.float32(x - xl)
    hy = np.float32(1) - ly
    hx = np.float32(1) - lx
    return (hy * hx * fm[:, yl, xl] + hy * lx * fm[:, yl, xh] +
            ly * hx * fm[:, yh, xl] + ly * lx * fm[:, yh, xh])


def roi_align_scalar(feat, rois, spatial_scale, ph, pw, sampling_ratio):
    """Scalar restatement of RoIAlignForward (legacy, non-aligned).  Slow; small cases only."""
    f = feat.numpy().astype(np.float32)
    K = rois.shape[0]
    out = np.zeros((K, f.shape[1], ph, pw), dtype=np.float32)
    for k in range(K):
        r = rois[k].numpy().astype(np.float32)
        bi = int(r[0])
        sc = np.float32(spatial_scale)
        x1, y1, x2, y2 = r[1] * sc, r[2] * sc, r[3] * sc, r[4] * sc
        rw = max(x2 - x1, np.float32(1))
        rh = max(y2 - y1, np.float32(1))
        bh = np.float32(rh / np.float32(ph))
        bw = np.float32(rw / np.float32(pw))
        gh = sampling_ratio if sampling_ratio > 0 else int(math.ceil(rh / ph))
        gw = sampling_ratio if sampling_ratio > 0 else int(math.ceil(rw / pw))
        for i in range(ph):
            for j in range(pw):
                acc = np.zeros(f.shape[1], dtype=np.float32)
                for iy in range(gh):
                    y = np.float32(y1 + np.float32(i) * bh + np.float32(iy + 0.5) * bh / np.float32(gh))
                    for ix in range(gw):
                        x = np.float32(x1 + np.float32(j) * bw + np.float32(ix + 0.5) * bw / np.float32(gw))
                        acc += _bilinear(f[bi], y, x)
                out[k, :, i, j] = acc / np.float32(gh * gw)
    return torch.from_numpy(out)


def frozen_bn_scale_bias(weight, bias, running_mean, running_var):
    """FrozenBatchNorm2d (dla.py:9,32): scale = w*rsqrt(var) (no eps), bias = b - mean*scale."""
    scale = weight * running_var.rsqrt()
    return scale, bias - running_mean * scale
