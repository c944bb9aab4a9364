"""Cross-checks for the restated upstream (maskrcnn_benchmark) primitives, which the reference
does not pin (it has no tests): each against an independent implementation or a hand-computed case."""
import math

import numpy as np
import torch

from oracle import prims


def _iou_plus1(a, b):
    w = min(a[2], b[2]) - max(a[0], b[0]) + 1
    h = min(a[3], b[3]) - max(a[1], b[1]) + 1
    inter = max(w, 0) * max(h, 0)
    aa = (a[2] - a[0] + 1) * (a[3] - a[1] + 1)
    ab = (b[2] - b[0] + 1) * (b[3] - b[1] + 1)
    return inter / (aa + ab - inter)


def test_nms_matches_bruteforce():
    g = torch.Generator().manual_seed(0)
    for n in (0, 1, 7, 200):
        xy = torch.rand(n, 2, generator=g) * 100
        wh = torch.rand(n, 2, generator=g) * 60 + 1
        boxes = torch.cat([xy, xy + wh], 1)
        scores = torch.rand(n, generator=g)
        keep = prims.nms_legacy(boxes, scores, 0.5).tolist()
        order = sorted(range(n), key=lambda i: (-float(scores[i]), i))
        alive, ref = [True] * n, []
        for a, i in enumerate(order):
            if not alive[a]:
                continue
            ref.append(i)
            for b in range(a + 1, n):
                if alive[b] and _iou_plus1(boxes[i].tolist(), boxes[order[b]].tolist()) > 0.5:
                    alive[b] = False
        assert keep == ref


def test_nms_ties_keep_index_order_and_strict_threshold():
    boxes = torch.tensor([[0., 0., 9., 9.], [0., 0., 9., 9.], [0., 5., 9., 14.]])
    scores = torch.tensor([0.5, 0.5, 0.4])
    assert prims.nms_legacy(boxes, scores, 0.5).tolist() == [0, 2]
    # IoU(+1) of box0 and box2 is exactly 50/150 = 1/3: thresh 1/3 must NOT suppress ('>' rule)
    assert prims.nms_legacy(boxes[[0, 2]], scores[[0, 2]], 1.0 / 3.0 + 1e-7).tolist() == [0, 1]


def test_cell_anchors_known_values():
    # Detectron's canonical stride-16 example: ratios (0.5,1,2) x size 128 (scale 8)
    a = prims.cell_anchors(16, (128,), (0.5, 1.0, 2.0))
    ref = torch.tensor([[-84., -40., 99., 55.], [-56., -56., 71., 71.], [-36., -80., 51., 95.]])
    assert torch.equal(a, ref)
    g = prims.grid_anchors(a, 16, 2, 3)
    assert g.shape == (18, 4)
    assert torch.equal(g[3 * 4 + 1], ref[1] + torch.tensor([16., 16., 16., 16.]))  # (y=1,x=1,a=1)


def test_box_decode_identity_and_clip():
    boxes = torch.tensor([[10., 20., 29., 59.]])
    out = prims.box_decode(torch.zeros(1, 4), boxes, (10., 10., 5., 5.))
    assert torch.allclose(out, boxes)
    big = prims.box_decode(torch.tensor([[0., 0., 100., 100.]]), boxes, (1., 1., 1., 1.))
    assert abs(float(big[0, 2] - big[0, 0] + 1) - 20 * 1000.0 / 16) < 1e-2


def test_level_mapper():
    b = torch.tensor([[0., 0., 59., 149.], [0., 0., 79., 199.], [0., 0., 299., 499.], [0., 0., 39., 99.],
                      [0., 0., 2000., 2000.]])
    assert prims.map_levels(b, 2, 5).tolist() == [0, 1, 2, 0, 3]


def test_roi_align_scalar_vs_torchvision():
    g = torch.Generator().manual_seed(1)
    feat = torch.randn(1, 3, 12, 17, generator=g)
    rois = torch.tensor([[0, 2.0, 3.0, 40.0, 30.0], [0, -20.0, -8.0, 10.0, 90.0], [0, 60.0, 40.0, 62.0, 41.0],
                         [0, 30.0, 10.0, 200.0, 100.0]])
    a = prims.roi_align_legacy(feat, rois, 0.25, 5, 5, 2)
    b = prims.roi_align_scalar(feat, rois, 0.25, 5, 5, 2)
    assert (a - b).abs().max() < 1e-5


def test_frozen_bn_has_no_eps():
    s, b = prims.frozen_bn_scale_bias(torch.tensor([2.0]), torch.tensor([1.0]), torch.tensor([3.0]), torch.tensor([4.0]))
    assert float(s) == 1.0 and float(b) == -2.0
