"""CPU fp32 oracle for the SiamMOT per-frame inference path (plain PyTorch, functional).

TEST INFRASTRUCTURE ONLY (see oracle/prims.py header): imported by tests/, by
__graft_entry__.smoke() and by bench.py's CPU-baseline / --impl reference legs, and only as the
checker.  The product (siammot_b200/) never imports this package.

This is a *restatement* of the reference algorithm, written from scratch around plain tensors
(no BoxList, no nn.Module tree), each stage citing the reference lines it follows.  It is pinned
two ways (tests/test_oracle_golden.py):
  * against golden vectors produced by the reference's OWN modules run unmodified from
    /root/reference over the maskrcnn_benchmark stand-in (tests/golden/make_golden.py);
  * live against that same reference when /root/reference is present.
Upstream maskrcnn_benchmark primitives (absent from /root/reference) are restated in prims.py;
for those, parity is unpinned by the reference and cross-checked independently.

Parameters are a flat dict keyed exactly like the reference ``state_dict()``.
"""
import copy
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import prims


# ----------------------------------------------------------------------------------------------
# backbone: DLA-34 (dla.py:241-313) + patched FPN (fpn_patch.py:29-61)
# ----------------------------------------------------------------------------------------------
def _conv_bn(P, x, conv, bn, stride=1, pad=0, relu=False, residual=None):
    """Conv2d(bias=False) -> FrozenBatchNorm2d [-> += residual] [-> ReLU]  (dla.py:43-57,181-189)."""
    y = F.conv2d(x, P[conv + ".weight"], None, stride, pad)
    scale, bias = prims.frozen_bn_scale_bias(P[bn + ".weight"], P[bn + ".bias"],
                                             P[bn + ".running_mean"], P[bn + ".running_var"])
    y = y * scale.reshape(1, -1, 1, 1) + bias.reshape(1, -1, 1, 1)
    if residual is not None:
        y = y + residual
    return F.relu(y) if relu else y


def _basic_block(P, pre, x, stride, residual=None):
    """DlaBasic.forward (dla.py:43-57)."""
    if residual is None:
        residual = x
    y = _conv_bn(P, x, pre + ".conv1", pre + ".bn1", stride, 1, relu=True)
    return _conv_bn(P, y, pre + ".conv2", pre + ".bn2", 1, 1, relu=True, residual=residual)


def deform_conv3x3(x, offset, weight, stride):
    """DCN v1 (upstream layers/dcn DeformConv, reached through DFConv2d at dla.py:74-78), 3x3 / pad 1, restated from the published
    operator: output pixel (oy, ox), tap k = 3i + j samples the input at (oy*stride - 1 + i + dy, ox*stride - 1 + j + dx) with
    (dy, dx) = offset channels (2k, 2k+1), bilinearly, corners outside the map contributing zero.  Cross-checked against
    torchvision.ops.deform_conv2d in tests/test_oracle_dla_family_cpu.py."""
    B, C, H, W = x.shape
    assert B == 1
    OH, OW = offset.shape[2], offset.shape[3]
    oy = torch.arange(OH, dtype=torch.float32).view(OH, 1) * stride - 1
    ox = torch.arange(OW, dtype=torch.float32).view(1, OW) * stride - 1
    cols = []
    img = x[0]
    for k in range(9):
        i, j = divmod(k, 3)
        yy = oy + i + offset[0, 2 * k]
        xx = ox + j + offset[0, 2 * k + 1]
        inside = (yy > -1) & (yy < H) & (xx > -1) & (xx < W)
        y0, x0 = torch.floor(yy), torch.floor(xx)
        ly, lx = yy - y0, xx - x0
        acc = torch.zeros((C, OH, OW))
        for dy_, dx_, wgt in ((0, 0, (1 - ly) * (1 - lx)), (0, 1, (1 - ly) * lx), (1, 0, ly * (1 - lx)), (1, 1, ly * lx)):
            yi, xi = (y0 + dy_).long(), (x0 + dx_).long()
            ok = inside & (yi >= 0) & (yi < H) & (xi >= 0) & (xi < W)
            v = img[:, yi.clamp(0, H - 1), xi.clamp(0, W - 1)]
            acc = acc + v * (wgt * ok)[None]
        cols.append(acc)
    col = torch.stack(cols, 1)                                   # (C, 9, OH, OW)
    return torch.einsum("ock,ckhw->ohw", weight.reshape(weight.shape[0], C, 9), col)[None]


def _bottleneck_block(P, pre, x, stride, residual=None):
    """DlaBottleneck.forward (dla.py:81-101), cardinality 1 / base width 64: 1x1 -> 3x3 (stride) -> 1x1, + residual, ReLU.
    With MODEL.DLA.STAGE_WITH_DCN the 3x3 is upstream's DFConv2d (dla.py:74-78): offsets from a regular 3x3 conv with bias."""
    if residual is None:
        residual = x
    y = _conv_bn(P, x, pre + ".conv1", pre + ".bn1", 1, 0, relu=True)
    if (pre + ".conv2.offset.weight") in P:
        off = F.conv2d(y, P[pre + ".conv2.offset.weight"], P[pre + ".conv2.offset.bias"], stride, 1)
        y = deform_conv3x3(y, off, P[pre + ".conv2.conv.weight"], stride)
        scale, bias = prims.frozen_bn_scale_bias(P[pre + ".bn2.weight"], P[pre + ".bn2.bias"], P[pre + ".bn2.running_mean"],
                                                 P[pre + ".bn2.running_var"])
        y = F.relu(y * scale.reshape(1, -1, 1, 1) + bias.reshape(1, -1, 1, 1))
    else:
        y = _conv_bn(P, y, pre + ".conv2", pre + ".bn2", stride, 1, relu=True)
    return _conv_bn(P, y, pre + ".conv3", pre + ".bn3", 1, 0, relu=True, residual=residual)


def _tree(P, pre, x, levels, cin, cout, stride, level_root, children=None, block=None, root_residual=False):
    """DlaTree.forward (dla.py:225-238).  The ``residual`` argument of the reference is always
    overwritten at dla.py:228, so it is not a parameter here.  ``block``: _basic_block (DLA-34) or _bottleneck_block;
    ``root_residual``: DlaRoot adds its first input (dla.py:185-186; DLA-102 / 169)."""
    block = block or _basic_block
    children = [] if children is None else children
    bottom = F.max_pool2d(x, stride, stride) if stride > 1 else x
    if cin != cout:
        residual = _conv_bn(P, bottom, pre + ".project.0", pre + ".project.1")
    else:
        residual = bottom
    if level_root:
        children.append(bottom)
    if levels == 1:
        x1 = block(P, pre + ".tree1", x, stride, residual)
        x2 = block(P, pre + ".tree2", x1, 1)
        cat = torch.cat([x2, x1] + children, 1)           # DlaRoot.forward dla.py:181-189
        return _conv_bn(P, cat, pre + ".root.conv", pre + ".root.bn", relu=True, residual=x2 if root_residual else None)
    x1 = _tree(P, pre + ".tree1", x, levels - 1, cin, cout, stride, False, None, block, root_residual)
    children.append(x1)
    return _tree(P, pre + ".tree2", x1, levels - 1, cout, cout, 1, False, children, block, root_residual)


def dla_forward(P, x, arch, pre="backbone.body"):
    """DLA.forward (dla.py:289-304) for the plain members of the family (dla.py:307-372): levels / channels / block /
    residual_root from siammot_b200.synthetic.DLA_ARCHS (the table is data about the reference, shared with the test inputs)."""
    from siammot_b200.synthetic import DLA_ARCHS
    A = DLA_ARCHS[arch]
    ch, lv = A["channels"], A["levels"]
    block = _bottleneck_block if A["block"] == "bottleneck" else _basic_block
    x = _conv_bn(P, x, pre + ".base_layer.0", pre + ".base_layer.1", 1, 3, relu=True)
    for name, n, stride in (("level0", lv[0], 1), ("level1", lv[1], 2)):      # _make_conv_level dla.py:278-287
        for i in range(n):
            x = _conv_bn(P, x, "%s.%s.%d" % (pre, name, 3 * i), "%s.%s.%d" % (pre, name, 3 * i + 1), stride if i == 0 else 1, 1, relu=True)
    outs = []
    for lvl in range(2, 6):
        x = _tree(P, "%s.level%d" % (pre, lvl), x, lv[lvl], ch[lvl - 1], ch[lvl], 2, lvl > 2, None, block, A["residual_root"])
        outs.append(x)
    return outs


DLA34_LEVELS = (1, 1, 1, 2, 2, 1)
DLA34_CHANNELS = (16, 32, 64, 128, 256, 512)


def dla34_forward(P, x, pre="backbone.body"):
    """DLA.forward (dla.py:289-304) for dla_34 (dla.py:307-313)."""
    ch = DLA34_CHANNELS
    x = _conv_bn(P, x, pre + ".base_layer.0", pre + ".base_layer.1", 1, 3, relu=True)
    x = _conv_bn(P, x, pre + ".level0.0", pre + ".level0.1", 1, 1, relu=True)
    x = _conv_bn(P, x, pre + ".level1.0", pre + ".level1.1", 2, 1, relu=True)
    x2 = _tree(P, pre + ".level2", x, DLA34_LEVELS[2], ch[1], ch[2], 2, False)
    x3 = _tree(P, pre + ".level3", x2, DLA34_LEVELS[3], ch[2], ch[3], 2, True)
    x4 = _tree(P, pre + ".level4", x3, DLA34_LEVELS[4], ch[3], ch[4], 2, True)
    x5 = _tree(P, pre + ".level5", x4, DLA34_LEVELS[5], ch[4], ch[5], 2, True)
    return [x2, x3, x4, x5]


R50_BLOCKS = (3, 4, 6, 3)


def resnet50_forward(P, x, pre="backbone.body", stride_in_1x1=True, blocks=R50_BLOCKS):
    """Upstream maskrcnn_benchmark ResNet.forward for "R-50-FPN" (modeling/backbone/resnet.py; un-vendored, restated):
    stem = 7x7/2 conv + FrozenBN + ReLU + 3x3/2 max-pool (pad 1); four stages of bottleneck blocks
    1x1 (stride here when STRIDE_IN_1X1) -> 3x3 -> 1x1, FrozenBN after each, ReLU after the first two and after the
    residual add; the first block of a stage projects the identity with a strided 1x1 + FrozenBN.  Returns C2..C5."""
    x = _conv_bn(P, x, pre + ".stem.conv1", pre + ".stem.bn1", 2, 3, relu=True)
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    outs = []
    for li, nb in enumerate(blocks):
        for b in range(nb):
            blk = "%s.layer%d.%d" % (pre, li + 1, b)
            stride = 2 if (b == 0 and li > 0) else 1
            s1, s3 = (stride, 1) if stride_in_1x1 else (1, stride)
            identity = x
            if (blk + ".downsample.0.weight") in P:
                identity = _conv_bn(P, x, blk + ".downsample.0", blk + ".downsample.1", stride, 0)
            y = _conv_bn(P, x, blk + ".conv1", blk + ".bn1", s1, 0, relu=True)
            y = _conv_bn(P, y, blk + ".conv2", blk + ".bn2", s3, 1, relu=True)
            x = _conv_bn(P, y, blk + ".conv3", blk + ".bn3", 1, 0, relu=True, residual=identity)
        outs.append(x)
    return outs


def fpn_forward(P, feats, pre="backbone.fpn"):
    """Patched FPN.forward (fpn_patch.py:29-61): bilinear resize-to-lateral-size top-down,
    LastLevelMaxPool = stride-2 subsample for P6."""
    def conv(name, t, pad):
        return F.conv2d(t, P[pre + "." + name + ".weight"], P[pre + "." + name + ".bias"], 1, pad)
    n = len(feats)
    last = conv("fpn_inner%d" % n, feats[-1], 0)
    results = [conv("fpn_layer%d" % n, last, 1)]
    for i in range(n - 1, 0, -1):
        lateral = conv("fpn_inner%d" % i, feats[i - 1], 0)
        top = F.interpolate(last, size=lateral.shape[-2:], mode="bilinear", align_corners=False)
        last = lateral + top
        results.insert(0, conv("fpn_layer%d" % i, last, 1))
    results.append(F.max_pool2d(results[-1], 1, 2, 0))
    return results


# ----------------------------------------------------------------------------------------------
# RPN (upstream head + anchors; patched post-processor rpn_patch.py:15-60)
# ----------------------------------------------------------------------------------------------
def rpn_head_forward(P, feats):
    """Upstream SingleConvRPNHead: 3x3 conv + ReLU, 1x1 -> A objectness, 1x1 -> 4A deltas, shared over levels."""
    logits, deltas = [], []
    for f in feats:
        t = F.relu(F.conv2d(f, P["rpn.head.conv.weight"], P["rpn.head.conv.bias"], 1, 1))
        logits.append(F.conv2d(t, P["rpn.head.cls_logits.weight"], P["rpn.head.cls_logits.bias"]))
        deltas.append(F.conv2d(t, P["rpn.head.bbox_pred.weight"], P["rpn.head.bbox_pred.bias"]))
    return logits, deltas


def rpn_select(cfg, logits_l, deltas_l, img_w, img_h):
    """Patched RPNPostProcessor (rpn_patch.py:15-60) + upstream select_over_all_levels.
    logits_l[i]: (1,A,H,W), deltas_l[i]: (1,4A,H,W).  Returns proposals (R,4), objectness (R,)."""
    R = cfg.MODEL.RPN
    boxes_all, scores_all = [], []
    for lvl, (logits, deltas) in enumerate(zip(logits_l, deltas_l)):
        _, A, H, W = logits.shape
        # permute_and_flatten -> order (h, w, a)                         rpn_patch.py:21-24
        logit = logits[0].permute(1, 2, 0).reshape(-1)
        reg = deltas[0].view(A, 4, H, W).permute(2, 3, 0, 1).reshape(-1, 4)
        k = min(R.PRE_NMS_TOP_N_TEST, logit.numel())
        # rpn_patch.py:28-29 is topk(sigmoid(logit), sorted=True), whose order among equal fp32
        # sigmoid values is implementation-defined.  The oracle fixes it: sigmoid is monotonic, so a
        # stable descending sort of the logits is one admissible order (ties -> lowest anchor index).
        idx = torch.sort(logit, descending=True, stable=True)[1][:k]
        obj = logit[idx].sigmoid()
        cell = prims.cell_anchors(R.ANCHOR_STRIDE[lvl], (R.ANCHOR_SIZES[lvl],), R.ASPECT_RATIOS)
        anchors = prims.grid_anchors(cell, R.ANCHOR_STRIDE[lvl], H, W)[idx]
        prop = prims.box_decode(reg[idx], anchors, (1.0, 1.0, 1.0, 1.0))   # rpn_patch.py:38
        if not cfg.INPUT.AMODAL:
            prop = prims.clip_boxes(prop, img_w, img_h)                   # rpn_patch.py:49-51
        keep = prims.remove_small_mask(prop, R.MIN_SIZE).nonzero().squeeze(1)
        prop, obj = prop[keep], obj[keep]
        keep = prims.nms_legacy(prop, obj, R.NMS_THRESH)[:R.POST_NMS_TOP_N_TEST]
        boxes_all.append(prop[keep])
        scores_all.append(obj[keep])
    boxes = torch.cat(boxes_all)
    scores = torch.cat(scores_all)
    k = min(R.FPN_POST_NMS_TOP_N_TEST, scores.numel())                    # select_over_all_levels
    inds = torch.sort(scores, descending=True, stable=True)[1][:k]       # topk, ties -> first
    return boxes[inds], scores[inds]


def rpn_forward(P, cfg, feats, img_w, img_h):
    """Returns proposals (R,4) and objectness (R,), R <= FPN_POST_NMS_TOP_N_TEST."""
    logits, deltas = rpn_head_forward(P, feats)
    return rpn_select(cfg, logits, deltas, img_w, img_h)


# ----------------------------------------------------------------------------------------------
# box head (box_head.py:23-52, inference.py:46-191)
# ----------------------------------------------------------------------------------------------
def pool_rois(feats, boxes, level_boxes, scales, res, sampling, rois=None):
    """Pooler / SRPooler.forward (sr_pool.py:53-91): level from ``level_boxes``, ROI = ``rois``
    (defaults to boxes).  Only the first len(scales) maps are used (zip-truncation sr_pool.py:86)."""
    rois = boxes if rois is None else rois
    n = rois.shape[0]
    out = torch.zeros((n, feats[0].shape[1], res, res), dtype=feats[0].dtype)
    if n == 0:
        return out
    k_min = -math.log2(scales[0])
    k_max = -math.log2(scales[-1])
    levels = prims.map_levels(level_boxes, k_min, k_max)
    r5 = torch.cat([torch.zeros((n, 1)), rois], dim=1)
    for lvl, sc in enumerate(scales):
        idx = torch.nonzero(levels == lvl).squeeze(1)
        out[idx] = prims.roi_align_legacy(feats[lvl], r5[idx], sc, res, res, sampling)
    return out


def box_head_features(P, cfg, feats, boxes):
    """FPN2MLPFeatureExtractor + FPNPredictor (upstream; box_head.py:46-48): logits (n,ncls), deltas (n,4*ncls)."""
    H = cfg.MODEL.ROI_BOX_HEAD
    pre = "roi_heads.box."
    x = pool_rois(feats, boxes, boxes, H.POOLER_SCALES, H.POOLER_RESOLUTION, H.POOLER_SAMPLING_RATIO)
    x = x.reshape(x.shape[0], -1)
    x = F.relu(F.linear(x, P[pre + "feature_extractor.fc6.weight"], P[pre + "feature_extractor.fc6.bias"]))
    x = F.relu(F.linear(x, P[pre + "feature_extractor.fc7.weight"], P[pre + "feature_extractor.fc7.bias"]))
    logits = F.linear(x, P[pre + "predictor.cls_score.weight"], P[pre + "predictor.cls_score.bias"])
    deltas = F.linear(x, P[pre + "predictor.bbox_pred.weight"], P[pre + "predictor.bbox_pred.bias"])
    return logits, deltas


def box_post(cfg, logits, deltas, boxes, img_w, img_h, ids=None, labels=None):
    """PostProcessor.forward + filter_results (inference.py:46-191).  ``ids``/``labels`` given => the
    boxes are tracks (inference.py:80-103).  Returns dict(boxes, scores, ids, labels)."""
    prob = F.softmax(logits, -1)
    n, ncls = prob.shape
    if cfg.MODEL.CLS_AGNOSTIC_BBOX_REG:
        deltas = deltas[:, -4:]
    dec = prims.box_decode(deltas, boxes, cfg.MODEL.ROI_HEADS.BBOX_REG_WEIGHTS)   # inference.py:69
    if cfg.MODEL.CLS_AGNOSTIC_BBOX_REG:
        dec = dec.repeat(1, ncls)
    if ids is None:
        ids = torch.full((n,), -1, dtype=torch.int64)
    if labels is not None:                                                # inference.py:93-103
        trk = (ids >= 0).nonzero().squeeze(1)
        if trk.numel() > 0:
            assert trk.numel() == n, "reference indexing prob[track_inds, labels] needs all-track input"
            cp = prob.clone()
            prob[trk, :] = 0.0
            prob[trk, labels] = cp[trk, labels] + 1.0
    dec = dec.reshape(-1, 4)
    if not cfg.INPUT.AMODAL:
        dec = prims.clip_boxes(dec, img_w, img_h)                         # inference.py:109-110
    dec = dec.reshape(n, ncls * 4)
    ob, os_, oi, ol = [], [], [], []
    for j in range(1, ncls):                                              # filter_results :145-191
        inds = (prob[:, j] > cfg.MODEL.ROI_HEADS.SCORE_THRESH).nonzero().squeeze(1)
        sj, bj, ij = prob[inds, j], dec[inds, j * 4:(j + 1) * 4], ids[inds]
        d = ij < 0
        keep = prims.nms_legacy(bj[d], sj[d], cfg.MODEL.ROI_HEADS.NMS)
        t = ij >= 0
        ob += [bj[d][keep], bj[t]]
        os_ += [sj[d][keep], sj[t]]
        oi += [ij[d][keep], ij[t]]
        ol.append(torch.full((keep.numel() + int(t.sum()),), j, dtype=torch.int64))
    return dict(boxes=torch.cat(ob), scores=torch.cat(os_), ids=torch.cat(oi), labels=torch.cat(ol))


def box_head_forward(P, cfg, feats, boxes, img_w, img_h, ids=None, labels=None):
    """ROIBoxHead.forward at inference (box_head.py:23-52)."""
    logits, deltas = box_head_features(P, cfg, feats, boxes)
    return box_post(cfg, logits, deltas, boxes, img_w, img_h, ids, labels)


# ----------------------------------------------------------------------------------------------
# EMM tracker (track_core.py, feature_extractor.py, xcorr.py, track_utils.py)
# ----------------------------------------------------------------------------------------------
def xcorr_depthwise(x, k):
    """xcorr.py:37-45: per (track, channel) valid cross-correlation."""
    n, c = k.shape[:2]
    out = F.conv2d(x.reshape(1, n * c, x.shape[2], x.shape[3]),
                   k.reshape(n * c, 1, k.shape[2], k.shape[3]), groups=n * c)
    return out.reshape(n, c, out.shape[2], out.shape[3])


def emm_predictor(P, resp, pre="roi_heads.track.tracker.predictor."):
    """EMMPredictor.forward (feature_extractor.py:62-69)."""
    def tower(name):
        y = F.conv2d(resp, P[pre + name + ".0.weight"], None, 1, 1)
        y = F.group_norm(y, 32, P[pre + name + ".1.weight"], P[pre + name + ".1.bias"], 1e-5)
        return F.relu(y)
    ct, rt = tower("cls_tower"), tower("reg_tower")
    cls = F.conv2d(ct, P[pre + "cls.weight"], P[pre + "cls.bias"], 1, 1)
    ctr = F.conv2d(ct, P[pre + "center.weight"], P[pre + "center.bias"], 1, 1)
    reg = F.relu(F.conv2d(rt, P[pre + "reg.weight"], P[pre + "reg.bias"], 1, 1))
    return cls, ctr, reg


def emm_decode(cls, ctr, reg, sr, tboxes, pad, tmpl_res, use_centerness, sigma, up=16):
    """bicubic x16 (track_core.py:69-71) + get_locations (:184-225) + decode_response (:101-135)."""
    cls = F.interpolate(cls, scale_factor=up, mode="bicubic")
    ctr = F.interpolate(ctr, scale_factor=up, mode="bicubic")
    reg = F.interpolate(reg, scale_factor=up, mode="bicubic")
    n = cls.shape[0]
    # locations: the SR box sampled on a (S*up)^2 grid, inner (S-2*border)*up kept    :190-209
    s_full = (cls.shape[-1] // up + 2 * (tmpl_res // 2)) * up
    border = (tmpl_res // 2) * up
    bw = sr[:, 2] - sr[:, 0]
    bh = sr[:, 3] - sr[:, 1]
    ar = torch.arange(0, s_full, dtype=torch.float32)
    gx = (sr[:, 0:1] + ar[None, :] * (bw / (s_full - 1))[:, None])[:, border:-border] - pad
    gy = (sr[:, 1:2] + ar[None, :] * (bh / (s_full - 1))[:, None])[:, border:-border] - pad
    # confidence map                                                                  :101-118
    p1 = F.softmax(cls, dim=1)[:, 1].reshape(n, -1)
    conf = p1 * torch.sigmoid(ctr).reshape(n, -1) if use_centerness else p1
    tlbr = reg.reshape(n, 4, -1)
    sw = (tlbr[:, 2] + tlbr[:, 0]) / (tboxes[:, 2] - tboxes[:, 0])[:, None]          # :138-152
    sh = (tlbr[:, 3] + tlbr[:, 1]) / (tboxes[:, 3] - tboxes[:, 1])[:, None]
    sw = torch.max(sw, 1 / sw)
    sh = torch.max(sh, 1 / sh)
    penalty = torch.exp((-sw * sh + 1) * 0.1)
    side = int(np.sqrt(tlbr.shape[-1]))
    hann = torch.hann_window(side, dtype=torch.float)                                  # :155-162
    window = torch.outer(hann, hann).reshape(-1)[None]
    score = (conf * penalty) * (1 - sigma) + sigma * window
    idx = torch.argmax(score, dim=1)                                                   # :120
    ar_n = torch.arange(n)
    iy, ix = idx // side, idx % side
    cx, cy = gx[ar_n, ix], gy[ar_n, iy]
    d = tlbr[ar_n, :, idx]
    bb = torch.stack((cx - d[:, 0], cy - d[:, 1], cx + d[:, 2], cy + d[:, 3]), dim=1)
    return bb, p1[ar_n, idx]                                                           # :132-133


def pad_features(feats, pad):
    """TrackUtils.pad_feature (track_utils.py:87-107)."""
    return [F.pad(f, [int(pad / ((2 ** i) * 4))] * 4) for i, f in enumerate(feats)]


def emm_forward(P, cfg, feats, memory, img_w, img_h):
    """EMM.forward inference branch (track_core.py:28-79).  memory = dict(feat, sr, boxes, ids, labels)."""
    T = cfg.MODEL.TRACK_HEAD
    res = T.POOLER_RESOLUTION
    sres = int(res * T.SEARCH_REGION)
    padded = pad_features(feats, T.PAD_PIXELS)
    srf = pool_rois(padded, memory["boxes"], memory["boxes"], T.POOLER_SCALES, sres,
                    T.POOLER_SAMPLING_RATIO, rois=memory["sr"])
    resp = xcorr_depthwise(srf, memory["feat"])
    cls, ctr, reg = emm_predictor(P, resp)
    bb, conf = emm_decode(cls, ctr, reg, memory["sr"], memory["boxes"], T.PAD_PIXELS, res,
                          T.EMM.USE_CENTERNESS, T.EMM.COSINE_WINDOW_WEIGHT)
    ids, labels = memory["ids"], memory["labels"]
    if not cfg.INPUT.AMODAL:                                              # track_core.py:177-178
        bb = prims.clip_boxes(bb, img_w, img_h)
        m = prims.nonempty_mask(bb)
        bb, conf, ids, labels = bb[m], conf[m], ids[m], labels[m]
    return dict(boxes=bb, scores=conf, ids=ids, labels=labels)


def search_region(boxes, pad, expansion, min_wh):
    """update_boxes_in_pad_images + extend_bbox (track_utils.py:62-85,109-135)."""
    sr = boxes + pad
    w = sr[:, 2] - sr[:, 0] + 1
    h = sr[:, 3] - sr[:, 1] + 1
    we = torch.max((min_wh - w) / (expansion * 2.0), w * (expansion / 2.0))
    he = torch.max((min_wh - h) / (expansion * 2.0), h * (expansion / 2.0))
    return torch.stack((sr[:, 0] - we, sr[:, 1] - he, sr[:, 2] + we, sr[:, 3] + he), dim=1)


# ----------------------------------------------------------------------------------------------
# track pool + solver + memory (track_utils.py:138-255, track_solver.py, track_head.py:54-110)
# ----------------------------------------------------------------------------------------------
class PoolState(object):
    """TrackPool: id allocator, active set, dormant table id -> last active frame, per-id cache.
    Container types mirror the reference (set / insertion-ordered dict) because set iteration
    order decides the order of dormant tracks in the memory (track_head.py:84)."""

    def __init__(self, max_dormant):
        self.max_dormant = max_dormant
        self.reset()

    def reset(self):
        self.active = set()
        self.dormant = {}
        self.cache = {}
        self.next_id = 0
        self.frame = 0

    def start(self):
        i = self.next_id
        self.next_id += 1
        self.active.add(i)
        return i

    def suspend(self, i):
        if i not in self.active:
            raise ValueError
        self.active.remove(i)
        self.dormant[i] = self.frame - 1

    def resume(self, i):
        if i not in self.dormant or i in self.active:
            raise ValueError
        self.active.add(i)
        self.dormant.pop(i)

    def expire(self):
        for i, last in list(self.dormant.items()):
            if self.frame - last >= self.max_dormant:
                self.dormant.pop(i)
                self.cache.pop(i, None)

    def dormant_ids(self):
        return set(self.dormant.keys())


def solver_forward(cfg, pool, det):
    """TrackSolver.forward (track_solver.py:36-108)."""
    T = cfg.MODEL.TRACK_HEAD
    if det["boxes"].shape[0] == 0:
        return det
    all_ids = det["ids"]
    scores = det["scores"].clone()
    active_mask = torch.tensor([int(i) in pool.active for i in all_ids], dtype=torch.bool)
    scores[active_mask] += 1.0                                            # :69
    keep = prims.nms_legacy(det["boxes"], scores, 0.5)                    # :22
    boxes, ids, sc, labels = det["boxes"][keep], all_ids[keep].clone(), scores[keep], det["labels"][keep]
    sc[sc >= 2.0] = sc[sc >= 2.0] - 2.0                                   # :31-32
    sc[sc >= 1.0] = sc[sc >= 1.0] - 1.0
    start_idx = ((ids < 0) & (sc >= T.START_TRACK_THRESH)).nonzero()      # :78
    inactive = (ids >= 0) & (sc < T.TRACK_THRESH)                         # :81
    nms_track_ids = set(ids[ids >= 0].tolist())
    all_track_ids = set(all_ids[all_ids >= 0].tolist())
    inactive_ids = set(ids[inactive].tolist()) | (all_track_ids - nms_track_ids)
    dormant_now = pool.dormant_ids()
    dormant_mask = torch.tensor([int(i) in dormant_now for i in ids], dtype=torch.bool)
    for i in ids[dormant_mask & (sc >= T.RESUME_TRACK_THRESH)].tolist():  # :89-92
        pool.resume(i)
    for j in start_idx:                                                   # :94-95
        ids[j] = pool.start()
    for i in inactive_ids:                                                # :97-100
        if i in pool.active:
            pool.suspend(i)
    ids[inactive] = -1                                                    # :103
    pool.expire()
    pool.frame += 1
    return dict(boxes=boxes, scores=sc, ids=ids, labels=labels)


def build_memory(P, cfg, pool, feats, det):
    """TrackHead.get_track_memory (track_head.py:54-110) + EMM.extract_cache (track_core.py:81-98)."""
    T = cfg.MODEL.TRACK_HEAD
    sel = torch.tensor([int(i) in pool.active for i in det["ids"]], dtype=torch.bool)
    boxes, ids, labels = det["boxes"][sel], det["ids"][sel], det["labels"][sel]
    if boxes.shape[0] == 0:
        feat = torch.zeros((0,))
        sr = boxes.clone()
    else:
        feat = pool_rois(feats, boxes, boxes, T.POOLER_SCALES, T.POOLER_RESOLUTION, T.POOLER_SAMPLING_RATIO)
        sr = search_region(boxes, T.PAD_PIXELS, T.SEARCH_REGION - 1.0, T.MINIMUM_SREACH_REGION)
    mem = dict(feat=feat, sr=sr, boxes=boxes, ids=ids, labels=labels)
    if pool.cache:                                                        # track_head.py:77-97
        dorm = [pool.cache[i] for i in pool.dormant_ids() if i in pool.cache]
        if dorm:
            feats_l = ([mem["feat"]] if mem["feat"].numel() > 0 else []) + [d["feat"][None] for d in dorm]
            mem = dict(feat=torch.cat(feats_l),
                       sr=torch.cat([mem["sr"]] + [d["sr"] for d in dorm]),
                       boxes=torch.cat([mem["boxes"]] + [d["boxes"] for d in dorm]),
                       ids=torch.cat([mem["ids"]] + [d["ids"] for d in dorm]),
                       labels=torch.cat([mem["labels"]] + [d["labels"] for d in dorm]))
    for k in range(mem["boxes"].shape[0]):                                # update_cache :180-197
        pool.cache[int(mem["ids"][k])] = dict(feat=mem["feat"][k], sr=mem["sr"][k:k + 1],
                                              boxes=mem["boxes"][k:k + 1], ids=mem["ids"][k:k + 1],
                                              labels=mem["labels"][k:k + 1])
    return mem


class OracleSiamMOT(object):
    """SiamMOT.forward at inference (rcnn.py:41-68) + CombinedROIHeads.forward (roi_heads.py:21-51)."""

    def __init__(self, cfg, params):
        self.cfg = cfg
        self.P = {k: v.detach().to(torch.float32).cpu() for k, v in params.items()}
        self.pool = PoolState(cfg.MODEL.TRACK_HEAD.MAX_DORMANT_FRAMES)
        self.memory = None
        self.trace = {}

    def reset(self):
        self.memory = None
        self.pool.reset()

    def features(self, image):
        if image.dim() == 3:
            image = image[None]
        body = self.cfg.MODEL.BACKBONE.CONV_BODY
        if body in ("R-50-FPN", "R-101-FPN"):
            return fpn_forward(self.P, resnet50_forward(self.P, image.to(torch.float32),
                                                        stride_in_1x1=self.cfg.MODEL.RESNETS.STRIDE_IN_1X1,
                                                        blocks=(3, 4, 23, 3) if body == "R-101-FPN" else R50_BLOCKS))
        if body == "DLA-34-FPN":
            return fpn_forward(self.P, dla34_forward(self.P, image.to(torch.float32)))
        return fpn_forward(self.P, dla_forward(self.P, image.to(torch.float32), body))

    @torch.no_grad()
    def forward(self, image, given_detection=None):
        P, cfg = self.P, self.cfg
        if image.dim() == 3:
            image = image[None]
        img_h, img_w = image.shape[-2:]
        feats = self.features(image)
        props, _ = rpn_forward(P, cfg, feats, img_w, img_h)
        tr = self.trace = dict(proposals=props)
        if given_detection is None:
            det = box_head_forward(P, cfg, feats, props, img_w, img_h)
        elif given_detection["boxes"].shape[0] > 0:
            det = box_head_forward(P, cfg, feats, given_detection["boxes"], img_w, img_h,
                                   given_detection.get("ids"), given_detection.get("labels"))
        else:
            det = given_detection
        tr["detections"] = det
        tracks = None
        if self.memory is None:
            self.pool.reset()                                             # track_head.py:39-40
        elif self.memory["feat"].numel() > 0:
            tracks = emm_forward(P, cfg, feats, self.memory, img_w, img_h)
            tr["tracks"] = tracks
        if tracks is not None:
            if tracks["boxes"].shape[0] == 0:
                # roi_heads.py:64-65 returns a bare BoxList here and :44 then fails on list + BoxList
                raise TypeError("reference fails when every propagated track is clipped away")
            ref = box_head_forward(P, cfg, feats, tracks["boxes"], img_w, img_h, tracks["ids"], tracks["labels"])
            if cfg.MODEL.TRACK_HEAD.TRACKTOR:
                sc = ref["scores"]
            else:
                sc = (ref["scores"] + (tracks["scores"] + 1.0)) / 2.0      # roi_heads.py:67,76
            tracks = dict(boxes=ref["boxes"], scores=sc, ids=ref["ids"], labels=ref["labels"])
            tr["refined"] = tracks
            det = {k: torch.cat([det[k], tracks[k]]) for k in ("boxes", "scores", "ids", "labels")}
        det = solver_forward(cfg, self.pool, det)
        self.memory = build_memory(P, cfg, self.pool, feats, det)
        return det
