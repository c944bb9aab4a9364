"""Decision margins of a scene (test infrastructure for the fp16 end-to-end parity tests).

Track ids are assigned by thresholding and ordering float scores (track_solver.py:78,81,90; inference.py:163; NMS 0.5), so
"bit-exact ids" from an fp16 engine is only a meaningful demand on a scene whose decisions are not within fp16 noise of
flipping.  This module runs the CPU oracle over a clip and reports, per frame, how close every id-deciding comparison came to
its threshold (in the oracle's own fp32 numbers):

  det_thresh   |p - SCORE_THRESH| over every (proposal, foreground class)                     inference.py:163
  det_nms      |IoU - NMS| over the pairs the per-class NMS compares, and the score gap of every pair one of which
               suppresses the other (a swap would keep the other box)                          inference.py:170-174
  solver_nms   the same for the solver's NMS over detections + refined tracks                  track_solver.py:21-34
  solver_thr   |score - START/TRACK/RESUME_TRACK_THRESH| of the kept candidates                track_solver.py:78,81,90
  emm_argmax   relative gap between the best and the best non-adjacent position of each track's score map (a flip moves the
               box by more than interpolation noise)                                            track_core.py:120

The RPN's rank cuts (top-1000 per level, top-300 overall) are not listed: among 225 060 anchors some always sit within noise
of a cut; what matters is whether a proposal that enters or leaves changes a detection, which the det_* margins above see.
"""
import numpy as np
import torch

from oracle import prims
from oracle import siammot_oracle as so


def _iou_legacy(a, b):
    """IoU with +1 widths (upstream nms.cu) of box a (4,) against boxes b (m,4), fp32."""
    a = a.astype(np.float32)
    b = b.astype(np.float32)
    one = np.float32(1)
    aa = (a[2] - a[0] + one) * (a[3] - a[1] + one)
    ab = (b[:, 2] - b[:, 0] + one) * (b[:, 3] - b[:, 1] + one)
    w = np.maximum(np.minimum(a[2], b[:, 2]) - np.maximum(a[0], b[:, 0]) + one, np.float32(0))
    h = np.maximum(np.minimum(a[3], b[:, 3]) - np.maximum(a[1], b[:, 1]) + one, np.float32(0))
    inter = w * h
    return inter / (aa + ab - inter)


def nms_margins(boxes, scores, thresh):
    """(min |IoU - thresh| over compared pairs, min score gap over pairs with IoU > thresh) of one legacy NMS call."""
    n = boxes.shape[0]
    if n < 2:
        return float("inf"), float("inf")
    order = torch.sort(scores, descending=True, stable=True)[1]
    b = boxes[order].to(torch.float32).numpy()
    s = scores[order].to(torch.float32).numpy()
    dead = np.zeros(n, dtype=bool)
    m_iou, m_gap = float("inf"), float("inf")
    for i in range(n - 1):
        iou = _iou_legacy(b[i], b[i + 1:])
        over = iou > np.float32(thresh)
        if over.any():   # ordering matters for every overlapping pair, suppressed or not (a swap changes who survives)
            m_gap = min(m_gap, float(np.min(s[i] - s[i + 1:][over])))
        if dead[i]:
            continue
        alive = ~dead[i + 1:]
        if alive.any():
            m_iou = min(m_iou, float(np.min(np.abs(iou[alive] - np.float32(thresh)))))
        dead[i + 1:] |= over
    return m_iou, m_gap


def emm_argmax_gap(orc, cfg, feats, memory):
    """Per track: (best score - best score at a position farther than 2 px of the 256x256 map) / best score."""
    T = cfg.MODEL.TRACK_HEAD
    res = T.POOLER_RESOLUTION
    sres = int(res * T.SEARCH_REGION)
    import torch.nn.functional as F
    padded = so.pad_features(feats, T.PAD_PIXELS)
    srf = so.pool_rois(padded, memory["boxes"], memory["boxes"], T.POOLER_SCALES, sres, T.POOLER_SAMPLING_RATIO, rois=memory["sr"])
    cls, ctr, reg = so.emm_predictor(orc.P, so.xcorr_depthwise(srf, memory["feat"]))
    up = 16
    cls, ctr, reg = (F.interpolate(x, scale_factor=up, mode="bicubic") for x in (cls, ctr, reg))
    n = cls.shape[0]
    p1 = F.softmax(cls, dim=1)[:, 1].reshape(n, -1)
    conf = p1 * torch.sigmoid(ctr).reshape(n, -1) if T.EMM.USE_CENTERNESS else p1
    tlbr = reg.reshape(n, 4, -1)
    tb = memory["boxes"]
    sw = (tlbr[:, 2] + tlbr[:, 0]) / (tb[:, 2] - tb[:, 0])[:, None]
    sh = (tlbr[:, 3] + tlbr[:, 1]) / (tb[:, 3] - tb[:, 1])[:, None]
    pen = torch.exp((-torch.max(sw, 1 / sw) * torch.max(sh, 1 / sh) + 1) * 0.1)
    side = cls.shape[-1]
    hann = torch.hann_window(side, dtype=torch.float)
    sig = T.EMM.COSINE_WINDOW_WEIGHT
    score = ((conf * pen) * (1 - sig) + sig * torch.outer(hann, hann).reshape(-1)[None]).reshape(n, side, side)
    gaps = []
    for i in range(n):
        s = score[i]
        k = int(torch.argmax(s))
        iy, ix = k // side, k % side
        m = s.clone()
        m[max(iy - 2, 0):iy + 3, max(ix - 2, 0):ix + 3] = -1
        gaps.append(float((s[iy, ix] - m.max()) / s[iy, ix]))
    return gaps


class MarginOracle(object):
    """OracleSiamMOT.forward plus the margins of the frame's id-deciding comparisons (recomputed from the oracle's own trace)."""

    def __init__(self, cfg, sd):
        self.cfg = cfg
        self.orc = so.OracleSiamMOT(cfg, sd)

    def inject(self, frame, boxes):
        orc, cfg = self.orc, self.cfg
        orc.reset()
        feats = orc.features(frame)
        ids = torch.tensor([orc.pool.start() for _ in range(len(boxes))])
        det = dict(boxes=boxes, scores=torch.full((len(boxes),), 0.9), ids=ids, labels=torch.ones(len(boxes), dtype=torch.int64))
        orc.memory = so.build_memory(orc.P, cfg, orc.pool, feats, det)
        orc.pool.frame += 1

    def step(self, frame, with_emm_gap=False):
        orc, cfg = self.orc, self.cfg
        mem_in = orc.memory
        active_before = set(orc.pool.active)
        dormant_before = set(orc.pool.dormant_ids())
        out = orc.forward(frame)
        tr = orc.trace
        H, T = cfg.MODEL.ROI_HEADS, cfg.MODEL.TRACK_HEAD
        m = {}
        img_h, img_w = frame.shape[-2:]
        feats = orc.features(frame)
        # ---- detection threshold + per-class NMS (inference.py:163-174), from the proposals' class probabilities
        logits, deltas = so.box_head_features(orc.P, cfg, feats, tr["proposals"])
        prob = torch.softmax(logits, -1)
        m["det_thresh"] = float((prob[:, 1:] - H.SCORE_THRESH).abs().min())
        if prob.shape[1] == 2:   # calibration aid (tools/parity_probe.py --calibrate): the largest foreground-vs-background logit gaps
            m["top_logit_diff"] = [round(float(v), 4) for v in torch.sort(logits[:, 1] - logits[:, 0], descending=True)[0][:40]]
        dec = prims.box_decode(deltas, tr["proposals"], H.BBOX_REG_WEIGHTS).reshape(-1, 4)
        if not cfg.INPUT.AMODAL:
            dec = prims.clip_boxes(dec, img_w, img_h)
        dec = dec.reshape(prob.shape[0], -1)
        mi, mg = float("inf"), float("inf")
        for j in range(1, prob.shape[1]):
            inds = (prob[:, j] > H.SCORE_THRESH).nonzero().squeeze(1)
            a, b = nms_margins(dec[inds, 4 * j:4 * j + 4], prob[inds, j], H.NMS)
            mi, mg = min(mi, a), min(mg, b)
        m["det_nms_iou"], m["det_nms_gap"] = mi, mg
        m["n_det"] = int(tr["detections"]["boxes"].shape[0])
        # ---- solver (track_solver.py:36-108): candidates = detections + refined tracks
        cand = tr["detections"]
        if "refined" in tr:
            cand = {k: torch.cat([cand[k], tr["refined"][k]]) for k in ("boxes", "scores", "ids", "labels")}
        sc = cand["scores"].clone()
        act = torch.tensor([int(i) in active_before for i in cand["ids"]], dtype=torch.bool)
        sc[act] += 1.0
        m["solver_nms_iou"], m["solver_nms_gap"] = nms_margins(cand["boxes"], sc, 0.5)
        keep = prims.nms_legacy(cand["boxes"], sc, 0.5)
        ks, kid = sc[keep].clone(), cand["ids"][keep]
        ks[ks >= 2.0] -= 2.0
        ks[ks >= 1.0] -= 1.0
        thr = [float("inf")]
        new = kid < 0
        if new.any():
            thr.append(float((ks[new] - T.START_TRACK_THRESH).abs().min()))
        if (~new).any():
            thr.append(float((ks[~new] - T.TRACK_THRESH).abs().min()))
        dm = torch.tensor([int(i) in dormant_before for i in kid], dtype=torch.bool)
        if dm.any():
            thr.append(float((ks[dm] - T.RESUME_TRACK_THRESH).abs().min()))
        m["solver_thr"] = min(thr)
        m["n_cand"], m["n_kept"], m["n_tracked"] = int(sc.numel()), int(keep.numel()), int((out["ids"] >= 0).sum())
        if with_emm_gap and mem_in is not None and mem_in["feat"].numel() > 0:
            g = emm_argmax_gap(orc, cfg, feats, mem_in)
            m["emm_argmax_gap"] = min(g) if g else float("inf")
            m["emm_gap_by_id"] = {int(i): float(v) for i, v in zip(mem_in["ids"].tolist(), g)}
        return out, m


ID_MARGINS = ("det_thresh", "det_nms_iou", "det_nms_gap", "solver_nms_iou", "solver_nms_gap", "solver_thr")


def min_margin(per_frame):
    """(value, name, frame) of the smallest id-deciding margin of a clip."""
    best = (float("inf"), None, None)
    for t, m in enumerate(per_frame):
        for k in ID_MARGINS:
            if m[k] < best[0]:
                best = (m[k], k, t)
    return best
