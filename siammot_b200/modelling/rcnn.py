"""``build_siammot(cfg)`` -> drop-in replacement of the reference's SiamMOT module
(/root/reference/siammot/modelling/rcnn.py:15-73) for inference.

Same surface: ``forward(images, targets=None, given_detection=None) -> [BoxList]``,
``reset_siammot_status()``, ``flush_memory(cache)``, ``track_memory``, sub-modules ``backbone``
(``body`` / ``fpn``), ``rpn``, ``roi_heads`` (``box`` / ``track`` / ``solver``), the reference's
state-dict keys, and the ``SIAMESE_TRACKER`` plugin registry.  All arithmetic runs in libsmot.so
through :class:`siammot_b200.engine.Engine`; this file is host control flow only
(CombinedROIHeads.forward roi_heads.py:21-51, TrackSolver.forward track_solver.py:36-108,
TrackHead.get_track_memory track_head.py:54-110), restructured so that a frame costs one
device->host copy.
"""
import torch
from torch import nn

from .. import ops
from ..engine import Engine, cell_anchors
from ..structures import BoxList
from ..synthetic import dla34_layout, make_state_dict
from . import registry
from .track_utils import build_track_utils


# ----------------------------------------------------------------------------------------------
# parameter containers (reference module tree / state-dict keys)
# ----------------------------------------------------------------------------------------------
class _Holder(nn.Module):
    """Plain container: parameters and buffers only, no forward."""


def _attach(root, key, tensor, as_buffer):
    parts = key.split(".")
    mod = root
    for p in parts[:-1]:
        if p not in mod._modules:
            mod.add_module(p, _Holder())
        mod = mod._modules[p]
    if as_buffer:
        mod.register_buffer(parts[-1], tensor)
    else:
        mod.register_parameter(parts[-1], nn.Parameter(tensor, requires_grad=False))


def _is_frozen_bn_key(key, bn_names):
    return key.rsplit(".", 1)[0] in bn_names


class Memory(object):
    """Track memory for the next frame: N = active tracks (first) + dormant tracks."""

    def __init__(self, feat, sr, boxes, ids, labels, n_active, device):
        self.feat = feat                  # device (N,T,T,C) activation dtype
        self.sr = sr                      # cpu (N,4) fp32, padded frame
        self.boxes = boxes                # cpu (N,4) fp32
        self.ids = ids                    # list[int]
        self.labels = labels              # cpu int64 (N,)
        self.n_active = n_active
        self.n = boxes.shape[0]
        if self.n:
            self.sr_dev = sr.to(device, non_blocking=True)
            self.boxes_dev = boxes.to(device, non_blocking=True)
            self.labels_dev = labels.to(torch.int32).to(device, non_blocking=True)
            act = torch.zeros(self.n, dtype=torch.float32)
            act[:n_active] = 1.0
            self.active_dev = act.to(device, non_blocking=True)


@registry.SIAMESE_TRACKER.register("EMM")
class EMM(nn.Module):
    """The Explicit Motion Model tracker (track_core.py:14-98) on the engine."""

    def __init__(self, cfg, track_utils):
        super().__init__()
        self.cfg = cfg
        self.track_utils = track_utils
        self.predictor = _Holder()
        self.engine = None  # set by SiamMOT

    def track_device(self, plan, mem):
        return self.engine.emm_track(plan, mem.feat, mem.sr_dev, mem.boxes_dev)

    def forward(self, features, boxes, sr, targets=None, template_features=None):
        """Reference contract: returns ({}, [BoxList], {}) with clipped, non-empty track boxes.
        ``features`` is the engine plan of the current frame."""
        dev = self.engine.device
        b, s = boxes[0], sr[0]
        tb, conf, valid = self.engine.emm_track(features, template_features, s.bbox.to(dev).contiguous(),
                                                b.bbox.to(dev).contiguous())
        keep = valid.bool()
        out = BoxList(tb[keep], b.size, mode="xyxy")
        out.add_field("ids", b.get_field("ids").to(dev)[keep])
        out.add_field("labels", b.get_field("labels").to(dev)[keep])
        out.add_field("scores", conf[keep])
        return {}, [out], {}

    def extract_cache(self, features, detection):
        dev = self.engine.device
        x = self.engine.templates(features, detection.bbox.to(dev).contiguous())
        sr = self.track_utils.extend_bbox(self.track_utils.update_boxes_in_pad_images([detection.to("cpu")]))
        return x, sr, [detection]


class TrackHead(nn.Module):
    def __init__(self, tracker, sampler, track_utils, track_pool):
        super().__init__()
        self.tracker = tracker
        self.sampler = sampler
        self.track_utils = track_utils
        self.track_pool = track_pool

    def reset_track_pool(self):
        self.track_pool.reset()


class TrackSolver(nn.Module):
    def __init__(self, track_pool, track_thresh=0.3, start_track_thresh=0.5, resume_track_thresh=0.4):
        super().__init__()
        self.track_pool = track_pool
        self.track_thresh = track_thresh
        self.start_thresh = start_track_thresh
        self.resume_track_thresh = resume_track_thresh

    def resolve(self, boxes, scores_adj, all_ids, labels, all_track_ids):
        """Host half of TrackSolver.forward (track_solver.py:71-106) given the NMS survivors, in NMS order.
        scores_adj still carries the +1 (dormant / refined) and +2 (active) offsets."""
        pool = self.track_pool
        _scores = scores_adj.clone()
        _scores[_scores >= 2.] = _scores[_scores >= 2.] - 2.
        _scores[_scores >= 1.] = _scores[_scores >= 1.] - 1.
        _ids = all_ids.clone()
        start_idxs = ((_ids < 0) & (_scores >= self.start_thresh)).nonzero()
        inactive_idxs = ((_ids >= 0) & (_scores < self.track_thresh))
        nms_track_ids = set(_ids[_ids >= 0].tolist())
        nms_removed_ids = all_track_ids - nms_track_ids
        inactive_ids = set(_ids[inactive_idxs].tolist()) | nms_removed_ids
        dormant_ids = pool.get_dormant_ids()
        dormant_mask = torch.tensor([int(x) in dormant_ids for x in _ids], dtype=torch.bool)
        resume_ids = _ids[dormant_mask & (_scores >= self.resume_track_thresh)]
        for _id in resume_ids.tolist():
            pool.resume_track(_id)
        for _idx in start_idxs:
            _ids[_idx] = pool.start_track()
        active_ids = pool.get_active_ids()
        for _id in inactive_ids:
            if _id in active_ids:
                pool.suspend_track(_id)
        _ids[inactive_idxs] = -1
        pool.expire_tracks()
        pool.increment_frame()
        return boxes, _scores, _ids, labels


class CombinedROIHeads(nn.ModuleDict):
    def __init__(self, cfg, heads):
        super().__init__(heads)
        self.cfg = cfg
        self.engine = None

    def reset_roi_status(self):
        if self.cfg.MODEL.TRACK_ON:
            self.track.reset_track_pool()

    # -- detections from externally provided boxes (roi_heads.py:26-34)
    def _given_detections(self, P, given):
        eng, dev, cfg = self.engine, self.engine.device, self.cfg
        rois = given.convert("xyxy").bbox.to(dev, torch.float32).contiguous()
        n = rois.shape[0]
        dec_b, dec_s = eng.box_head_eager(P, rois)
        ncls = eng.ncls
        cap = n * (ncls - 1)
        det_boxes = torch.zeros((cap, 4), dtype=torch.float32, device=dev)
        det_scores = torch.full((cap,), -1.0, dtype=torch.float32, device=dev)
        det_labels = torch.zeros((cap,), dtype=torch.int32, device=dev)
        det_count = torch.zeros((1,), dtype=torch.int32, device=dev)
        H = cfg.MODEL.ROI_HEADS
        for j in range(1, ncls):
            ops.sort_nms(dec_b[:, j], dec_s[:, j], det_count, n_max=n, min_score=H.SCORE_THRESH, thresh=H.NMS,
                         max_keep=n, tag=j, out_boxes=det_boxes, out_scores=det_scores, out_tag=det_labels,
                         workspace=eng.nms_workspace(n), box_stride=4 * ncls, score_stride=ncls)
        return det_boxes, det_scores, det_labels, det_count

    def run_frame(self, P, mem, given_detection=None):
        """One frame after the static stage.  Returns (BoxList on device, Memory for the next frame)."""
        eng, dev, cfg = self.engine, self.engine.device, self.cfg
        pool = self.track.track_pool
        img_size = (P.W, P.H)
        if given_detection is None:
            det_boxes, det_scores, det_labels, det_count = P.det_boxes, P.det_scores, P.det_labels, P.det_count
        elif len(given_detection[0]) > 0:
            det_boxes, det_scores, det_labels, det_count = self._given_detections(P, given_detection[0])
        else:
            det_boxes = torch.zeros((0, 4), dtype=torch.float32, device=dev)
            det_scores = torch.zeros((0,), dtype=torch.float32, device=dev)
            det_labels = torch.zeros((0,), dtype=torch.int32, device=dev)
            det_count = torch.zeros((1,), dtype=torch.int32, device=dev)
        ncap = det_boxes.shape[0]

        if not cfg.MODEL.TRACK_ON:
            raise NotImplementedError("MODEL.TRACK_ON False")
        if mem is None:
            pool.reset()                                                  # track_head.py:39-40
        n_trk = mem.n if (mem is not None and mem.feat.numel() > 0) else 0
        if n_trk:
            tb, conf, valid = self.track.tracker.track_device(P, mem)     # EMM.forward
            dec_b, dec_s = eng.box_head_eager(P, tb, mem.labels_dev)      # _refine_tracks roi_heads.py:60-84
            rows = torch.arange(n_trk, device=dev)
            lab = mem.labels_dev.long()
            ref_boxes = dec_b[rows, lab]
            det_part = dec_s[rows, lab]                                   # p + 1   (inference.py:103)
            if cfg.MODEL.TRACK_HEAD.TRACKTOR:
                trk_scores = det_part
            else:
                trk_scores = (det_part + (conf + 1.)) / 2.                # roi_heads.py:67,76
            trk_scores = trk_scores + mem.active_dev                      # track_solver.py:69
            trk_scores = torch.where(valid > 0, trk_scores, torch.full_like(trk_scores, -1.0))
            cat_boxes = torch.cat([det_boxes, ref_boxes])
            cat_scores = torch.cat([det_scores, trk_scores])
        else:
            cat_boxes, cat_scores = det_boxes, det_scores
        total = cat_boxes.shape[0]
        keep_idx = torch.zeros((max(total, 1),), dtype=torch.int32, device=dev)
        keep_cnt = torch.zeros((1,), dtype=torch.int32, device=dev)
        if total:
            ops.sort_nms(cat_boxes, cat_scores, keep_cnt, n_max=total, min_score=-0.5, thresh=0.5, max_keep=total,
                         out_index=keep_idx, workspace=eng.nms_workspace(total))
        # ---- the frame's single device->host copy
        pack = torch.cat([keep_cnt.view(torch.float32), det_count.view(torch.float32), keep_idx.view(torch.float32),
                          cat_boxes.reshape(-1), cat_scores, det_labels.view(torch.float32)])
        host = pack.cpu()
        k = int(host[0:1].view(torch.int32))
        o = 2
        h_keep = host[o:o + max(total, 1)].view(torch.int32)[:k].long(); o += max(total, 1)
        h_boxes = host[o:o + 4 * total].view(total, 4); o += 4 * total
        h_scores = host[o:o + total]; o += total
        h_dlabels = host[o:o + ncap].view(torch.int32).long()

        # ids / labels of every candidate row (detections first, then memory rows)
        all_ids = torch.full((total,), -1, dtype=torch.int64)
        all_labels = torch.zeros((total,), dtype=torch.int64)
        all_labels[:ncap] = h_dlabels
        if n_trk:
            all_ids[ncap:] = torch.tensor(mem.ids, dtype=torch.int64)
            all_labels[ncap:] = mem.labels
            trk_valid = h_scores[ncap:] > -0.5
            if not bool(trk_valid.any()):
                # roi_heads.py:64-65 returns a bare BoxList here and :44 then evaluates list + BoxList
                raise TypeError("can only concatenate list (not \"BoxList\") to list")
            all_track_ids = set(all_ids[ncap:][trk_valid].tolist())
        else:
            all_track_ids = set()

        if k == 0 and not (h_scores > -0.5).any():                        # track_solver.py:51-52
            boxes, scores = torch.zeros((0, 4)), torch.zeros((0,))
            ids, labels = torch.zeros((0,), dtype=torch.int64), torch.zeros((0,), dtype=torch.int64)
        else:
            boxes, scores, ids, labels = self.solver.resolve(h_boxes[h_keep], h_scores[h_keep], all_ids[h_keep],
                                                             all_labels[h_keep], all_track_ids)
        new_mem = self._build_memory(P, boxes, ids, labels)
        result = BoxList(boxes.to(dev, non_blocking=True), img_size, mode="xyxy")
        result.add_field("scores", scores.to(dev, non_blocking=True))
        result.add_field("ids", ids.to(dev, non_blocking=True))
        result.add_field("labels", labels.to(dev, non_blocking=True))
        return result, new_mem

    def _build_memory(self, P, boxes, ids, labels):
        """TrackHead.get_track_memory (track_head.py:54-110) + EMM.extract_cache (track_core.py:81-98)."""
        eng, dev = self.engine, self.engine.device
        pool = self.track.track_pool
        tu = self.track.track_utils
        active_ids = pool.get_active_ids()
        sel = torch.tensor([int(i) in active_ids for i in ids.tolist()], dtype=torch.bool)
        a_boxes, a_ids, a_labels = boxes[sel], ids[sel].tolist(), labels[sel]
        n_act = a_boxes.shape[0]
        cache = pool.get_cache()
        dormant = [cache[i] for i in pool.get_dormant_ids() if i in cache] if cache else []
        m_boxes = torch.cat([a_boxes] + [d["box"] for d in dormant]) if dormant else a_boxes
        a_sr = tu.search_region(a_boxes) if n_act else torch.zeros((0, 4))
        m_sr = torch.cat([a_sr] + [d["sr"] for d in dormant]) if dormant else a_sr
        m_ids = a_ids + [d["id"] for d in dormant]
        m_labels = torch.cat([a_labels] + [d["label"] for d in dormant]) if dormant else a_labels
        n = m_boxes.shape[0]
        if n == 0:
            feat = torch.zeros((0,), device=dev)
        else:
            parts = []
            if n_act:
                parts.append(eng.templates(P, a_boxes.to(dev).contiguous()))
            parts += [d["feat"][None] for d in dormant]
            feat = torch.cat(parts) if len(parts) > 1 else parts[0]
        mem = Memory(feat, m_sr, m_boxes, m_ids, m_labels, n_act, dev)
        pool.update_cache({m_ids[r]: dict(feat=feat[r], sr=m_sr[r:r + 1], box=m_boxes[r:r + 1], id=m_ids[r],
                                          label=m_labels[r:r + 1]) for r in range(n)})
        return mem


class SiamMOT(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        # ---- parameter tree with the reference's names; values: seeded synthetic init (no checkpoints offline)
        bn_names = set("backbone.body." + n for kind, n, _ in dla34_layout() if kind == "bn")
        self.backbone = _Holder()
        self.backbone.out_channels = cfg.MODEL.DLA.BACKBONE_OUT_CHANNELS
        self.rpn = _Holder()
        track_utils, track_pool = build_track_utils(cfg)
        tracker = registry.SIAMESE_TRACKER[cfg.MODEL.TRACK_HEAD.MODEL](cfg, track_utils)
        sampler = registry.TRACKER_SAMPLER.get(cfg.MODEL.TRACK_HEAD.MODEL, lambda c, t: None)(cfg, track_utils)
        T = cfg.MODEL.TRACK_HEAD
        heads = [("box", _Holder()), ("track", TrackHead(tracker, sampler, track_utils, track_pool)),
                 ("solver", TrackSolver(track_pool, T.TRACK_THRESH, T.START_TRACK_THRESH, T.RESUME_TRACK_THRESH))]
        self.roi_heads = CombinedROIHeads(cfg, heads)
        for key, val in make_state_dict(cfg, seed=0).items():
            _attach(self, key, val, as_buffer=_is_frozen_bn_key(key, bn_names))
        R = cfg.MODEL.RPN
        for i, (st, sz) in enumerate(zip(R.ANCHOR_STRIDE, R.ANCHOR_SIZES)):
            _attach(self, "rpn.anchor_generator.cell_anchors.%d" % i, cell_anchors(st, (sz,), R.ASPECT_RATIOS), True)
        self.track_memory = None
        self._mem = None
        self._engine = None
        self._engine_stale = True
        self.eval()

    # ---- engine management
    def engine(self):
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("siammot_b200 runs on CUDA only (model is on %s): call .to('cuda'); there is no CPU path" % dev)
        if self._engine is None or self._engine.device != dev:
            self._engine = Engine(self.cfg, device=dev)
            self._engine_stale = True
        if self._engine_stale:
            self._engine.load_state_dict(self.state_dict())
            self._engine_stale = False
        self.roi_heads.engine = self._engine
        self.roi_heads.track.tracker.engine = self._engine
        return self._engine

    def load_state_dict(self, state_dict, strict=True):
        sd = {(k[7:] if k.startswith("module.") else k): v for k, v in state_dict.items()}
        r = super().load_state_dict(sd, strict=strict)
        self._engine_stale = True
        return r

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self._engine_stale = True
        return r

    # ---- reference API
    def flush_memory(self, cache=None):
        self.track_memory = cache
        self._mem = cache if (cache is None or isinstance(cache, Memory)) else self._memory_from_tuple(cache)

    def reset_siammot_status(self):
        self.flush_memory()
        self.roi_heads.reset_roi_status()

    def _memory_from_tuple(self, cache):
        feats, sr, boxes = cache
        b = boxes[0].to("cpu")
        n_act = sum(1 for i in b.get_field("ids").tolist() if i in self.roi_heads.track.track_pool.get_active_ids())
        return Memory(feats, sr[0].bbox.to("cpu").float(), b.bbox.float(), b.get_field("ids").tolist(),
                      b.get_field("labels").to(torch.int64), n_act, next(self.parameters()).device)

    @torch.no_grad()
    def forward(self, images, targets=None, given_detection=None):
        if self.training:
            raise NotImplementedError("siammot_b200 is an inference engine: call .eval() (training is out of scope)")
        eng = self.engine()
        if hasattr(images, "tensors"):
            images = images.tensors
        P = eng.run_static(images)
        result, mem = self.roi_heads.run_frame(P, self._mem, given_detection)
        self._mem = mem
        self.track_memory = mem
        return [result]


def build_siammot(cfg):
    return SiamMOT(cfg)
