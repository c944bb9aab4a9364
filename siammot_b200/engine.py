"""Per-frame execution engine: weight preparation, static launch plan, dynamic tracker stage.

This is the B200-native replacement for everything beneath ``SiamMOT.forward``
(/root/reference/siammot/modelling/rcnn.py:41-68).  Design:

* activations live in NHWC device buffers allocated once per input resolution; DLA roots read their
  children through channel-slice views (no torch.cat, dla.py:183), FrozenBN / bias / residual / ReLU
  ride in the conv epilogues;
* the frame-independent stage (backbone -> FPN -> RPN head -> proposal selection -> box head ->
  per-class NMS) is a fixed list of C-ABI launches with device-side counts, captured once into a
  CUDA graph and replayed per frame;
* the track-dependent stage (search-region ROIAlign with virtual padding -> xcorr -> EMM towers ->
  fused decode -> box-head refinement -> solver NMS) is launched eagerly with N = tracks in memory;
* exactly one device->host copy per frame (the solver needs ids on the host, track_solver.py:62-106),
  against >= 10 hidden syncs in the reference.

All compute is libsmot.so (hand-written sm_100a CUDA); torch is used for memory, streams, graphs and
a few single-element glue ops on tiny tensors.  No CPU fallback exists.
"""
import ctypes as C
import math

import torch

from . import _lib, ops
from ._lib import check, lib


def _ohwi(w, dtype, device):
    return w.detach().to(torch.float32).permute(0, 2, 3, 1).contiguous().to(device=device, dtype=dtype)


def _f32(t, device):
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


def cell_anchors(stride, sizes, aspect_ratios):
    """Detectron-style cell anchors (upstream rpn/anchor_generator.py generate_anchors): float64
    maths, legacy rounding, cast to fp32.  Returned as an (A,4) CPU tensor; also the value of the
    reference's ``rpn.anchor_generator.cell_anchors.*`` buffers."""
    import numpy as np
    scales = np.array(sizes, dtype=np.float64) / stride
    ratios = np.array(aspect_ratios, dtype=np.float64)

    def whctr(a):
        w, h = a[2] - a[0] + 1, a[3] - a[1] + 1
        return w, h, a[0] + 0.5 * (w - 1), a[1] + 0.5 * (h - 1)

    def mk(ws, hs, xc, yc):
        ws, hs = ws[:, None], hs[:, None]
        return np.hstack((xc - 0.5 * (ws - 1), yc - 0.5 * (hs - 1), xc + 0.5 * (ws - 1), yc + 0.5 * (hs - 1)))

    w, h, xc, yc = whctr(np.array([0, 0, stride - 1, stride - 1], dtype=np.float64))
    ws = np.round(np.sqrt(w * h / ratios))
    hs = np.round(ws * ratios)
    rows = []
    for ra in mk(ws, hs, xc, yc):
        w, h, xc, yc = whctr(ra)
        rows.append(mk(w * scales, h * scales, xc, yc))
    return torch.from_numpy(np.vstack(rows)).float()


class _Plan(object):
    """Launch list for one input resolution.  Each step is (callable, args-before-stream)."""

    def __init__(self, engine, H, W):
        self.e = engine
        self.H, self.W = H, W
        self.steps = []
        self.keep = []  # ctypes objects / tensors that must outlive the plan
        self.graph = None
        self.dev = engine.device
        self.dtype = engine.dtype

    def new(self, H, W, Cc, dtype=None, B=1):
        t = torch.zeros((B, H, W, Cc), dtype=dtype or self.dtype, device=self.dev)
        self.keep.append(t)
        return t

    def conv(self, x, name, out, residual=None, stride=1, pad=0, relu=False):
        w, scale, bias = self.e.weights[name]
        d = ops.conv_desc(x, w, out, scale, bias, residual, stride, pad, relu)
        self.keep.append(d)
        self.steps.append((lib().smot_conv2d, (C.byref(d),), "conv:" + name))
        return out

    def call(self, fn, args, tag):
        self.steps.append((fn, args, tag))

    def run_eager(self):
        st = _lib.stream_ptr()
        for fn, args, tag in self.steps:
            check(fn(*args, st), tag)

    def run(self):
        if self.e.use_graph:
            if self.graph is None:
                self.run_eager()  # warm-up: sets function attributes, surfaces argument errors
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self.run_eager()
                self.graph = g
            self.graph.replay()
        else:
            self.run_eager()


class Engine(object):
    def __init__(self, cfg, device="cuda", dtype=None, use_graph=True):
        if not torch.cuda.is_available():
            raise RuntimeError("siammot_b200 needs a CUDA device (sm_100a); there is no CPU path")
        lib()  # fail loudly if the CUDA library is missing
        self.cfg = cfg
        self.device = torch.device(device)
        if dtype is None:
            dtype = {"float32": torch.float32, "float16": torch.float16}[cfg.DTYPE]
        self.dtype = dtype
        self.use_graph = use_graph
        self.weights = {}
        self.plans = {}
        self.C = cfg.MODEL.DLA.BACKBONE_OUT_CHANNELS
        self.ncls = cfg.MODEL.ROI_BOX_HEAD.NUM_CLASSES
        if cfg.MODEL.BACKBONE.CONV_BODY != "DLA-34-FPN":
            raise NotImplementedError("only the DLA-34-FPN body is implemented (SURVEY.md section 8)")
        if cfg.MODEL.CLS_AGNOSTIC_BBOX_REG:
            raise NotImplementedError("CLS_AGNOSTIC_BBOX_REG")
        R = cfg.MODEL.RPN
        self.cells = [cell_anchors(st, (sz,), R.ASPECT_RATIOS) for st, sz in zip(R.ANCHOR_STRIDE, R.ANCHOR_SIZES)]
        self.n_anchor = self.cells[0].shape[0]
        T = cfg.MODEL.TRACK_HEAD
        self.t_res = T.POOLER_RESOLUTION
        self.s_res = int(T.POOLER_RESOLUTION * T.SEARCH_REGION)
        self.o_res = self.s_res - self.t_res + 1
        self.up = 16
        self.hann = torch.hann_window(self.o_res * self.up, dtype=torch.float).to(self.device)
        self.pads = [int(T.PAD_PIXELS / ((2 ** i) * 4)) for i in range(len(T.POOLER_SCALES))]
        self._nms_ws = {}
        self.timers = None  # optional dict name -> list of (start_event, end_event), see timed()

    def timed(self, name):
        """Context manager: when self.timers is a dict, brackets the enclosed launches with CUDA events on
        the launching stream (bench.py's live per-kernel timing)."""
        eng = self

        class _T(object):
            def __enter__(self_inner):
                if eng.timers is not None:
                    self_inner.e0 = torch.cuda.Event(enable_timing=True)
                    self_inner.e1 = torch.cuda.Event(enable_timing=True)
                    self_inner.e0.record()

            def __exit__(self_inner, *exc):
                if eng.timers is not None:
                    self_inner.e1.record()
                    eng.timers.setdefault(name, []).append((self_inner.e0, self_inner.e1))
                return False
        return _T()

    # ------------------------------------------------------------------------------------------
    # weights
    # ------------------------------------------------------------------------------------------
    def load_state_dict(self, sd):
        """sd: flat dict with the reference key layout."""
        dev, dt, Wt = self.device, self.dtype, self.weights
        Wt.clear()

        def bn(prefix):
            scale = sd[prefix + ".weight"].float() * sd[prefix + ".running_var"].float().rsqrt()
            bias = sd[prefix + ".bias"].float() - sd[prefix + ".running_mean"].float() * scale
            return _f32(scale, dev), _f32(bias, dev)

        body = "backbone.body."
        pairs = []
        for k in sd:
            if k.startswith(body) and k.endswith(".weight") and sd[k].dim() == 4:
                conv = k[:-len(".weight")]
                leaf = conv.rsplit(".", 1)[1]
                parent = conv.rsplit(".", 1)[0]
                if leaf in ("conv1", "conv2"):
                    bnn = parent + ".bn" + leaf[-1]
                elif leaf == "conv":
                    bnn = parent + ".bn"
                else:  # Sequential: conv at index i, bn at i+1
                    bnn = parent + "." + str(int(leaf) + 1)
                pairs.append((conv, bnn))
        for conv, bnn in pairs:
            s, b = bn(bnn)
            Wt[conv[len("backbone."):]] = (_ohwi(sd[conv + ".weight"], dt, dev), s, b)
        for i in range(1, 5):
            for kind in ("fpn_inner", "fpn_layer"):
                k = "backbone.fpn.%s%d" % (kind, i)
                Wt["fpn.%s%d" % (kind, i)] = (_ohwi(sd[k + ".weight"], dt, dev), None, _f32(sd[k + ".bias"], dev))
        Wt["rpn.conv"] = (_ohwi(sd["rpn.head.conv.weight"], dt, dev), None, _f32(sd["rpn.head.conv.bias"], dev))
        wp = torch.cat([sd["rpn.head.cls_logits.weight"], sd["rpn.head.bbox_pred.weight"]], 0)
        bp = torch.cat([sd["rpn.head.cls_logits.bias"], sd["rpn.head.bbox_pred.bias"]], 0)
        Wt["rpn.pred"] = (_ohwi(wp, dt, dev), None, _f32(bp, dev))
        pre = "roi_heads.box."
        res = self.cfg.MODEL.ROI_BOX_HEAD.POOLER_RESOLUTION
        w6 = sd[pre + "feature_extractor.fc6.weight"].float()
        rep = w6.shape[0]
        # reference flattens (C, res, res); ROIAlign here emits (res, res, C): permute fc6's input axis once
        w6 = w6.view(rep, self.C, res, res).permute(0, 2, 3, 1).reshape(rep, 1, 1, res * res * self.C)
        Wt["box.fc6"] = (w6.contiguous().to(dev, dt), None, _f32(sd[pre + "feature_extractor.fc6.bias"], dev))
        w7 = sd[pre + "feature_extractor.fc7.weight"].float()
        Wt["box.fc7"] = (w7.reshape(w7.shape[0], 1, 1, w7.shape[1]).contiguous().to(dev, dt), None,
                         _f32(sd[pre + "feature_extractor.fc7.bias"], dev))
        wc = torch.cat([sd[pre + "predictor.cls_score.weight"], sd[pre + "predictor.bbox_pred.weight"]], 0).float()
        bc = torch.cat([sd[pre + "predictor.cls_score.bias"], sd[pre + "predictor.bbox_pred.bias"]], 0)
        Wt["box.pred"] = (wc.reshape(wc.shape[0], 1, 1, wc.shape[1]).contiguous().to(dev, dt), None, _f32(bc, dev))
        pre = "roi_heads.track.tracker.predictor."
        wt = torch.cat([sd[pre + "cls_tower.0.weight"], sd[pre + "reg_tower.0.weight"]], 0)
        Wt["emm.towers"] = (_ohwi(wt, dt, dev), None, None)
        self.gn_gamma = _f32(torch.cat([sd[pre + "cls_tower.1.weight"], sd[pre + "reg_tower.1.weight"]]), dev)
        self.gn_beta = _f32(torch.cat([sd[pre + "cls_tower.1.bias"], sd[pre + "reg_tower.1.bias"]]), dev)
        wcc = torch.cat([sd[pre + "cls.weight"], sd[pre + "center.weight"]], 0)
        bcc = torch.cat([sd[pre + "cls.bias"], sd[pre + "center.bias"]], 0)
        Wt["emm.clsctr"] = (_ohwi(wcc, dt, dev), None, _f32(bcc, dev))
        Wt["emm.reg"] = (_ohwi(sd[pre + "reg.weight"], dt, dev), None, _f32(sd[pre + "reg.bias"], dev))
        self.plans.clear()

    # ------------------------------------------------------------------------------------------
    # static plan
    # ------------------------------------------------------------------------------------------
    def _tree(self, P, name, x, levels, cin, cout, stride, level_root, out=None, rootbuf=None):
        """DlaTree (dla.py:192-238) as launches.  x / out are NHWC views.  Returns the output view."""
        _, H, W, _ = x.shape
        Ho, Wo = H // stride, W // stride
        if levels == 1:
            total = 2 * cout + (cin if level_root else 0)
            if rootbuf is None:
                rootbuf = P.new(Ho, Wo, total)
            else:
                assert not level_root
            x2v, x1v = rootbuf[..., 0:cout], rootbuf[..., cout:2 * cout]
            if stride > 1:
                bottom = rootbuf[..., 2 * cout:2 * cout + cin] if level_root else P.new(Ho, Wo, cin)
                P.call(lib().smot_maxpool2x2, self._pool_args(x, bottom), "maxpool:" + name)
            else:
                bottom = x
            if cin != cout:
                residual = P.new(Ho, Wo, cout)
                P.conv(bottom, "body." + name + ".project.0", residual)
            else:
                residual = bottom
            a = P.new(Ho, Wo, cout)
            P.conv(x, "body." + name + ".tree1.conv1", a, stride=stride, pad=1, relu=True)
            P.conv(a, "body." + name + ".tree1.conv2", x1v, residual=residual, pad=1, relu=True)
            b = P.new(Ho, Wo, cout)
            P.conv(x1v, "body." + name + ".tree2.conv1", b, pad=1, relu=True)
            P.conv(b, "body." + name + ".tree2.conv2", x2v, residual=x1v, pad=1, relu=True)
            if out is None:
                out = P.new(Ho, Wo, cout)
            P.conv(rootbuf, "body." + name + ".root.conv", out, relu=True)
            return out
        assert levels == 2, "DLA-34 only nests two tree levels"
        total = 2 * cout + (cin if level_root else 0) + cout
        rootbuf = P.new(Ho, Wo, total)
        off = 2 * cout
        if level_root:
            P.call(lib().smot_maxpool2x2, self._pool_args(x, rootbuf[..., off:off + cin]), "maxpool:" + name)
            off += cin
        # the outer project (dla.py:228 overwrites its result) is dead code: skipped, results identical
        t1 = rootbuf[..., off:off + cout]
        self._tree(P, name + ".tree1", x, 1, cin, cout, stride, False, out=t1)
        return self._tree(P, name + ".tree2", t1, 1, cout, cout, 1, False, out=out, rootbuf=rootbuf)

    def _pool_args(self, x, out):
        B, H, W, Cc, ld = ops._nhwc(x)
        return (ops._ptr(x), ops._ptr(out), B, H, W, Cc, ld, ops._nhwc(out)[4], _lib.dtype_code(x.dtype))

    def plan(self, H, W):
        key = (H, W)
        if key in self.plans:
            return self.plans[key]
        if not self.weights:
            raise RuntimeError("Engine.load_state_dict() must be called before the first frame")
        if H % 32 or W % 32:
            raise ValueError("DLA-34 needs an input divisible by 32 (got %dx%d); the reference resizes to such a size "
                             "(DATALOADER.SIZE_DIVISIBILITY 32) and fails in dla.py:54 otherwise" % (H, W))
        cfg, dev, dt = self.cfg, self.device, self.dtype
        P = _Plan(self, H, W)
        L = lib()
        dc = _lib.dtype_code(dt)
        # ---- input
        P.img_in = torch.zeros((3, H, W), dtype=torch.float32, device=dev)
        img = P.new(H, W, 4)
        P.call(L.smot_image_to_nhwc, (ops._ptr(P.img_in), ops._ptr(img), 3, H, W, 4, dc), "image_to_nhwc")
        # ---- DLA-34 body (dla.py:289-304)
        ch = (16, 32, 64, 128, 256, 512)
        x = P.conv(img[..., :3], "body.base_layer.0", P.new(H, W, ch[0]), pad=3, relu=True)
        x = P.conv(x, "body.level0.0", P.new(H, W, ch[0]), pad=1, relu=True)
        x = P.conv(x, "body.level1.0", P.new(H // 2, W // 2, ch[1]), stride=2, pad=1, relu=True)
        x2 = self._tree(P, "level2", x, 1, ch[1], ch[2], 2, False)
        x3 = self._tree(P, "level3", x2, 2, ch[2], ch[3], 2, True)
        x4 = self._tree(P, "level4", x3, 2, ch[3], ch[4], 2, True)
        x5 = self._tree(P, "level5", x4, 1, ch[4], ch[5], 2, True)
        body = [x2, x3, x4, x5]
        # ---- FPN (fpn_patch.py:29-61)
        Cc = self.C
        feats = [None] * 5
        last = None
        for i in range(4, 0, -1):
            f = body[i - 1]
            inner = P.conv(f, "fpn.fpn_inner%d" % i, P.new(f.shape[1], f.shape[2], Cc))
            if last is not None:
                P.call(L.smot_upsample_add, (ops._ptr(last), last.shape[1], last.shape[2], Cc, ops._ptr(inner),
                                             inner.shape[1], inner.shape[2], Cc, Cc, dc), "upsample_add%d" % i)
            feats[i - 1] = P.conv(inner, "fpn.fpn_layer%d" % i, P.new(f.shape[1], f.shape[2], Cc), pad=1)
            last = inner
        p5 = feats[3]
        feats[4] = P.new((p5.shape[1] - 1) // 2 + 1, (p5.shape[2] - 1) // 2 + 1, Cc)
        P.call(L.smot_subsample2, (ops._ptr(p5), ops._ptr(feats[4]), p5.shape[1], p5.shape[2], Cc, Cc, Cc, dc), "p6")
        P.feats = feats
        # ---- RPN head + selection
        R = cfg.MODEL.RPN
        A = self.n_anchor
        hld = ((5 * A + 3) // 4) * 4
        heads = []
        for l, f in enumerate(feats):
            t = P.conv(f, "rpn.conv", P.new(f.shape[1], f.shape[2], Cc), pad=1, relu=True)
            hbuf = P.new(f.shape[1], f.shape[2], hld, dtype=torch.float32)
            P.conv(t, "rpn.pred", hbuf[..., :5 * A])
            heads.append(hbuf)
        P.rpn_levels = ops.rpn_levels(heads, R.ANCHOR_STRIDE, self.cells)
        P.keep.append(P.rpn_levels)
        nprop = R.FPN_POST_NMS_TOP_N_TEST
        P.props = torch.zeros((nprop, 4), dtype=torch.float32, device=dev)
        P.prop_scores = torch.zeros((nprop,), dtype=torch.float32, device=dev)
        P.prop_count = torch.zeros((1,), dtype=torch.int32, device=dev)
        ws = ops.rpn_select_workspace(len(feats), R.PRE_NMS_TOP_N_TEST, dev)
        P.keep.append(ws)
        P.call(L.smot_rpn_select, (P.rpn_levels, len(feats), R.PRE_NMS_TOP_N_TEST, R.POST_NMS_TOP_N_TEST,
                                   R.NMS_THRESH, float(R.MIN_SIZE), nprop, W, H, int(cfg.INPUT.AMODAL),
                                   ops._ptr(P.props), ops._ptr(P.prop_scores), ops._ptr(P.prop_count), ops._ptr(ws),
                                   ws.numel()), "rpn_select")
        # ---- box head on the proposals (box_head.py:46-51, inference.py:46-191)
        P.box = self._box_buffers(nprop)
        self._box_steps(P, P.box, P.props, P.prop_count, nprop, None)
        ncls = self.ncls
        cap = nprop * (ncls - 1)
        P.det_boxes = torch.zeros((cap, 4), dtype=torch.float32, device=dev)
        P.det_scores = torch.zeros((cap,), dtype=torch.float32, device=dev)
        P.det_labels = torch.zeros((cap,), dtype=torch.int32, device=dev)
        P.det_count = torch.zeros((1,), dtype=torch.int32, device=dev)
        nws = ops.sort_nms_workspace(nprop, dev)
        P.keep.append(nws)
        P.call(lambda st: self._fill_dets(P), (), "det_init")
        H_ = cfg.MODEL.ROI_HEADS
        for j in range(1, ncls):
            P.call(L.smot_sort_nms, (C.c_void_p(P.box["dec_boxes"].data_ptr() + 16 * j), 4 * ncls,
                                     C.c_void_p(P.box["dec_scores"].data_ptr() + 4 * j), ncls, ops._ptr(P.prop_count),
                                     nprop, H_.SCORE_THRESH, H_.NMS, nprop, j, None, ops._ptr(P.det_boxes),
                                     ops._ptr(P.det_scores), ops._ptr(P.det_labels), ops._ptr(P.det_count), ops._ptr(nws),
                                     nws.numel()), "det_nms%d" % j)
        self.plans[key] = P
        return P

    def _fill_dets(self, P):
        P.det_scores.fill_(-1.0)
        P.det_count.zero_()
        return 0

    def _box_buffers(self, n):
        dev, dt, ncls = self.device, self.dtype, self.ncls
        res = self.cfg.MODEL.ROI_BOX_HEAD.POOLER_RESOLUTION
        rep = self.cfg.MODEL.ROI_BOX_HEAD.MLP_HEAD_DIM
        hld = ((5 * ncls + 3) // 4) * 4
        return dict(pooled=torch.zeros((n, res, res, self.C), dtype=dt, device=dev),
                    fc6=torch.zeros((1, 1, n, rep), dtype=dt, device=dev),
                    fc7=torch.zeros((1, 1, n, rep), dtype=dt, device=dev),
                    head=torch.zeros((1, 1, n, hld), dtype=torch.float32, device=dev),
                    dec_boxes=torch.zeros((n, ncls, 4), dtype=torch.float32, device=dev),
                    dec_scores=torch.zeros((n, ncls), dtype=torch.float32, device=dev), n=n)

    def _box_steps(self, P, B, rois, count, n, track_labels):
        """ROIAlign 7x7 -> fc6 -> fc7 -> [cls | bbox] -> softmax/decode.  Appends launches to plan P."""
        cfg, L = self.cfg, lib()
        Hh = cfg.MODEL.ROI_BOX_HEAD
        res = Hh.POOLER_RESOLUTION
        pyr = ops.make_pyramid(P.feats, Hh.POOLER_SCALES)
        P.keep.append(pyr)
        P.call(L.smot_roi_align, (C.byref(pyr), ops._ptr(rois), None, ops._ptr(count), n, self.C, res,
                                  Hh.POOLER_SAMPLING_RATIO, ops._ptr(B["pooled"]), _lib.dtype_code(self.dtype)), "box_roi_align")
        P.conv(B["pooled"].view(1, 1, n, res * res * self.C), "box.fc6", B["fc6"], relu=True)
        P.conv(B["fc6"], "box.fc7", B["fc7"], relu=True)
        P.conv(B["fc7"], "box.pred", B["head"][..., :5 * self.ncls])
        w4 = (C.c_float * 4)(*[float(w) for w in cfg.MODEL.ROI_HEADS.BBOX_REG_WEIGHTS])
        P.keep.append(w4)
        P.call(L.smot_box_decode, (ops._ptr(B["head"]), B["head"].shape[3], ops._ptr(rois), ops._ptr(count), n, self.ncls,
                                   C.byref(w4), P.W, P.H, int(cfg.INPUT.AMODAL), ops._ptr(track_labels),
                                   ops._ptr(B["dec_boxes"]), ops._ptr(B["dec_scores"])), "box_decode")

    # ------------------------------------------------------------------------------------------
    # per-frame entry points
    # ------------------------------------------------------------------------------------------
    def run_static(self, image):
        """image: (3,H,W) or (1,3,H,W) float tensor (any device).  Runs backbone..detections."""
        if image.dim() == 4:
            if image.shape[0] != 1:
                raise ValueError("one image per forward (track_core.py:75 asserts the same)")
            image = image[0]
        P = self.plan(image.shape[1], image.shape[2])
        P.img_in.copy_(image, non_blocking=True)
        with self.timed("static"):
            P.run()
        return P

    def box_head_eager(self, P, rois, track_labels=None):
        """Box head on an arbitrary host-sized set of boxes (track refinement roi_heads.py:69, given
        detections roi_heads.py:29).  Returns (dec_boxes (n,ncls,4), dec_scores (n,ncls))."""
        n = rois.shape[0]
        B = self._box_buffers(n)
        Q = _Plan(self, P.H, P.W)
        Q.feats = P.feats
        self._box_steps(Q, B, rois, None, n, track_labels)
        Q.keep.append(B)
        Q.run_eager()
        return B["dec_boxes"], B["dec_scores"]

    def emm_track(self, P, mem_feat, mem_sr, mem_boxes):
        """EMM.forward inference branch (track_core.py:28-79) for N tracks; device tensors in, device out:
        boxes (N,4), conf (N,), valid (N,) int32."""
        cfg = self.cfg
        T = cfg.MODEL.TRACK_HEAD
        n = mem_boxes.shape[0]
        with self.timed("sr_roi_align"):
            srf = ops.roi_align(P.feats, mem_sr, T.POOLER_SCALES, self.s_res, T.POOLER_SAMPLING_RATIO,
                                level_boxes=mem_boxes, pads=self.pads)
        with self.timed("xcorr"):
            resp = ops.xcorr(srf, mem_feat)
        O, Cc = self.o_res, self.C
        w, _, _ = self.weights["emm.towers"]
        tower = ops.conv2d(resp, w, pad=1)
        ops.groupnorm_relu_(tower, self.gn_gamma, self.gn_beta, 2 * self.cfg.MODEL.GROUP_NORM.NUM_GROUPS,
                            self.cfg.MODEL.GROUP_NORM.EPSILON, True)
        maps = torch.zeros((n, O, O, 8), dtype=torch.float32, device=self.device)
        w, _, b = self.weights["emm.clsctr"]
        ops.conv2d(tower[..., :Cc], w, None, b, pad=1, out=maps[..., 0:3])
        w, _, b = self.weights["emm.reg"]
        ops.conv2d(tower[..., Cc:], w, None, b, pad=1, relu=True, out=maps[..., 3:7])
        self.last_maps = maps
        return ops.emm_decode(maps, mem_sr, mem_boxes, self.hann, self.up, self.t_res, T.PAD_PIXELS,
                              T.EMM.USE_CENTERNESS, T.EMM.COSINE_WINDOW_WEIGHT, P.W, P.H, cfg.INPUT.AMODAL)

    def templates(self, P, boxes_dev):
        """EMM.extract_cache feature part (track_core.py:92): ROIAlign T x T on the unpadded pyramid."""
        T = self.cfg.MODEL.TRACK_HEAD
        return ops.roi_align(P.feats, boxes_dev, T.POOLER_SCALES, self.t_res, T.POOLER_SAMPLING_RATIO)

    def nms_workspace(self, n):
        n = max(64, 1 << (max(n, 1) - 1).bit_length())
        if n not in self._nms_ws:
            self._nms_ws[n] = ops.sort_nms_workspace(n, self.device)
        return self._nms_ws[n]
