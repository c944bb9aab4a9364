"""Seeded synthetic weights with the reference state-dict layout.

There are no checkpoints offline (all are S3 downloads: /root/reference/readme/model_zoo.md), so
parity tests and the benchmark run on random weights of the real architecture.  The recipe keeps
activations O(1) through the 39 DLA convs (He-style gains; FrozenBN gamma < 1 on residual
branches) and scales the head layers so that scores spread over (0,1) and the EMM regression
outputs are tens of pixels -- i.e. detections start tracks, tracks move, get suspended and
resumed, instead of everything sitting at 0.5 as with the default initialisers.

Pure function of (cfg, seed): torch CPU generator, so the same tensors are produced in the
authoring container and on the GPU box."""
import math

import torch

DLA34_LEVELS = (1, 1, 1, 2, 2, 1)
DLA34_CHANNELS = (16, 32, 64, 128, 256, 512)


def dla34_layout():
    """Yield (kind, name, shape) for every DLA-34 parameter/buffer group, in reference order
    (dla.py:241-313).  kind in {"conv", "bn"}."""
    ch = DLA34_CHANNELS
    out = [("conv", "base_layer.0", (ch[0], 3, 7, 7)), ("bn", "base_layer.1", ch[0]),
           ("conv", "level0.0", (ch[0], ch[0], 3, 3)), ("bn", "level0.1", ch[0]),
           ("conv", "level1.0", (ch[1], ch[0], 3, 3)), ("bn", "level1.1", ch[1])]

    def block(pre, cin, cout):
        out.extend([("conv", pre + ".conv1", (cout, cin, 3, 3)), ("bn", pre + ".bn1", cout),
                    ("conv", pre + ".conv2", (cout, cout, 3, 3)), ("bn", pre + ".bn2", cout)])

    def tree(pre, levels, cin, cout, level_root, root_dim=0):
        if root_dim == 0:
            root_dim = 2 * cout
        if level_root:
            root_dim += cin
        if levels == 1:
            block(pre + ".tree1", cin, cout)
            block(pre + ".tree2", cout, cout)
            out.extend([("conv", pre + ".root.conv", (cout, root_dim, 1, 1)), ("bn", pre + ".root.bn", cout)])
        else:
            tree(pre + ".tree1", levels - 1, cin, cout, False, 0)
            tree(pre + ".tree2", levels - 1, cout, cout, False, root_dim + cout)
        if cin != cout:
            out.extend([("conv", pre + ".project.0", (cout, cin, 1, 1)), ("bn", pre + ".project.1", cout)])

    for lvl in range(2, 6):
        tree("level%d" % lvl, DLA34_LEVELS[lvl], ch[lvl - 1], ch[lvl], lvl > 2)
    return out


# The plain (non-grouped, non-DCN) members of the reference's DLA family (dla.py:307-372): tree depths per level, channels,
# block type, residual roots.  "DLA-46-XC" / "DLA-60-RES2NET" use grouped / Res2Net blocks and are not covered.
DLA_ARCHS = {
    "DLA-34-FPN": dict(levels=(1, 1, 1, 2, 2, 1), channels=(16, 32, 64, 128, 256, 512), block="basic", residual_root=False),
    "DLA-46-C-FPN": dict(levels=(1, 1, 1, 2, 2, 1), channels=(16, 32, 64, 64, 128, 256), block="bottleneck", residual_root=False),
    "DLA-60-FPN": dict(levels=(1, 1, 1, 2, 3, 1), channels=(16, 32, 128, 256, 512, 1024), block="bottleneck", residual_root=False),
    "DLA-102-FPN": dict(levels=(1, 1, 1, 3, 4, 1), channels=(16, 32, 128, 256, 512, 1024), block="bottleneck", residual_root=True),
    "DLA-169-FPN": dict(levels=(1, 1, 2, 3, 5, 1), channels=(16, 32, 128, 256, 512, 1024), block="bottleneck", residual_root=True),
}


def dla_layout(arch, dcn=(False,) * 6):
    """(kind, name, shape) of every parameter group of a DLA body in the reference's module order (dla.py:241-304): DlaTree
    registers tree1, tree2, then root (levels == 1), then project; DlaBottleneck has mid = out / 2 channels (dla.py:63-96).
    dcn[level]: MODEL.DLA.STAGE_WITH_DCN -- the bottlenecks' 3x3 becomes upstream's DFConv2d (dla.py:74-78): kind "convb"
    (conv with bias) for its offset predictor, then the deformable conv's weight.  DlaBasic ignores the flag (dla.py:33 **_)."""
    A = DLA_ARCHS[arch]
    ch, levels, bottleneck = A["channels"], A["levels"], A["block"] == "bottleneck"
    out = [("conv", "base_layer.0", (ch[0], 3, 7, 7)), ("bn", "base_layer.1", ch[0])]

    def conv_level(name, cin, cout, convs):
        for i in range(convs):
            out.extend([("conv", "%s.%d" % (name, 3 * i), (cout, cin, 3, 3)), ("bn", "%s.%d" % (name, 3 * i + 1), cout)])
            cin = cout
    conv_level("level0", ch[0], ch[0], levels[0])
    conv_level("level1", ch[0], ch[1], levels[1])

    def block(pre, cin, cout, with_dcn=False):
        if bottleneck:
            mid = cout // 2
            out.extend([("conv", pre + ".conv1", (mid, cin, 1, 1)), ("bn", pre + ".bn1", mid)])
            if with_dcn:
                out.extend([("convb", pre + ".conv2.offset", (18, mid, 3, 3)), ("conv", pre + ".conv2.conv", (mid, mid, 3, 3))])
            else:
                out.append(("conv", pre + ".conv2", (mid, mid, 3, 3)))
            out.extend([("bn", pre + ".bn2", mid),
                        ("conv", pre + ".conv3", (cout, mid, 1, 1)), ("bn", pre + ".bn3", cout)])
        else:
            out.extend([("conv", pre + ".conv1", (cout, cin, 3, 3)), ("bn", pre + ".bn1", cout),
                        ("conv", pre + ".conv2", (cout, cout, 3, 3)), ("bn", pre + ".bn2", cout)])

    def tree(pre, lv, cin, cout, level_root, with_dcn, root_dim=0):
        if root_dim == 0:
            root_dim = 2 * cout
        if level_root:
            root_dim += cin
        if lv == 1:
            block(pre + ".tree1", cin, cout, with_dcn)
            block(pre + ".tree2", cout, cout, with_dcn)
            out.extend([("conv", pre + ".root.conv", (cout, root_dim, 1, 1)), ("bn", pre + ".root.bn", cout)])
        else:
            tree(pre + ".tree1", lv - 1, cin, cout, False, with_dcn, 0)
            tree(pre + ".tree2", lv - 1, cout, cout, False, with_dcn, root_dim + cout)
        if cin != cout:
            out.extend([("conv", pre + ".project.0", (cout, cin, 1, 1)), ("bn", pre + ".project.1", cout)])

    for lvl in range(2, 6):
        tree("level%d" % lvl, levels[lvl], ch[lvl - 1], ch[lvl], lvl > 2, bool(dcn[lvl]) and bottleneck)
    return out


R50_BLOCKS = (3, 4, 6, 3)
RESNET_BLOCKS = {"R-50-FPN": R50_BLOCKS, "R-101-FPN": (3, 4, 23, 3)}   # upstream resnet.py stage specs (FPN variants, stages 2..5)


def resnet50_layout(blocks=R50_BLOCKS, stem=64, res2=256, width=64):
    """(kind, name, shape) of every ResNet-50 body parameter group in upstream module order (maskrcnn_benchmark
    modeling/backbone/resnet.py: stem, then per block downsample / conv1..3), stride on the first 1x1."""
    out = [("conv", "stem.conv1", (stem, 3, 7, 7)), ("bn", "stem.bn1", stem)]
    cin = stem
    for li, nb in enumerate(blocks):
        mid, cout = width * 2 ** li, res2 * 2 ** li
        for b in range(nb):
            pre = "layer%d.%d" % (li + 1, b)
            if cin != cout:
                out.extend([("conv", pre + ".downsample.0", (cout, cin, 1, 1)), ("bn", pre + ".downsample.1", cout)])
            out.extend([("conv", pre + ".conv1", (mid, cin, 1, 1)), ("bn", pre + ".bn1", mid),
                        ("conv", pre + ".conv2", (mid, mid, 3, 3)), ("bn", pre + ".bn2", mid),
                        ("conv", pre + ".conv3", (cout, mid, 1, 1)), ("bn", pre + ".bn3", cout)])
            cin = cout
    return out


def is_resnet(cfg):
    return cfg.MODEL.BACKBONE.CONV_BODY.startswith("R-")


def body_layout(cfg):
    """Layout of the configured body: DLA-34 (dla.py) or ResNet-50 (upstream resnet.py)."""
    body = cfg.MODEL.BACKBONE.CONV_BODY
    if body == "DLA-34-FPN":
        return dla34_layout()
    if body in DLA_ARCHS:
        return dla_layout(body, tuple(cfg.MODEL.DLA.STAGE_WITH_DCN))
    if body in RESNET_BLOCKS:
        R = cfg.MODEL.RESNETS
        return resnet50_layout(RESNET_BLOCKS[body], R.STEM_OUT_CHANNELS, R.RES2_OUT_CHANNELS, R.NUM_GROUPS * R.WIDTH_PER_GROUP)
    raise NotImplementedError("body %s (implemented: %s)" % (body, ", ".join(sorted(DLA_ARCHS) + sorted(RESNET_BLOCKS))))


def backbone_channels(cfg):
    """(FPN input channels per stage, FPN output channels) -- backbone_ext.py:17-23 / upstream build_resnet_fpn_backbone;
    the EMM head takes the same output width (feature_extractor.py:47-52)."""
    if is_resnet(cfg):
        c2 = cfg.MODEL.RESNETS.RES2_OUT_CHANNELS
        return (c2, c2 * 2, c2 * 4, c2 * 8), cfg.MODEL.RESNETS.BACKBONE_OUT_CHANNELS
    D = cfg.MODEL.DLA
    return ((D.DLA_STAGE2_OUT_CHANNELS, D.DLA_STAGE3_OUT_CHANNELS, D.DLA_STAGE4_OUT_CHANNELS, D.DLA_STAGE5_OUT_CHANNELS),
            D.BACKBONE_OUT_CHANNELS)


def make_state_dict(cfg, seed=1):
    g = torch.Generator().manual_seed(seed)

    def randn(*shape, std=1.0):
        return torch.randn(*shape, generator=g) * std

    def rand(*shape):
        return torch.rand(*shape, generator=g)

    sd = {}
    resnet = is_resnet(cfg)
    for kind, name, shape in body_layout(cfg):
        key = "backbone.body." + name
        if kind == "conv":
            fan_in = shape[1] * shape[2] * shape[3]
            sd[key + ".weight"] = randn(*shape, std=math.sqrt(2.0 / fan_in))
        elif kind == "convb":   # DFConv2d's offset predictor: offsets of a fraction of a pixel up to ~2 px
            fan_in = shape[1] * shape[2] * shape[3]
            sd[key + ".weight"] = randn(*shape, std=0.7 / math.sqrt(fan_in))
            sd[key + ".bias"] = randn(shape[0], std=0.3)
        else:
            # the branch that is added to the identity gets a smaller gain so that activations stay O(1) with depth
            # (16 bottleneck blocks in the ResNet: a much smaller gain than for DLA's 8 basic blocks)
            arch = DLA_ARCHS.get(cfg.MODEL.BACKBONE.CONV_BODY, {})
            bottleneck = resnet or arch.get("block") == "bottleneck"
            residual_branch = (name.endswith("bn3") or (".project" in name and not resnet)) if bottleneck else \
                (name.endswith("bn2") or ".project" in name)
            gain = ((0.25 if bottleneck else 0.6) if residual_branch else 0.9)
            if arch.get("residual_root") and name.endswith("root.bn"):
                gain = 0.4                                  # DlaRoot adds its first input (dla.py:185-186): damp the conv branch
            sd[key + ".weight"] = gain + 0.1 * rand(shape)
            sd[key + ".bias"] = randn(shape, std=0.1)
            sd[key + ".running_mean"] = randn(shape, std=0.1)
            sd[key + ".running_var"] = 1.0 + 0.1 * rand(shape)
    stage_channels, C = backbone_channels(cfg)
    for i, cin in enumerate(stage_channels, 1):
        sd["backbone.fpn.fpn_inner%d.weight" % i] = randn(C, cin, 1, 1, std=math.sqrt(1.0 / cin))
        sd["backbone.fpn.fpn_inner%d.bias" % i] = randn(C, std=0.1)
        sd["backbone.fpn.fpn_layer%d.weight" % i] = randn(C, C, 3, 3, std=math.sqrt(1.0 / (9 * C)))
        sd["backbone.fpn.fpn_layer%d.bias" % i] = randn(C, std=0.1)
    A = len(cfg.MODEL.RPN.ASPECT_RATIOS)
    sd["rpn.head.conv.weight"] = randn(C, C, 3, 3, std=math.sqrt(2.0 / (9 * C)))
    sd["rpn.head.conv.bias"] = randn(C, std=0.1)
    sd["rpn.head.cls_logits.weight"] = randn(A, C, 1, 1, std=0.4 / math.sqrt(C))
    sd["rpn.head.cls_logits.bias"] = randn(A, std=0.1) - 1.0
    sd["rpn.head.bbox_pred.weight"] = randn(4 * A, C, 1, 1, std=0.3 / math.sqrt(C))
    sd["rpn.head.bbox_pred.bias"] = randn(4 * A, std=0.05)
    H = cfg.MODEL.ROI_BOX_HEAD
    d_in, rep, ncls = C * H.POOLER_RESOLUTION ** 2, H.MLP_HEAD_DIM, H.NUM_CLASSES
    pre = "roi_heads.box."
    sd[pre + "feature_extractor.fc6.weight"] = randn(rep, d_in, std=math.sqrt(2.0 / d_in))
    sd[pre + "feature_extractor.fc6.bias"] = randn(rep, std=0.1)
    sd[pre + "feature_extractor.fc7.weight"] = randn(rep, rep, std=math.sqrt(2.0 / rep))
    sd[pre + "feature_extractor.fc7.bias"] = randn(rep, std=0.1)
    sd[pre + "predictor.cls_score.weight"] = randn(ncls, rep, std=2.0 / math.sqrt(rep))
    sd[pre + "predictor.cls_score.bias"] = randn(ncls, std=0.1) + torch.tensor([3.5] + [0.0] * (ncls - 1))
    nreg = 2 if cfg.MODEL.CLS_AGNOSTIC_BBOX_REG else ncls
    sd[pre + "predictor.bbox_pred.weight"] = randn(4 * nreg, rep, std=0.5 / math.sqrt(rep))
    sd[pre + "predictor.bbox_pred.bias"] = randn(4 * nreg, std=0.05)
    pre = "roi_heads.track.tracker.predictor."
    for tower in ("cls_tower", "reg_tower"):
        sd[pre + tower + ".0.weight"] = randn(C, C, 3, 3, std=math.sqrt(2.0 / (9 * C)))
        sd[pre + tower + ".1.weight"] = 0.9 + 0.2 * rand(C)
        sd[pre + tower + ".1.bias"] = randn(C, std=0.1)
    sd[pre + "cls.weight"] = randn(2, C, 3, 3, std=1.5 / math.sqrt(9 * C))
    sd[pre + "cls.bias"] = torch.tensor([0.0, 1.0])
    sd[pre + "center.weight"] = randn(1, C, 3, 3, std=1.0 / math.sqrt(9 * C))
    sd[pre + "center.bias"] = torch.tensor([1.0])
    sd[pre + "reg.weight"] = randn(4, C, 3, 3, std=6.0 / math.sqrt(9 * C))
    sd[pre + "reg.bias"] = torch.tensor([30.0, 60.0, 30.0, 60.0])
    return sd
