"""Seeded scenarios shared by the golden-vector generator, the oracle tests and the GPU parity tests."""
import os

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_DIR = os.path.join(REPO, "tests", "golden")

# name -> dict(yaml, overrides (yacs list form), H, W, frames, n_obj, clip_seed, weight_seed, inject)
SCENARIOS = {
    # default TAO-style config, natural start / suspend / resume dynamics
    "emm_256x384": dict(yaml="DLA_34_FPN_EMM.yaml", overrides=[], H=256, W=384, frames=6, n_obj=6,
                        clip_seed=0, weight_seed=1, inject=None),
    # MOT17 config: amodal (no clipping anywhere) + short dormancy so tracks expire
    "emm_amodal_expire_192x320": dict(yaml="DLA_34_FPN_EMM_MOT17.yaml",
                                      overrides=["MODEL.TRACK_HEAD.MAX_DORMANT_FRAMES", 2,
                                                 "INFERENCE.USE_GIVEN_DETECTIONS", False],
                                      H=192, W=320, frames=7, n_obj=5, clip_seed=3, weight_seed=2, inject=None),
    # BASELINE.json configs[0]: one 720p frame pair (704x1280 after the reference resize rule), 4 injected
    # tracks (cx, cy, w, h) hitting FPN levels 0,1,2,0
    "pair_720p_4tracks": dict(yaml="DLA_34_FPN_EMM.yaml", overrides=[], H=704, W=1280, frames=2, n_obj=6,
                              clip_seed=0, weight_seed=1,
                              inject=[(200., 300., 60., 150.), (500., 350., 80., 200.),
                                      (800., 352., 300., 500.), (1100., 400., 40., 100.)]),
}


# public-detection path: the reference is fed `given_detection` every frame (roi_heads.py:26-34);
# oracle-only fixtures (the engine's given-detection path is checked against the oracle in tests/test_paths_gpu.py)
GIVEN_SCENARIOS = {
    "given_det_192x320": dict(yaml="DLA_34_FPN_EMM_MOT17.yaml",
                              # thresholds lowered so that tracks start / persist / lapse on random public boxes
                              overrides=["MODEL.TRACK_HEAD.MAX_DORMANT_FRAMES", 2, "INFERENCE.USE_GIVEN_DETECTIONS", False,
                                         "MODEL.TRACK_HEAD.START_TRACK_THRESH", 0.12, "MODEL.TRACK_HEAD.TRACK_THRESH", 0.06,
                                         "MODEL.TRACK_HEAD.RESUME_TRACK_THRESH", 0.09],
                              H=192, W=320, frames=5, n_obj=5, clip_seed=3, weight_seed=2, inject=None,
                              det_seed=4, det_per_frame=(24, 24, 0, 24, 30)),   # frame 2: no public detections at all
}


# further reference-generated fixtures that pin the ORACLE only (CPU tests); the engine is compared with the oracle on the
# same switches in tests/test_paths_gpu.py
ORACLE_SCENARIOS = {
    # two foreground classes (the reference's person_vehicle models): per-class NMS, labels carried by tracks
    "emm_3class_192x320": dict(yaml="DLA_34_FPN_EMM.yaml", overrides=["MODEL.ROI_BOX_HEAD.NUM_CLASSES", 3],
                               H=192, W=320, frames=5, n_obj=5, clip_seed=5, weight_seed=3, inject=None),
    # TRACKTOR scoring switch (roi_heads.py:72-76) + centerness off (track_core.py:106-108)
    "emm_tracktor_nocenter_256x384": dict(yaml="DLA_34_FPN_EMM.yaml",
                                          overrides=["MODEL.TRACK_HEAD.TRACKTOR", True, "MODEL.TRACK_HEAD.EMM.USE_CENTERNESS", False],
                                          H=256, W=384, frames=5, n_obj=6, clip_seed=0, weight_seed=1, inject=None),
    # AOT geometry: 7x7 templates, search region r = 5 (35x35 windows, 29x29 responses, x16 ->
    # 464x464 score maps), PAD_PIXELS 256, anchors 6..96, centerness off, cosine-window weight 0.1; thresholds lowered so that
    # tracks start / lapse / resume on random weights (the shipped 0.95 / 0.6 start nothing)
    "emm_aot_geometry_256x384": dict(yaml="DLA_34_FPN_EMM_AOT.yaml",
                                     overrides=["MODEL.TRACK_HEAD.START_TRACK_THRESH", 0.45, "MODEL.TRACK_HEAD.TRACK_THRESH", 0.35,
                                                "MODEL.TRACK_HEAD.RESUME_TRACK_THRESH", 0.4, "MODEL.TRACK_HEAD.MAX_DORMANT_FRAMES", 3],
                                     H=256, W=384, frames=6, n_obj=6, clip_seed=0, weight_seed=3, inject=None),
    # BASELINE.json configs[4]: upstream maskrcnn_benchmark "R-50-FPN" body (stride on the first 1x1), 256-channel FPN / RPN / box
    # head / EMM.  The body is NOT reference code (un-vendored upstream, restated in oracle/shim/.../backbone/resnet.py and
    # cross-checked against torchvision in tests/test_oracle_resnet.py): this fixture pins everything SiamMOT-authored around it
    "emm_r50_192x320": dict(yaml="DLA_34_FPN_EMM.yaml",
                            overrides=["MODEL.BACKBONE.CONV_BODY", "R-50-FPN", "MODEL.RESNETS.BACKBONE_OUT_CHANNELS", 256,
                                       "MODEL.TRACK_HEAD.MAX_DORMANT_FRAMES", 3],
                            H=192, W=320, frames=5, n_obj=5, clip_seed=5, weight_seed=11, inject=None),
    # a deeper member of the reference's DLA family (dla.py:353-360): DLA-102 -- bottleneck blocks, trees three and four levels
    # deep, residual roots.  (MODEL.WEIGHT must name an existing path: for every body but DLA-34 the reference's own lookup of
    # the ImageNet URL raises KeyError -- dla.py:387-405 map "DLA-102-FPN" to 'dla_102' while model_urls has 'dla102'.)
    "emm_dla102_192x320": dict(yaml="DLA_34_FPN_EMM.yaml",
                               overrides=["MODEL.BACKBONE.CONV_BODY", "DLA-102-FPN", "MODEL.DLA.DLA_STAGE2_OUT_CHANNELS", 128,
                                          "MODEL.DLA.DLA_STAGE3_OUT_CHANNELS", 256, "MODEL.DLA.DLA_STAGE4_OUT_CHANNELS", 512,
                                          "MODEL.DLA.DLA_STAGE5_OUT_CHANNELS", 1024, "MODEL.TRACK_HEAD.MAX_DORMANT_FRAMES", 3,
                                          "MODEL.WEIGHT", "/dev/null"],
                               H=192, W=320, frames=5, n_obj=5, clip_seed=5, weight_seed=32, inject=None),
    # deformable stages (MODEL.DLA.STAGE_WITH_DCN, dla.py:74-78; the reference's "-DCN" models deform levels 3..5): DLA-60 whose
    # bottleneck 3x3 convs there are upstream's DFConv2d.  DFConv2d / DeformConv are NOT reference code (un-vendored upstream,
    # restated in oracle/shim/maskrcnn_benchmark/layers over torchvision.ops.deform_conv2d; the oracle's own restatement of
    # the operator is cross-checked against torchvision in tests/test_oracle_dla_family_cpu.py)
    "emm_dla60_dcn_192x320": dict(yaml="DLA_34_FPN_EMM.yaml",
                                  overrides=["MODEL.BACKBONE.CONV_BODY", "DLA-60-FPN",
                                             "MODEL.DLA.STAGE_WITH_DCN", (False, False, False, True, True, True),
                                             "MODEL.DLA.DLA_STAGE2_OUT_CHANNELS", 128, "MODEL.DLA.DLA_STAGE3_OUT_CHANNELS", 256,
                                             "MODEL.DLA.DLA_STAGE4_OUT_CHANNELS", 512, "MODEL.DLA.DLA_STAGE5_OUT_CHANNELS", 1024,
                                             "MODEL.TRACK_HEAD.MAX_DORMANT_FRAMES", 3, "MODEL.WEIGHT", "/dev/null"],
                                  H=192, W=320, frames=5, n_obj=5, clip_seed=5, weight_seed=6, inject=None),
    # class-agnostic box regression (upstream MODEL.CLS_AGNOSTIC_BBOX_REG; inference.py:66-72) with two foreground classes
    "emm_cls_agnostic_192x320": dict(yaml="DLA_34_FPN_EMM.yaml",
                                     overrides=["MODEL.ROI_BOX_HEAD.NUM_CLASSES", 3, "MODEL.CLS_AGNOSTIC_BBOX_REG", True],
                                     H=192, W=320, frames=5, n_obj=5, clip_seed=5, weight_seed=3, inject=None),
}


# Full-size scenes of the fp16 end-to-end parity tests (tests/test_fp16_e2e_gpu.py): the geometry of BASELINE.json configs[1] /
# [2] / [4] (network input 3x704x1280 or 3x1056x1920; 30 / 80 tracks injected into the memory on frame 0, then 8 frames).
# ``tweak`` reshapes the synthetic HEAD weights (tests/fp16_scene.py: apply_tweak) -- found with tools/parity_probe.py
# --calibrate -- so that every id-deciding comparison of the clip has a wide margin (tests/decisive.py measures them; the
# margins are stored in the fixture and re-checked by the test).  Expected outputs: the reference itself (make_golden.py).
FULL_SCENARIOS = {
    "full_720p30": dict(yaml="DLA_34_FPN_EMM.yaml", overrides=[], workload="720p30", H=704, W=1280, frames=9, n_obj=12,
                        clip_seed=13, weight_seed=3, tracks=30, tweak="clsx4+bg26.41+emm4"),
    # BASELINE.json configs[2]: native 1080p input (INPUT.MIN/MAX_SIZE_TEST 1080/1920 -> 3x1056x1920), 80 tracks in memory
    "full_1080p80": dict(yaml="DLA_34_FPN_EMM.yaml", overrides=["INPUT.MIN_SIZE_TEST", 1080, "INPUT.MAX_SIZE_TEST", 1920],
                         workload="1080p80", H=1056, W=1920, frames=9, n_obj=12, clip_seed=13, weight_seed=3, tracks=80,
                         tweak="clsx4+bg26.62+emm4"),
    # BASELINE.json configs[4]: upstream R-50-FPN body, 256-channel FPN / RPN / box head / EMM.  Six tracked frames: with random
    # weights the 256-channel EMM regresses the tracks towards each other, and on the seventh frame two of them overlap by
    # an IoU within 0.001 of the solver's 0.5 (tools/parity_probe.py), which no fp16 engine can be asked to reproduce
    "full_r50_720p30": dict(yaml="DLA_34_FPN_EMM.yaml",
                            overrides=["MODEL.BACKBONE.CONV_BODY", "R-50-FPN", "MODEL.RESNETS.BACKBONE_OUT_CHANNELS", 256],
                            workload="r50_720p30", H=704, W=1280, frames=7, n_obj=12, clip_seed=13, weight_seed=3, tracks=30,
                            tweak="clsx4+bg108.46+emm6"),
}


def given_boxes(sc):
    """The public detections of every frame of a GIVEN_SCENARIOS entry (seeded)."""
    g = torch.Generator().manual_seed(sc["det_seed"])
    out = []
    for n in sc["det_per_frame"]:
        xy = torch.rand(n, 2, generator=g) * torch.tensor([sc["W"] * 0.78, sc["H"] * 0.62])
        wh = torch.rand(n, 2, generator=g) * torch.tensor([50., 60.]) + 10
        out.append(torch.cat([xy, xy + wh], 1))
    return out


def inject_boxes(spec):
    b = torch.tensor(spec, dtype=torch.float32)
    return torch.stack((b[:, 0] - b[:, 2] / 2, b[:, 1] - b[:, 3] / 2,
                        b[:, 0] + b[:, 2] / 2, b[:, 1] + b[:, 3] / 2), dim=1)


def golden_path(name):
    return os.path.join(GOLDEN_DIR, name + ".pt")
