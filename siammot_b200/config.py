"""Configuration: a small yacs-compatible ``CfgNode`` plus the defaults of the reference.

Mirrors the reference's config surface so its yaml files load unchanged:
upstream maskrcnn_benchmark defaults (subset actually read on the inference path)
overlaid with /root/reference/siammot/configs/defaults.py:5-109 (same keys, same values).
Unknown keys in a yaml/override are accepted and stored (the reference's training-only keys are
not all enumerated here)."""
import copy

import yaml


def _literal(text):
    """A Python literal written in a yaml value / command-line override ("(30000, 40000)", "0.5", "True"), as yacs decodes
    them: ast.literal_eval -- never eval -- falling back to the raw string."""
    import ast
    try:
        return ast.literal_eval(text)
    except (ValueError, SyntaxError):
        return text


class CfgNode(dict):
    def __init__(self, init_dict=None):
        super().__init__()
        self.__dict__["_frozen"] = False
        for k, v in (init_dict or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, name):
        if name in self:
            return self[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if self.__dict__.get("_frozen"):
            raise AttributeError("attempt to modify a frozen CfgNode: {}".format(name))
        self[name] = value

    def clone(self):
        return copy.deepcopy(self)

    def __deepcopy__(self, memo):
        out = CfgNode()
        for k, v in self.items():
            dict.__setitem__(out, k, copy.deepcopy(v, memo))
        return out

    def freeze(self):
        self._set_frozen(True)

    def defrost(self):
        self._set_frozen(False)

    def _set_frozen(self, flag):
        self.__dict__["_frozen"] = flag
        for v in self.values():
            if isinstance(v, CfgNode):
                v._set_frozen(flag)

    @staticmethod
    def _coerce(old, new):
        if isinstance(new, str) and isinstance(old, (tuple, list)):
            new = _literal(new)  # yaml has no tuple syntax: "(30000, 40000)" arrives as a string
        if isinstance(old, tuple) and isinstance(new, list):
            new = tuple(new)
        elif isinstance(old, list) and isinstance(new, tuple):
            new = list(new)
        elif isinstance(old, float) and isinstance(new, int) and not isinstance(new, bool):
            new = float(new)
        return new

    def _merge(self, other):
        if self.__dict__.get("_frozen"):
            raise AttributeError("attempt to modify a frozen CfgNode")
        for k, v in other.items():
            if isinstance(v, dict):
                if k not in self or not isinstance(self[k], CfgNode):
                    dict.__setitem__(self, k, CfgNode())
                self[k]._merge(v)
            else:
                dict.__setitem__(self, k, self._coerce(self[k], v) if k in self else v)

    def merge_from_file(self, path):
        with open(path, "r") as f:
            self._merge(yaml.safe_load(f) or {})

    def merge_from_other_cfg(self, other):
        self._merge(other)

    def merge_from_list(self, lst):
        if len(lst) % 2:
            raise ValueError("override list must be [key, value, key, value, ...]")
        for key, val in zip(lst[0::2], lst[1::2]):
            node = self
            parts = key.split(".")
            for p in parts[:-1]:
                node = node[p]
            if isinstance(val, str):
                val = _literal(val)
            node._merge({parts[-1]: val})


_DEFAULTS = {
    "MODEL": {
        "META_ARCHITECTURE": "GeneralizedRCNN", "DEVICE": "cuda", "WEIGHT": "", "RPN_ONLY": False,
        "MASK_ON": False, "KEYPOINT_ON": False, "RETINANET_ON": False, "CLS_AGNOSTIC_BBOX_REG": False,
        "BOX_ON": True, "TRACK_ON": True,
        "BACKBONE": {"CONV_BODY": "DLA-34-FPN", "FREEZE_CONV_BODY_AT": 2},
        "FPN": {"USE_GN": False, "USE_RELU": False},
        "GROUP_NORM": {"DIM_PER_GP": -1, "NUM_GROUPS": 32, "EPSILON": 1e-5},
        # upstream maskrcnn_benchmark defaults (the "R-50-FPN" body of BASELINE.json configs[4])
        "RESNETS": {"NUM_GROUPS": 1, "WIDTH_PER_GROUP": 64, "STRIDE_IN_1X1": True, "TRANS_FUNC": "BottleneckWithFixedBatchNorm",
                    "STEM_FUNC": "StemWithFixedBatchNorm", "RES5_DILATION": 1, "BACKBONE_OUT_CHANNELS": 1024,
                    "RES2_OUT_CHANNELS": 256, "STEM_OUT_CHANNELS": 64, "STAGE_WITH_DCN": (False, False, False, False)},
        "DLA": {"DLA_STAGE2_OUT_CHANNELS": 64, "DLA_STAGE3_OUT_CHANNELS": 128, "DLA_STAGE4_OUT_CHANNELS": 256,
                "DLA_STAGE5_OUT_CHANNELS": 512, "BACKBONE_OUT_CHANNELS": 128,
                "STAGE_WITH_DCN": (False, False, False, False, False, False)},
        "RPN": {"USE_FPN": True, "ANCHOR_SIZES": (32, 64, 128, 256, 512), "ANCHOR_STRIDE": (4, 8, 16, 32, 64),
                "ASPECT_RATIOS": (0.5, 1.0, 2.0), "STRADDLE_THRESH": 0, "FG_IOU_THRESHOLD": 0.7,
                "BG_IOU_THRESHOLD": 0.3, "BATCH_SIZE_PER_IMAGE": 256, "POSITIVE_FRACTION": 0.5,
                "PRE_NMS_TOP_N_TRAIN": 2000, "PRE_NMS_TOP_N_TEST": 1000, "POST_NMS_TOP_N_TRAIN": 2000,
                "POST_NMS_TOP_N_TEST": 300, "NMS_THRESH": 0.7, "MIN_SIZE": 0, "FPN_POST_NMS_TOP_N_TRAIN": 2000,
                "FPN_POST_NMS_TOP_N_TEST": 300, "FPN_POST_NMS_PER_BATCH": True, "RPN_HEAD": "SingleConvRPNHead"},
        "ROI_HEADS": {"USE_FPN": True, "FG_IOU_THRESHOLD": 0.5, "BG_IOU_THRESHOLD": 0.5,
                      "BBOX_REG_WEIGHTS": (10.0, 10.0, 5.0, 5.0), "BATCH_SIZE_PER_IMAGE": 256,
                      "POSITIVE_FRACTION": 0.25, "SCORE_THRESH": 0.05, "NMS": 0.5, "DETECTIONS_PER_IMG": 100},
        "ROI_BOX_HEAD": {"FEATURE_EXTRACTOR": "FPN2MLPFeatureExtractor", "PREDICTOR": "FPNPredictor",
                         "POOLER_RESOLUTION": 7, "POOLER_SAMPLING_RATIO": 2,
                         "POOLER_SCALES": (0.25, 0.125, 0.0625, 0.03125), "NUM_CLASSES": 2, "MLP_HEAD_DIM": 1024,
                         "USE_GN": False, "DILATION": 1, "CONV_HEAD_DIM": 256, "NUM_STACKED_CONVS": 4},
        "TRACK_HEAD": {"TRACKTOR": False, "POOLER_SCALES": (0.25, 0.125, 0.0625, 0.03125),
                       "POOLER_RESOLUTION": 15, "POOLER_SAMPLING_RATIO": 2, "PAD_PIXELS": 512,
                       "SEARCH_REGION": 2.0, "MINIMUM_SREACH_REGION": 0, "MODEL": "EMM",
                       "TRACK_THRESH": 0.4, "START_TRACK_THRESH": 0.6, "RESUME_TRACK_THRESH": 0.4,
                       "MAX_DORMANT_FRAMES": 1, "PROPOSAL_PER_IMAGE": 256, "FG_IOU_THRESHOLD": 0.65,
                       "BG_IOU_THRESHOLD": 0.35,
                       "IMM": {"FC_HEAD_DIM_MULTIPLIER": 2, "FC_HEAD_DIM": 256},
                       "EMM": {"USE_CENTERNESS": True, "POS_RATIO": 0.25, "HN_RATIO": 0.25,
                               "TRACK_LOSS_WEIGHT": 1.0, "CLS_POS_REGION": 0.8, "COSINE_WINDOW_WEIGHT": 0.4}},
    },
    "INPUT": {"MIN_SIZE_TRAIN": (800,), "MAX_SIZE_TRAIN": 1333, "MIN_SIZE_TEST": 800, "MAX_SIZE_TEST": 1333,
              "PIXEL_MEAN": [102.9801, 115.9465, 122.7717], "PIXEL_STD": [1.0, 1.0, 1.0], "TO_BGR255": True,
              "BRIGHTNESS": 0.0, "CONTRAST": 0.0, "SATURATION": 0.0, "HUE": 0.0,
              "HORIZONTAL_FLIP_PROB_TRAIN": 0.5, "VERTICAL_FLIP_PROB_TRAIN": 0.0,
              "MOTION_LIMIT": 0.1, "COMPRESSION_LIMIT": 50, "MOTION_BLUR_PROB": 0.5, "AMODAL": False},
    "VIDEO": {"TEMPORAL_WINDOW": 8, "TEMPORAL_SAMPLING": 4, "RANDOM_FRAMES_PER_CLIP": 2},
    "INFERENCE": {"USE_GIVEN_DETECTIONS": False, "CLIP_LEN": 1},
    "DATASETS": {"TRAIN": (), "TEST": (), "ROOT_DIR": ""},
    "DATALOADER": {"NUM_WORKERS": 4, "SIZE_DIVISIBILITY": 0, "ASPECT_RATIO_GROUPING": True},
    "SOLVER": {"MAX_ITER": 40000, "BASE_LR": 0.001, "WEIGHT_DECAY": 0.0005, "STEPS": (30000,),
               "CHECKPOINT_PERIOD": 5000, "IMS_PER_BATCH": 16, "VIDEO_CLIPS_PER_BATCH": 16},
    "TEST": {"EXPECTED_RESULTS": [], "IMS_PER_BATCH": 8, "DETECTIONS_PER_IMG": 100, "BBOX_AUG": {"ENABLED": False}},
    "OUTPUT_DIR": ".",
    # "float32" is the reference's arithmetic; "float16" selects fp16 storage / tensor-core convs.
    "DTYPE": "float32",
}

cfg = CfgNode(_DEFAULTS)


def get_cfg():
    """A fresh, unfrozen copy of the defaults (the reference mutates a global; this avoids that)."""
    return CfgNode(_DEFAULTS)
