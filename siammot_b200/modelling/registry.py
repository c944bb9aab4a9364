"""Operator registries with the reference's names (/root/reference/siammot/utils/registry.py:3-4 and
upstream maskrcnn_benchmark.modeling.registry.BACKBONES).  A tracker registered under
``SIAMESE_TRACKER[name]`` is built as ``Tracker(cfg, track_utils)`` and selected by
``cfg.MODEL.TRACK_HEAD.MODEL`` exactly as in track_head.py:118-124."""


class Registry(dict):
    def register(self, name, fn=None):
        if fn is not None:
            self[name] = fn
            return fn

        def deco(f):
            self[name] = f
            return f
        return deco


SIAMESE_TRACKER = Registry()
TRACKER_SAMPLER = Registry()
BACKBONES = Registry()
