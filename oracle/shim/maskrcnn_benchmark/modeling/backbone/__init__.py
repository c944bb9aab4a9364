# upstream's package __init__ imports the builders, which is what registers "R-50-FPN" & co. in registry.BACKBONES when
# siammot/modelling/backbone/backbone_ext.py:6 imports maskrcnn_benchmark.modeling.backbone.fpn
from . import backbone  # noqa: F401
