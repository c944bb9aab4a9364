from .rcnn import SiamMOT, build_siammot  # noqa: F401
