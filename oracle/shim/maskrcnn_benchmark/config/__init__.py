from .defaults import _C as cfg
