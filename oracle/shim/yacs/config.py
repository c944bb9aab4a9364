"""``yacs`` is not installed in this image; the reference's ``siammot/configs/defaults.py`` only
needs ``CfgNode``, which the product ships (same semantics: attribute access, clone, merge_from_file,
merge_from_list, freeze).  TEST INFRASTRUCTURE shim: re-exports it under the yacs name."""
from siammot_b200.config import CfgNode  # noqa: F401
