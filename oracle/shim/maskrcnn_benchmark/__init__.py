"""Pure-PyTorch CPU stand-in for the parts of facebookresearch/maskrcnn-benchmark that
SiamMOT's inference path imports.  TEST INFRASTRUCTURE: it exists so the
reference's own ``siammot/modelling/**`` runs unmodified from /root/reference in the
authoring container and produces the golden vectors under tests/golden/.  It is a
restatement of the upstream *semantics* (built on oracle/prims.py), not a copy."""
