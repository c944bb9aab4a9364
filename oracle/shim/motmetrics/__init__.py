"""``motmetrics`` is not installable offline; siammot/eval/eval_clears_mot.py imports it at module level and inferencer.py imports
that module.  Empty stand-in (the evaluation itself is out of scope).  TEST INFRASTRUCTURE."""
