"""GPU parity cases whose fixtures were added after the round's GPU budget was spent (run last: the file name sorts after
every other test file, so whatever happens here cannot disturb the validated tests).  Expected outputs come from the reference
itself (tests/golden/make_golden.py, ORACLE_SCENARIOS); the CPU oracle is pinned to the same fixtures in
tests/test_oracle_golden.py."""
import pytest
import torch

from helpers import load_golden
from scenarios import ORACLE_SCENARIOS
from test_e2e_gpu import BOX_TOL, SCORE_TOL, run_engine_scenario

pytestmark = pytest.mark.gpu


@pytest.mark.xfail(strict=False, reason="fixtures added after this round's GPU budget was spent: the oracle is pinned to them on "
                                       "the CPU (tests/test_oracle_golden.py); the engine's first GPU run on them is pending")
@pytest.mark.parametrize("name", list(ORACLE_SCENARIOS))
def test_engine_fp32_matches_reference_golden_more_switches(name):
    """Two foreground classes; TRACKTOR scoring + centerness off; the AOT geometry (7x7 templates, 35x35 search windows,
    29x29 responses, PAD_PIXELS 256) -- expected outputs from the reference itself."""
    gold = load_golden(name)["frames"]
    got = run_engine_scenario(name, "float32")
    assert len(got) == len(gold)
    for t, (g, o) in enumerate(zip(gold, got)):
        assert o["boxes"].shape == g["boxes"].shape, "frame %d: %d boxes vs %d" % (t, o["boxes"].shape[0], g["boxes"].shape[0])
        assert torch.equal(o["ids"], g["ids"]), "frame %d: track ids differ" % t
        assert torch.equal(o["labels"], g["labels"])
        assert float((o["boxes"] - g["boxes"]).abs().max()) <= BOX_TOL, "frame %d boxes" % t
        assert float((o["scores"] - g["scores"]).abs().max()) <= SCORE_TOL, "frame %d scores" % t
        assert o["active"] == g["active"] and o["dormant"] == g["dormant"]
