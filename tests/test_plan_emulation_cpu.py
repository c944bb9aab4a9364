"""Host-side wiring of the static launch plan, checked without a GPU: the plan is built over host buffers and its backbone +
FPN calls are interpreted with torch CPU ops (tests/plan_emulator.py), then compared with the oracle's FPN maps.  The DLA-34
case (whose GPU execution is validated by the e2e parity tests) validates the interpreter; the R-50-FPN case is the check of
the wiring written without GPU access."""
import pytest
import torch

from helpers import scenario_inputs
from plan_emulator import build_engine_on_host, run_backbone


@pytest.mark.parametrize("name", ["emm_amodal_expire_192x320", "emm_r50_192x320"])
def test_backbone_and_fpn_wiring_matches_oracle(name, monkeypatch):
    from oracle.siammot_oracle import OracleSiamMOT
    cfg, sd, clip = scenario_inputs(name)
    cfg.DTYPE = "float32"
    eng = build_engine_on_host(cfg, sd, monkeypatch)
    image = clip[0]
    P = eng.plan(image.shape[1], image.shape[2])
    n = run_backbone(P, image)
    assert n >= 50, n
    ref = OracleSiamMOT(cfg, sd).features(image)
    assert len(P.feats) == len(ref) == 5
    for l, (got, want) in enumerate(zip(P.feats, ref)):
        got = got.permute(0, 3, 1, 2)
        assert got.shape == want.shape, (l, got.shape, want.shape)
        err = float((got - want).abs().max() / want.abs().max())
        assert err <= 1e-4, "FPN level %d: relative error %g" % (l, err)
