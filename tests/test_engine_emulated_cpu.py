"""The engine's host side, end to end, without a GPU: build_siammot(cfg) -> model(frame) runs over tests/cabi_emulator.py
(every libsmot entry point emulated with the oracle's primitives on the pointers the engine passes) and must reproduce the
goldens the reference itself produced -- ids / labels bit-exact, boxes <= 1e-3 px -- on every scenario, including the ones
whose GPU run is still pending (two foreground classes, TRACKTOR, AOT geometry, R-50-FPN body).  This pins launch plans,
arenas, result-block packing, the host solver and the next-frame memory; the CUDA kernels are what the `-m gpu` tests pin."""
import pytest
import torch

import cabi_emulator
from helpers import load_golden, scenario_inputs
from scenarios import ORACLE_SCENARIOS, SCENARIOS, inject_boxes

BOX_TOL, SCORE_TOL = 1e-3, 1e-3


def _run(name, monkeypatch, clip_api=False, env=None):
    from siammot_b200.modelling import build_siammot
    for k, v in (env or {}).items():
        monkeypatch.setenv(k, v)
    fake = cabi_emulator.install(monkeypatch)
    sc = SCENARIOS.get(name) or ORACLE_SCENARIOS[name]
    cfg, sd, clip = scenario_inputs(name)
    cfg.DTYPE = "float32"
    model = build_siammot(cfg)
    model.load_state_dict(sd, strict=False)
    model.eval()
    model.reset_siammot_status()
    out, start = [], 0
    pool = model.roi_heads.track.track_pool
    if sc["inject"] is not None:
        eng = model.engine()
        P = eng.run_static(clip[0])
        pool.reset()
        boxes = inject_boxes(sc["inject"])
        ids = torch.tensor([pool.start_track() for _ in range(len(boxes))])
        model.flush_memory(model.roi_heads._build_memory(P, boxes.numpy(), ids.numpy(), torch.ones(len(boxes), dtype=torch.int64).numpy()))
        pool.increment_frame()
        start = 1
    frames = [clip[t] for t in range(start, sc["frames"])]
    if clip_api:
        states = []
        results = model.forward_clip(frames, before_frame=lambda t: states.append((sorted(pool.get_active_ids()), sorted(pool._dormant_ids))))
        states = states[1:] + [(sorted(pool.get_active_ids()), sorted(pool._dormant_ids))]
    else:
        results, states = [], []
        for f in frames:
            results.append(model(f)[0])
            states.append((sorted(pool.get_active_ids()), sorted(pool._dormant_ids)))
    for r, (act, dor) in zip(results, states):
        out.append(dict(boxes=r.bbox, scores=r.get_field("scores"), ids=r.get_field("ids"), labels=r.get_field("labels"),
                        active=act, dormant=dor))
    return out, fake


def _compare(gold, got):
    assert len(got) == len(gold)
    for t, (g, o) in enumerate(zip(gold, got)):
        assert o["boxes"].shape == g["boxes"].shape, "frame %d: %d boxes vs %d" % (t, o["boxes"].shape[0], g["boxes"].shape[0])
        assert torch.equal(o["ids"], g["ids"]), "frame %d: track ids differ" % t
        assert torch.equal(o["labels"], g["labels"]), "frame %d: labels differ" % t
        if g["boxes"].numel():
            assert float((o["boxes"] - g["boxes"]).abs().max()) <= BOX_TOL, "frame %d boxes" % t
            assert float((o["scores"] - g["scores"]).abs().max()) <= SCORE_TOL, "frame %d scores" % t
        assert o["active"] == g["active"] and o["dormant"] == g["dormant"], "frame %d pool state" % t


SMALL = [n for n in list(SCENARIOS) + list(ORACLE_SCENARIOS) if n != "pair_720p_4tracks"]


@pytest.mark.parametrize("name", SMALL)
def test_emulated_engine_matches_reference_golden(name, monkeypatch):
    got, fake = _run(name, monkeypatch)
    _compare(load_golden(name)["frames"], got)
    assert fake.calls.get("smot_xcorr", 0) > 0 and fake.calls.get("smot_sort_nms", 0) > 0


def test_emulated_engine_720p_pair_with_injected_tracks(monkeypatch):
    """BASELINE.json configs[0] (704x1280, four injected tracks on FPN levels 0,1,2,0)."""
    got, _ = _run("pair_720p_4tracks", monkeypatch)
    _compare(load_golden("pair_720p_4tracks")["frames"], got)


def test_emulated_forward_clip_equals_golden(monkeypatch):
    """The clip API's bookkeeping (double-buffered plans, next_P for the memory) on the host."""
    got, _ = _run("emm_amodal_expire_192x320", monkeypatch, clip_api=True)
    _compare(load_golden("emm_amodal_expire_192x320")["frames"], got)


def test_emulated_planar_window_exchange_wiring(monkeypatch):
    """Host wiring of the SMOT_XCORR_PLANAR switch (arena buffer, the two planar calls in the track plan): same goldens.
    The emulation is fp32, so the dtype condition of the switch is lifted for this test only."""
    from siammot_b200 import engine
    monkeypatch.setattr(engine.Engine, "xcorr_planar_ok", lambda self: self.xcorr_planar and self.s_res == 30 and self.t_res == 15)
    got, fake = _run("emm_256x384", monkeypatch, env={"SMOT_XCORR_PLANAR": "1"})
    _compare(load_golden("emm_256x384")["frames"], got)
    assert fake.calls.get("smot_xcorr_planar", 0) > 0 and fake.calls.get("smot_roi_align_planar", 0) > 0
    assert fake.calls.get("smot_xcorr", 0) == 0


@pytest.mark.parametrize("slots", ["2", "3"])
def test_emulated_three_stage_clip_equals_golden(slots, monkeypatch):
    """SMOT_CLIP_SPLIT=1: the static plan cut at the proposal selection, K plan copies -- same results (this checks the
    bookkeeping: plan slices, slot rotation, next_P; the stream / event ordering can only be exercised on a GPU)."""
    env = {"SMOT_CLIP_SPLIT": "1", "SMOT_CLIP_SLOTS": slots}
    for name in ("emm_amodal_expire_192x320", "emm_256x384"):
        got, fake = _run(name, monkeypatch, clip_api=True, env=env)
        _compare(load_golden(name)["frames"], got)
