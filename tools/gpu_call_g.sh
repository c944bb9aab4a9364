#!/bin/bash
# Round-2 GPU call G (1 GPU): validation of the committed defaults (frame pairs on, helper thread off, multi-row ROIAlign,
# FeaturesView, INTEGRATION.md stubs) + the three bench workloads with the CPU arm.
set +e
OUT=gpurun_out/r02g
mkdir -p "$OUT"
timeout 1200 python -m pytest tests -q -m gpu --durations=8 > "$OUT/pytest_gpu.txt" 2>&1
echo "rc=$?" >> "$OUT/pytest_gpu.txt"
timeout 600 python -m pytest tests/test_fp16_e2e_gpu.py -q -s > "$OUT/pytest_fp16_e2e.txt" 2>&1
echo "rc=$?" >> "$OUT/pytest_fp16_e2e.txt"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > "$OUT/smoke.txt" 2>&1
echo "rc=$?" >> "$OUT/smoke.txt"
timeout 600 python bench.py > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"
timeout 400 python bench.py --steps 100 --warmup 10 --workload 1080p80 --no-cpu-baseline > "$OUT/bench_1080p80.json" 2> "$OUT/bench_1080p80.err"
timeout 400 python bench.py --steps 100 --warmup 10 --workload r50_720p30 --no-cpu-baseline > "$OUT/bench_r50_720p30.json" 2> "$OUT/bench_r50_720p30.err"
tail -n 12 "$OUT/pytest_gpu.txt"
grep -n "float16:\|float32:\|passed\|failed\|Error\|rc=" "$OUT/pytest_fp16_e2e.txt" | head -20
tail -n 3 "$OUT/smoke.txt"
for f in "$OUT"/bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    if d.get("impl") == "reference":
        print(json.dumps(d)[:600])
    else:
        print("value", d["value"], "e2e", d["e2e"]["value"], "per_frame", d["e2e"]["per_frame_call"]["value"], "static", d["stage_ms"]["static_graph"], d["e2e"]["clip_error"])
        print("   roofline", d["roofline"], "cpu", d.get("cpu_baseline"), "clocks", d.get("clocks"))
        print("   spread", d["spread"])
except Exception as e:
    print("ERR", e)
PY
done
