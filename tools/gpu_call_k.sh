#!/bin/bash
# Round-2 GPU call K (1 GPU): full validation of the state with HaloPlan + conflict-free template scatter; all bench workloads; launch list.
set +e
OUT=gpurun_out/r02k
mkdir -p "$OUT"
timeout 1200 python -m pytest tests -q -m gpu --durations=5 > "$OUT/pytest_gpu.txt" 2>&1
echo "rc=$?" >> "$OUT/pytest_gpu.txt"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > "$OUT/smoke.txt" 2>&1
echo "rc=$?" >> "$OUT/smoke.txt"
timeout 600 python bench.py > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"
timeout 400 python bench.py --steps 100 --warmup 10 --workload 1080p80 --no-cpu-baseline --experimental off > "$OUT/bench_1080p80.json" 2> "$OUT/bench_1080p80.err"
timeout 400 python bench.py --steps 100 --warmup 10 --workload r50_720p30 --no-cpu-baseline --experimental off > "$OUT/bench_r50_720p30.json" 2> "$OUT/bench_r50_720p30.err"
timeout 300 python tools/xcorr_lab.py --out "$OUT/xcorr_lab.json" > "$OUT/xcorr_lab.log" 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file "$OUT/launches_720p30.csv" \
    python tools/run_frames.py --frames 3 --eager > "$OUT/ncu_launches.log" 2>&1
python tools/launch_report.py "$OUT/launches_720p30.csv" > "$OUT/launches_720p30_summary.txt" 2>&1
tail -n 9 "$OUT/pytest_gpu.txt"; tail -n 2 "$OUT/smoke.txt"
grep "'mma_mode': 1" "$OUT/xcorr_lab.log" | cut -c1-200
for f in "$OUT"/bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    if d.get("impl") == "reference":
        print("reference value", d["value"], d["cpu_baseline"])
    else:
        print("value", d["value"], "e2e", d["e2e"]["value"], "per_frame", d["e2e"]["per_frame_call"]["value"], "static", d["stage_ms"]["static_graph"], "xcorr us", d["roofline"]["us_per_launch"], "frac", d["roofline"]["frac"], "traffic", d["roofline"]["traffic"], d["e2e"]["clip_error"])
        print("   spread", d["spread"]["value_fps"], d["spread"]["e2e_fps"], "clocks", d.get("clocks"), "launches", d.get("gpu_launches"))
except Exception as e:
    print("ERR", e)
PY
done
head -8 "$OUT/launches_720p30_summary.txt"
