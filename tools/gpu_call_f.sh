#!/bin/bash
# Round-2 GPU call F (1 GPU): helper-thread clip pipeline, host optimisations, multi-row ROIAlign CTAs; R-50 scene probe.
set +e
OUT=gpurun_out/r02f
mkdir -p "$OUT"
timeout 900 python -m pytest tests -q -m gpu > "$OUT/pytest_gpu.txt" 2>&1
echo "rc=$?" >> "$OUT/pytest_gpu.txt"
B="--steps 100 --warmup 10 --experimental off --no-cpu-baseline"
timeout 300 python bench.py $B > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"
SMOT_CLIP_THREAD=0 timeout 300 python bench.py $B > "$OUT/bench_nothread.json" 2> "$OUT/bench_nothread.err"
SMOT_CLIP_PAIRS=1 timeout 300 python bench.py $B > "$OUT/bench_pairs.json" 2> "$OUT/bench_pairs.err"
SMOT_CLIP_SLOTS=4 timeout 300 python bench.py $B > "$OUT/bench_k4.json" 2> "$OUT/bench_k4.err"
timeout 300 python bench.py --steps 20 --warmup 5 --experimental off --no-cpu-baseline > "$OUT/bench_k20.json" 2> "$OUT/bench_k20.err"
timeout 300 python tools/parity_probe.py --workload r50_720p30 --frames 4 --variants "$PROBE_VARIANTS" --out "$OUT/probe_r50.json" > "$OUT/probe_r50.log" 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file "$OUT/launches_720p30.csv" \
    python tools/run_frames.py --frames 3 --eager > "$OUT/ncu_launches.log" 2>&1
python tools/launch_report.py "$OUT/launches_720p30.csv" > "$OUT/launches_720p30_summary.txt" 2>&1
tail -n 6 "$OUT/pytest_gpu.txt"
for f in "$OUT"/bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("value", d["value"], "e2e", d["e2e"]["value"], "per_frame", d["e2e"]["per_frame_call"]["value"], "static", d["stage_ms"]["static_graph"], d["e2e"]["clip_error"])
    print("   host", d["stage_ms"].get("host_per_frame_ms"), "spread", d["spread"]["value_fps"], d["spread"]["e2e_fps"])
except Exception as e:
    print("ERR", e)
PY
done
python - "$OUT/probe_r50.log" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    try: r=json.loads(l)
    except Exception: continue
    c=r.get("float16",{})
    print(r["variant"], "min_margin", r["min_margin"], "fp16 ids_exact", c.get("ids_exact_all_frames"), "box", c.get("max_box_err"), "score", c.get("max_score_err"), [ (f.get("ids_equal"), f.get("common_ids"), round(f.get("box_err",0),2)) for f in c.get("frames",[])])
PY
grep -n "roi_align" "$OUT/launches_720p30_summary.txt" | head
