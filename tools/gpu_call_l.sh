#!/bin/bash
# Round-2 GPU call L (1 GPU): flat form of the planar correlation kernel (one CTA per SM, 4-plane units).
set +e
OUT=gpurun_out/r02l
mkdir -p "$OUT"
timeout 600 python -m pytest tests/test_more_gpu.py tests/test_fp16_e2e_gpu.py tests/test_e2e_gpu.py -q -m gpu > "$OUT/pytest_sel.txt" 2>&1
echo "rc=$?" >> "$OUT/pytest_sel.txt"
timeout 600 python tools/xcorr_lab.py --out "$OUT/xcorr_lab.json" > "$OUT/xcorr_lab.log" 2>&1
echo "rc=$?" >> "$OUT/xcorr_lab.log"
B="--steps 100 --warmup 10 --experimental off --no-cpu-baseline"
timeout 300 python bench.py $B > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"
SMOT_XCORR_FLAT=0 timeout 300 python bench.py $B > "$OUT/bench_noflat.json" 2> "$OUT/bench_noflat.err"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:xcorr_flat_kernel -s 2 -c 1 -f -o "$OUT/xcorr_flat" \
    python tools/run_frames.py --frames 4 --eager > "$OUT/ncu_xcorr.log" 2>&1
ncu -i "$OUT/xcorr_flat.ncu-rep" --page raw --csv > "$OUT/xcorr_flat_raw.csv" 2> /dev/null
tail -n 4 "$OUT/pytest_sel.txt"
grep "'mma_mode': 1" "$OUT/xcorr_lab.log" | cut -c1-200
for f in "$OUT"/bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("value", d["value"], "e2e", d["e2e"]["value"], "per_frame", d["e2e"]["per_frame_call"]["value"], "static", d["stage_ms"]["static_graph"], d["roofline"]["kernel"], "xcorr us", d["roofline"]["us_per_launch"], "frac", d["roofline"]["frac"], d["e2e"]["clip_error"])
except Exception as e:
    print("ERR", e)
PY
done
