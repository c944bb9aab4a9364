#!/bin/bash
# Round-2 GPU call M (1 GPU): A/B of the conv-family rules under today's pipeline + ncu --set full of every tcgen05 conv launch of a frame.
set +e
OUT=gpurun_out/r02m
mkdir -p "$OUT"
B="--steps 100 --warmup 10 --no-cpu-baseline"
timeout 300 python bench.py $B > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"
SMOT_TC_NOSPLIT=1 timeout 300 python bench.py $B > "$OUT/bench_nosplit.json" 2> "$OUT/bench_nosplit.err"
SMOT_TC_MINCTAS=48 timeout 300 python bench.py $B > "$OUT/bench_minctas48.json" 2> "$OUT/bench_minctas48.err"
SMOT_TC_MINCTAS=148 timeout 300 python bench.py $B > "$OUT/bench_minctas148.json" 2> "$OUT/bench_minctas148.err"
SMOT_BODY_BRANCHES=1 timeout 300 python bench.py $B > "$OUT/bench_bodybranches.json" 2> "$OUT/bench_bodybranches.err"
SMOT_PDL=0 timeout 300 python bench.py $B > "$OUT/bench_nopdl.json" 2> "$OUT/bench_nopdl.err"
timeout 600 ncu --set full --clock-control none -k regex:"conv_tc_kernel|splitk_reduce" -s 70 -c 75 -f -o "$OUT/conv_tc" \
    python tools/run_frames.py --frames 3 --eager > "$OUT/ncu_conv_tc.log" 2>&1
ncu -i "$OUT/conv_tc.ncu-rep" --page raw --csv > "$OUT/conv_tc_raw.csv" 2> /dev/null
python tools/ncu_summary.py "$OUT/conv_tc_raw.csv" > "$OUT/conv_tc_summary.csv" 2>&1
rm -f "$OUT/conv_tc.ncu-rep"
for f in "$OUT"/bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("value", d["value"], "e2e", d["e2e"]["value"], "per_frame", d["e2e"]["per_frame_call"]["value"], "static", d["stage_ms"]["static_graph"], d["e2e"]["clip_error"], d["spread"]["value_fps"])
except Exception as e:
    print("ERR", e)
PY
done
head -30 "$OUT/conv_tc_summary.csv" | cut -c1-220
