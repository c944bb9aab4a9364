#!/bin/bash
# Round-2 GPU call J (1 GPU): correlation kernel after the B-word revert (per-copy mbarriers + late dependency wait kept);
# ncu --set full of the roofline kernel and of the four persistent hi-res kernels.
set +e
OUT=gpurun_out/r02j
mkdir -p "$OUT"
timeout 600 python -m pytest tests/test_more_gpu.py tests/test_fp16_e2e_gpu.py -q -m gpu -k "xcorr or planar or fp16" > "$OUT/pytest_xcorr.txt" 2>&1
echo "rc=$?" >> "$OUT/pytest_xcorr.txt"
timeout 600 python tools/xcorr_lab.py --out "$OUT/xcorr_lab.json" > "$OUT/xcorr_lab.log" 2>&1
echo "rc=$?" >> "$OUT/xcorr_lab.log"
B="--steps 100 --warmup 10 --experimental off --no-cpu-baseline"
timeout 300 python bench.py $B > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:xcorr_planar_kernel -s 2 -c 1 -f -o "$OUT/xcorr_planar" \
    python tools/run_frames.py --frames 4 --eager > "$OUT/ncu_xcorr.log" 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:persist_kernel -s 4 -c 4 -f -o "$OUT/hires_persist" \
    python tools/run_frames.py --frames 3 --eager > "$OUT/ncu_hires.log" 2>&1
for r in xcorr_planar hires_persist; do
  ncu -i "$OUT/$r.ncu-rep" --page raw --csv > "$OUT/${r}_raw.csv" 2> /dev/null
done
tail -n 4 "$OUT/pytest_xcorr.txt"
grep "'mma_mode': 1" "$OUT/xcorr_lab.log" | cut -c1-200
python - "$OUT/xcorr_lab.json" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
for t in d["trace"][:1]:
    print("TRACE n=%d C=%d cg=%d ctas=%d per_sm=%s" % (t["n"],t["C"],t["channel_group"],t["ctas"],t["ctas_per_sm"]))
    for k,v in t["timeline_ns_since_first_cta_start (MMA warps; copy warp where said)"].items(): print("   T %-32s %s" % (k,v))
    for k,v in t["phase_cycles_per_warp (clock64)"].items(): print("   C %-32s %s" % (k,v))
PY
python - "$OUT/bench_default.json" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print("value", d["value"], "e2e", d["e2e"]["value"], "per_frame", d["e2e"]["per_frame_call"]["value"], "static", d["stage_ms"]["static_graph"], "xcorr us", d["roofline"]["us_per_launch"], "frac", d["roofline"]["frac"], "traffic", d["roofline"]["traffic"], d["e2e"]["clip_error"])
PY
ls -la "$OUT"
