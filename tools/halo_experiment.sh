mkdir -p gpurun_out
for pr in 0 256 1536 1792; do
  SMOT_LAYERS="level3,rpn conv,level4" SMOT_TC_HALO=16 SMOT_TC_PROBE=$pr timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv --log-file gpurun_out/probe_$pr.csv python tools/bench_conv.py > gpurun_out/probe_$pr.log 2>&1
done
SMOT_LAYERS="level3,rpn conv,level4" timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv --log-file gpurun_out/probe_old.csv python tools/bench_conv.py > gpurun_out/probe_old.log 2>&1
ls -la gpurun_out
