#!/bin/bash
# c3: ncu launch list (short run) + full capture of the tail kernel and of the bin kernel; batch-size sweep
OUT=gpurun_out; mkdir -p $OUT
ARGS="--frames 592 --steps 1 --warmup 1 --quick --no-cpu"
timeout 600 python bench.py $ARGS > $OUT/r2b_plain.log 2>&1 || { echo plain failed; tail -5 $OUT/r2b_plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $OUT/r2b_c3_launches.csv python bench.py $ARGS > $OUT/r2b_ncu1.log 2>&1
echo "launch list rc=$?"
python tools/launch_summary.py $OUT/r2b_c3_launches.csv | tee $OUT/r2b_c3_launch_summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pc_large_tail -s 2 -c 1 -o $OUT/r2b_tail_prof -f python bench.py $ARGS > $OUT/r2b_ncu2.log 2>&1
echo "tail full rc=$?"
ncu -i $OUT/r2b_tail_prof.ncu-rep --page details > $OUT/r2b_tail_ncu_full.txt 2>&1
for b in 74 148 296; do
  for sl in 2 3 4; do
    timeout 600 python bench.py --frames 2960 --steps 3 --warmup 3 --no-cpu --batch $b --slots $sl > $OUT/r2b_sweep_b${b}_s${sl}.log 2>&1
    python - <<PY
import json
try:
    d=json.loads(open("$OUT/r2b_sweep_b${b}_s${sl}.log").read().strip().splitlines()[-1])
    print("batch $b slots $sl: %.0f frames/s, %.3f ms/step, frac %.3f, seq %.3f ms/batch" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["dominant_kernel"]["scan_sequence_ms_per_batch"]))
except Exception as e:
    print("batch $b slots $sl failed", e)
PY
  done
done
