#!/bin/bash
# walk kernel: tests of the large path, c3 short bench in the three tail modes, launch list + full capture of the walk kernel
OUT=gpurun_out; mkdir -p $OUT
T=${1:-r2c}
echo "== pytest large"; timeout 1200 python -m pytest tests/test_gpu_large.py -m gpu -q --timeout 600 -x > $OUT/${T}_pytest_large.log 2>&1; echo "rc=$?"; tail -30 $OUT/${T}_pytest_large.log | cut -c1-300
for m in 1 2; do
  PC_TAIL=$m timeout 600 python bench.py --frames 2960 --steps 3 --warmup 3 --no-cpu > $OUT/${T}_c3_tail$m.log 2>&1
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/${T}_c3_tail$m.log").read().strip().splitlines()[-1])
    print("tail $m: %.0f frames/s, %.3f ms/step, frac %.3f, seq %.3f ms/batch, checked %s, launches %d" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["dominant_kernel"]["scan_sequence_ms_per_batch"], d["checked"]["ok"], d["gpu_launches"]))
except Exception as e:
    print("tail $m failed", e); print(open("$OUT/${T}_c3_tail$m.log").read()[-1500:])
PY
done
ARGS="--frames 592 --steps 1 --warmup 1 --quick --no-cpu"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $OUT/${T}_c3_launches.csv python bench.py $ARGS > $OUT/${T}_ncu1.log 2>&1
python tools/launch_summary.py $OUT/${T}_c3_launches.csv | tee $OUT/${T}_c3_launch_summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pc_large_tail -s 2 -c 1 -o $OUT/${T}_tail_prof -f python bench.py $ARGS > $OUT/${T}_ncu2.log 2>&1
echo "walk full rc=$?"
ncu -i $OUT/${T}_tail_prof.ncu-rep --page details > $OUT/${T}_tail_ncu_full.txt 2>&1
