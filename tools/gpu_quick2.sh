#!/bin/bash
# A/B of the two tail-kernel forms: tests, c3 short bench each, kernel times
OUT=gpurun_out; mkdir -p $OUT
T=${1:-q}
timeout 900 python -m pytest tests/test_gpu_large.py -m gpu -q --timeout 600 -x > $OUT/${T}_pytest_large.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/${T}_pytest_large.log | cut -c1-300
for impl in 1 0; do export PC_L2HINT=$impl;
timeout 600 python bench.py --frames 2960 --steps 3 --warmup 3 --no-cpu > $OUT/${T}_c3_$impl.log 2>&1
python - <<PY
import json
try:
    d=json.loads(open("$OUT/${T}_c3_$impl.log").read().strip().splitlines()[-1])
    print("l2hint $impl c3: %.0f frames/s, %.3f ms/step, frac %.3f, seq %.3f ms/batch, checked %s, launches %d" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["dominant_kernel"]["scan_sequence_ms_per_batch"], d["checked"]["ok"], d["gpu_launches"]))
except Exception as e:
    print("bench failed", e); print(open("$OUT/${T}_c3_$impl.log").read()[-1500:])
PY
done
unset PC_L2HINT; ARGS="--frames 592 --steps 1 --warmup 1 --quick --no-cpu"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $OUT/${T}_c3_launches.csv python bench.py $ARGS > $OUT/${T}_ncu1.log 2>&1
python tools/launch_summary.py $OUT/${T}_c3_launches.csv | grep -v "ffma\|synth\|pack_gid" | tee $OUT/${T}_c3_launch_summary.txt
