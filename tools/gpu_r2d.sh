#!/bin/bash
# round 2: fixed cost of a step (strong-scaling shares of the c3 trajectory on ONE GPU) + C5 subset with the round-2 kernels
OUT=gpurun_out; mkdir -p $OUT
for f in 1250 2500 5000; do
  timeout 300 python bench.py --frames $f --no-cpu --steps 5 --warmup 3 > $OUT/ovh_f$f.log 2> $OUT/ovh_f$f.err; echo "frames=$f rc=$?"
  python - <<PY
import json
d=json.loads([l for l in open("$OUT/ovh_f$f.log").read().splitlines() if l.startswith("{")][-1])
print("frames %d: %.3f ms/step, batch %d x %d, launches %d, seq %.3f ms/batch" % ($f, d["ms_per_step"], d["config"]["batch_frames"], d["config"]["batches_per_gpu_per_step"], d["gpu_launches"], d["roofline"]["dominant_kernel"]["scan_sequence_ms_per_batch"]))
PY
done
timeout 600 python tools/sweep_c5.py 10000,100000 5,8,12 > $OUT/c5_r2.md 2> $OUT/c5_r2.err; echo "c5 rc=$?"; cat $OUT/c5_r2.md
