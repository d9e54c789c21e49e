#!/bin/bash
# A/B of tail-kernel builds and thread counts: c3 short bench per variant
OUT=gpurun_out; mkdir -p $OUT
for v in "" inl; do
 for nt in 1024 768 512; do
  PYCONTACT_B200_LIB=$( [ -n "$v" ] && echo $PWD/pycontact_b200/libpycontact_b200_$v.so ) PC_TAIL_NT=$nt timeout 600 python bench.py --frames 2960 --steps 3 --warmup 3 --no-cpu > $OUT/ab_${v}_${nt}.log 2>&1
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/ab_${v}_${nt}.log").read().strip().splitlines()[-1])
    print("variant '$v' nt $nt: %.0f frames/s, %.3f ms/step, frac %.3f, seq %.3f ms/batch, checked %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["dominant_kernel"]["scan_sequence_ms_per_batch"], d["checked"]["ok"]))
except Exception as e:
    print("variant '$v' nt $nt failed", e); print(open("$OUT/ab_${v}_${nt}.log").read()[-800:])
PY
 done
done
