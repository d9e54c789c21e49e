#!/bin/bash
# A/B of library variants (PYCONTACT_B200_LIB) on the c3 short bench: tools/gpu_ab.sh "" u1 u4
OUT=gpurun_out; mkdir -p $OUT
for v in "$@"; do
  LIB=""; [ -n "$v" ] && LIB=$PWD/pycontact_b200/libpycontact_b200_$v.so
  for rep in 1 2; do
  PYCONTACT_B200_LIB=$LIB timeout 600 python bench.py --frames 2960 --steps 3 --warmup 3 --no-cpu > $OUT/ab_${v}_$rep.log 2>&1
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/ab_${v}_$rep.log").read().strip().splitlines()[-1])
    print("variant '$v' run $rep: %.0f frames/s, %.3f ms/step, frac %.3f, seq %.3f ms/batch, checked %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["dominant_kernel"]["scan_sequence_ms_per_batch"], d["checked"]["ok"]))
except Exception as e:
    print("variant '$v' failed", e); print(open("$OUT/ab_${v}_$rep.log").read()[-800:])
PY
  done
done
