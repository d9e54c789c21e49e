#!/bin/bash
# A/B of library builds on the c3 short bench (2960 frames, two runs each), then the large-path tests with the LAST build
# usage: tools/gpu_c3_ab.sh lib1.so lib2.so ...
OUT=gpurun_out; mkdir -p $OUT
for VAR in "$@"; do
  V=$(basename $VAR .so)
  for rep in 1 2; do
  PYCONTACT_B200_LIB=$VAR timeout 200 python bench.py --frames 2960 --no-cpu --steps 3 --warmup 3 > $OUT/v_$V.log 2> $OUT/v_$V.err
  python - $OUT/v_$V.log $V <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
    k=d["roofline"]["dominant_kernel"]
    print("%-28s %.0f frames/s, %.3f ms/step, frac %.3f, bin %.3f ms, seq %.3f ms, checked %s" % (sys.argv[2], d["value"], d["ms_per_step"], d["roofline"]["frac"], k["ms_per_launch"], k["scan_sequence_ms_per_batch"], (d.get("checked") or {}).get("ok")))
except Exception as e:
    print(sys.argv[2], "failed", e); print(open(sys.argv[1].replace(".log",".err")).read()[-800:])
PY
  done
done
PYCONTACT_B200_LIB=$VAR timeout 600 python -m pytest tests/test_gpu_large.py -x -q 2>&1 | tail -3
