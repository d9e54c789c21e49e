#!/bin/bash
# bench lines of one workload with several library builds: tools/gpu_variants.sh TAG WORKLOAD lib1.so lib2.so ...
TAG=$1; W=$2; shift 2
OUT=gpurun_out; mkdir -p $OUT
for VAR in "$@"; do
  V=$(basename $VAR .so)
  PYCONTACT_B200_LIB=$VAR timeout 300 python bench.py --workload $W --no-cpu --steps 5 --warmup 3 > $OUT/${TAG}_${W}_$V.log 2> $OUT/${TAG}_${W}_$V.err
  python - $OUT/${TAG}_${W}_$V.log <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
    k=d["roofline"]["dominant_kernel"]
    print("%s: %.0f frames/s, %.3f ms/step, kernel %.4f ms/launch (%d frames), checked %s" % (sys.argv[1], d["value"], d["ms_per_step"], k["ms_per_launch"], d["config"]["batch_frames"], (d.get("checked") or {}).get("ok")))
except Exception as e:
    print(sys.argv[1], "failed", e); print(open(sys.argv[1]).read()[-800:])
PY
done
