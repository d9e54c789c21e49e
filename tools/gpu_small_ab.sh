#!/bin/bash
# small-frame kernel A/B: parity tests, then c2 / c1 / small bench lines with the default library and a variant
# usage: tools/gpu_small_ab.sh TAG [variant.so ...]
TAG=${1:-ab}; shift; VARS="$@"
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/${TAG}_pytest.log
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
    k=d["roofline"]["dominant_kernel"]
    print("%s: %.0f frames/s, %.3f ms/step, kernel %.4f ms/launch (%d frames), checked %s" % (sys.argv[1], d["value"], d["ms_per_step"], k["ms_per_launch"], d["config"]["batch_frames"], (d.get("checked") or {}).get("ok")))
except Exception as e:
    print(sys.argv[1], "failed", e); print(open(sys.argv[1]).read()[-800:])
PY
}
for w in c2 c1; do
  timeout 300 python bench.py --workload $w --no-cpu --steps 5 --warmup 3 > $OUT/${TAG}_$w.log 2> $OUT/${TAG}_$w.err; show $OUT/${TAG}_$w.log
  for VAR in $VARS; do
    V=$(basename $VAR .so)
    PYCONTACT_B200_LIB=$VAR timeout 300 python bench.py --workload $w --no-cpu --steps 5 --warmup 3 > $OUT/${TAG}_${w}_$V.log 2> $OUT/${TAG}_${w}_$V.err; show $OUT/${TAG}_${w}_$V.log
  done
done
