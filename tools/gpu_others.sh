#!/bin/bash
# the other workloads once: c2, c2i (interdigitated chains), c1 (fixture), c4s, c4; c2 at 2 ranks when two GPUs are there
OUT=gpurun_out; mkdir -p $OUT
for w in c2 c2i c1 c4s c4; do
  extra=""; [ "$w" = "c4" ] && extra="--steps 2 --warmup 3"
  timeout 900 python bench.py --workload $w $extra > $OUT/o_$w.log 2> $OUT/o_$w.err; echo "$w rc=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$OUT/o_$w.log").read().strip().splitlines() if l.startswith("{")][-1])
    k=d["roofline"]["dominant_kernel"]; f=d["roofline"]["fp32"]
    print("$w: %.0f frames/s, %.3f ms/step, frac %.3f, kernel %.0f GB/s (%.3f), path %s, e2e %.0f, checked %s, cpu %s, fp32 frac_alg %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], k["achieved"], k["frac"], d["config"]["kernel_path"][:30], d["e2e"]["value"], d["checked"]["ok"], (d["cpu_baseline"] or {}).get("value"), f["frac_alg"]))
except Exception as e:
    print("$w failed", e); print(open("$OUT/o_$w.log").read()[-600:]); print(open("$OUT/o_$w.err").read()[-1200:])
PY
done
