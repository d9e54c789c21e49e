#!/bin/bash
# full GPU suite + default bench (c3, 10k frames, CPU legs) + reference arm
OUT=gpurun_out; mkdir -p $OUT
T=${1:-full}
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/${T}_smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > $OUT/${T}_pytest.log 2>&1; echo "rc=$?"; tail -8 $OUT/${T}_pytest.log | cut -c1-300
echo "== bench default"; timeout 1200 python bench.py > $OUT/${T}_bench.log 2>&1; echo "rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("$OUT/${T}_bench.log").read().strip().splitlines()[-1])
    k=d["roofline"]["dominant_kernel"]
    print("c3 full: %.0f frames/s, %.3f ms/step, frac %.3f, seq %.3f ms/batch, binB %.3f ms (%.0f GB/s, share %.2f), e2e %.0f, checked %s, cpu %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], k["scan_sequence_ms_per_batch"], k["ms_per_launch"], k["achieved"], k["share_of_step"], d["e2e"]["value"], d["checked"], d["cpu_baseline"]["value"] if d["cpu_baseline"] else None))
except Exception as e:
    print("failed", e); print(open("$OUT/${T}_bench.log").read()[-1500:])
PY
