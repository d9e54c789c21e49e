#!/bin/bash
# HEAD check in one call: smoke, full GPU suite, the driver's two bench commands, c4 with a launch list
OUT=gpurun_out; mkdir -p $OUT
T=${1:-head}
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/${T}_smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > $OUT/${T}_pytest.log 2>&1; echo "rc=$?"; tail -8 $OUT/${T}_pytest.log | cut -c1-300
S=$(date +%s); echo "== bench default"; timeout 1200 python bench.py > $OUT/${T}_bench.log 2> $OUT/${T}_bench.err; echo "rc=$? $(( $(date +%s) - S )) s"
S=$(date +%s); echo "== bench reference arm"; timeout 1200 python bench.py --impl reference > $OUT/${T}_ref.log 2> $OUT/${T}_ref.err; echo "rc=$? $(( $(date +%s) - S )) s"; tail -c 1200 $OUT/${T}_ref.log
python - <<PY
import json
try:
    d=json.loads(open("$OUT/${T}_bench.log").read().strip().splitlines()[-1])
    k=d["roofline"]["dominant_kernel"]
    print("c3 full: %.0f frames/s, %.3f ms/step, frac %.3f, seq %.3f ms/batch, binB %.3f ms (%.0f GB/s, share %.2f), e2e %.0f, checked %s, cpu %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], k["scan_sequence_ms_per_batch"], k["ms_per_launch"], k["achieved"], k["share_of_step"], d["e2e"]["value"], d["checked"]["ok"], d["cpu_baseline"]["value"] if d["cpu_baseline"] else None))
except Exception as e:
    print("failed", e); print(open("$OUT/${T}_bench.log").read()[-1500:]); print(open("$OUT/${T}_bench.err").read()[-1500:])
PY
echo "== c4"; CMD="python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu"
timeout 900 $CMD > $OUT/${T}_c4.log 2> $OUT/${T}_c4.err; RC=$?; echo "rc=$RC"; tail -c 1500 $OUT/${T}_c4.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("$OUT/${T}_c4.log").read().strip().splitlines() if l.startswith("{")][-1])
    k=d["roofline"]["dominant_kernel"]; f=d["roofline"]["fp32"]
    print("c4: %.0f frames/s, %.3f ms/step (%d frames), frac %.4f, seq %.3f ms/batch, e2e %.0f, checked %s, fp32 frac_alg %s, results %s" % (d["value"], d["ms_per_step"], d["config"]["frames_per_step"], d["roofline"]["frac"], k["scan_sequence_ms_per_batch"], d["e2e"]["value"], d["checked"]["ok"], f["frac_alg"], d["config"]["results"][:80]))
except Exception as e:
    print("c4 failed", e); print(open("$OUT/${T}_c4.log").read()[-600:])
PY
if [ $RC -eq 0 ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/${T}_c4_launches.csv $CMD --quick > $OUT/${T}_c4_ncu1.log 2>&1
  python tools/launch_summary.py $OUT/${T}_c4_launches.csv | head -30
fi
