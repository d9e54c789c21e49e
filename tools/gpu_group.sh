#!/bin/bash
# group kernel: its tests, the large-path suite, c4 / c4s bench lines with and without it, launch list
OUT=gpurun_out; mkdir -p $OUT
T=${1:-grp}
timeout 900 python -m pytest tests/test_gpu_large.py -m gpu -q --timeout 600 -k "group" > $OUT/${T}_pytest_group.log 2>&1; echo "group tests rc=$?"; tail -25 $OUT/${T}_pytest_group.log | cut -c1-250
timeout 1200 python -m pytest tests/test_gpu_large.py -m gpu -q --timeout 600 -k "not group" > $OUT/${T}_pytest_large.log 2>&1; echo "large rc=$?"; tail -5 $OUT/${T}_pytest_large.log | cut -c1-250
for g in ${GROUPS_TO_RUN:-1 0}; do
for w in c4 c4s; do
  PC_GROUP=$g timeout 900 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu > $OUT/${T}_${w}_g$g.log 2> $OUT/${T}_${w}_g$g.err; echo "$w group=$g rc=$?"; tail -c 600 $OUT/${T}_${w}_g$g.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$OUT/${T}_${w}_g$g.log").read().strip().splitlines() if l.startswith("{")][-1])
    k=d["roofline"]["dominant_kernel"]; f=d["roofline"]["fp32"]
    print("$w g=$g: %.0f frames/s, %.3f ms/step (%d frames), frac %.4f, seq %.3f ms/batch, e2e %.0f, checked %s, fp32 frac_alg %s, launches %s, per_step %s" % (d["value"], d["ms_per_step"], d["config"]["frames_per_step"], d["roofline"]["frac"], k["scan_sequence_ms_per_batch"], d["e2e"]["value"], d["checked"]["ok"], f["frac_alg"], d["gpu_launches"], d["per_step"]))
except Exception as e:
    print("$w failed", e); print(open("$OUT/${T}_${w}_g$g.log").read()[-600:])
PY
done
done
CMD="python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu --quick"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/${T}_c4_launches.csv $CMD > $OUT/${T}_c4_ncu1.log 2>&1
python tools/launch_summary.py $OUT/${T}_c4_launches.csv | head -24
