#!/bin/bash
# Round-end evidence in one gpurun call (one GPU): smoke, the whole GPU suite, the default bench line with its CPU legs, the
# reference arm, c2 / c1 lines, then (only after the plain commands exited 0) ncu: launch list of the default command and
# --set full of the small-frame kernel on c2.
# usage: tools/gpu_final.sh [tag]
TAG=${1:-fin}
OUT=gpurun_out; mkdir -p $OUT
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/${TAG}_smoke.log
echo "== pytest gpu"; S=$(date +%s); timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$? $(( $(date +%s) - S )) s"; tail -3 $OUT/${TAG}_pytest.log
echo "== bench default (c3)"; S=$(date +%s); timeout 900 python bench.py > $OUT/${TAG}_bench_c3.log 2> $OUT/${TAG}_bench_c3.err; BRC=$?; echo "rc=$BRC $(( $(date +%s) - S )) s"
echo "== reference arm"; S=$(date +%s); timeout 900 python bench.py --impl reference > $OUT/${TAG}_bench_c3_ref.log 2> $OUT/${TAG}_bench_c3_ref.err; echo "rc=$? $(( $(date +%s) - S )) s"; tail -c 600 $OUT/${TAG}_bench_c3_ref.log
for w in c2 c1; do
  echo "== bench $w"; timeout 600 python bench.py --workload $w > $OUT/${TAG}_bench_$w.log 2> $OUT/${TAG}_bench_$w.err; echo "rc=$?"
done
python - $OUT/${TAG}_bench_c3.log $OUT/${TAG}_bench_c2.log $OUT/${TAG}_bench_c1.log <<'PY'
import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])
        print("%s: %.0f frames/s, %.3f ms/step, frac %.3f, e2e %.0f, cpu %s (%s), checked %s" % (f, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], (d.get("cpu_baseline") or {}).get("value"), (d.get("cpu_baseline") or {}).get("kind"), (d.get("checked") or {}).get("ok")))
    except Exception as e:
        print(f, "failed", e); print(open(f).read()[-800:])
PY
if [ $BRC -eq 0 ]; then
  echo "== ncu launch list (default command, short)"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_c3_launches.csv python bench.py --frames 2960 --no-cpu --steps 2 --warmup 3 > $OUT/${TAG}_c3_ncu1.log 2>&1; echo "rc=$?"
  echo "== ncu --set full: small-frame kernel on c2"
  timeout 600 python bench.py --workload c2 --no-cpu --steps 2 --warmup 3 > $OUT/${TAG}_c2_plain.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:pc_scan_small_kernel -s 4 -c 1 -o $OUT/${TAG}_c2_small_prof -f python bench.py --workload c2 --no-cpu --steps 2 --warmup 3 > $OUT/${TAG}_c2_ncu2.log 2>&1; echo "rc=$?"
fi
