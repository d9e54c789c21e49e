#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python tools/ingest_bench.py --system c2 --frames 2000 > $OUT/m_ingest.log 2> $OUT/m_ingest.err; echo "ingest rc=$?"; tail -c 1500 $OUT/m_ingest.log; tail -3 $OUT/m_ingest.err
for w in c4 c2i; do
  extra=""; [ "$w" = "c4" ] && extra="--steps 2 --warmup 3"
  timeout 1200 python bench.py --workload $w $extra > $OUT/m_$w.log 2> $OUT/m_$w.err; echo "$w rc=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$OUT/m_$w.log").read().strip().splitlines() if l.startswith("{")][-1])
    k=d["roofline"]["dominant_kernel"]; f=d["roofline"]["fp32"]
    print("$w: %.0f frames/s, %.3f ms/step, frac %.3f, kernel %.0f GB/s (%.3f), e2e %.0f, checked %s, cpu %s, fp32 frac_alg %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], k["achieved"], k["frac"], d["e2e"]["value"], d["checked"], (d["cpu_baseline"] or {}).get("value"), f["frac_alg"]))
except Exception as e:
    print("$w failed", e); print(open("$OUT/m_$w.log").read()[-600:]); print(open("$OUT/m_$w.err").read()[-1500:])
PY
done
timeout 600 python bench.py --frames 2960 --steps 3 --warmup 3 --no-cpu > $OUT/m_c3s.log 2>&1; python - <<PY
import json
d=json.loads(open("$OUT/m_c3s.log").read().strip().splitlines()[-1]); print("c3 short: %.0f fps frac %.3f batch %d entries: %s" % (d["value"], d["roofline"]["frac"], d["config"]["batch_frames"], d["config"]["results"][-120:]))
PY
