#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for wa in 1 2; do for wb in 1 2 4 8; do
PC_WAVES_A=$wa PC_WAVES_B=$wb timeout 600 python bench.py --frames 2960 --steps 3 --warmup 3 --no-cpu > $OUT/wv_${wa}_${wb}.log 2>&1
python - <<PY
import json
try:
    d=json.loads(open("$OUT/wv_${wa}_${wb}.log").read().strip().splitlines()[-1])
    k=d["roofline"]["dominant_kernel"]
    print("wavesA $wa wavesB $wb: %.0f frames/s, frac %.3f, seq %.3f ms/batch, binB %.3f ms (%.0f GB/s), checked %s" % (d["value"], d["roofline"]["frac"], k["scan_sequence_ms_per_batch"], k["ms_per_launch"], k["achieved"], d["checked"]["ok"]))
except Exception as e:
    print("failed", e); print(open("$OUT/wv_${wa}_${wb}.log").read()[-800:])
PY
done; done
