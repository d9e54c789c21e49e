#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > $OUT/n2_bench.log 2> $OUT/n2_bench.err; echo "rc=$?"
tail -c 3000 $OUT/n2_bench.err | tail -15
python - <<PY
import json
try:
    d=json.loads([l for l in open("$OUT/n2_bench.log").read().strip().splitlines() if l.startswith("{")][-1])
    print("n2: %.0f frames/s, %.3f ms/step, frac %.3f, scaling %s, frames/gpu %d, e2e %.0f, checked %s, ngpu %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["scaling"], d["config"]["frames_per_gpu_per_step"], d["e2e"]["value"], d["checked"]["ok"], d["ngpu_equals_1gpu"]))
except Exception as e:
    print("failed", e); print(open("$OUT/n2_bench.log").read()[-1500:])
PY
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "two_gpus" 2>&1 | tail -2
