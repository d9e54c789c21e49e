#!/bin/bash
# strong scaling of the default bench at 8 and 4 ranks (one call on an 8-GPU box)
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L | head -8
for n in ${NS:-8 4}; do
  S=$(date +%s)
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 > $OUT/n${n}_bench.log 2> $OUT/n${n}_bench.err; echo "n=$n rc=$? $(( $(date +%s) - S )) s"
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$OUT/n${n}_bench.log").read().strip().splitlines() if l.startswith("{")][-1])
    print("n=$n: %.0f frames/s, %.3f ms/step, frac %.3f, scaling %s, frames/gpu %d, e2e %.0f, checked %s, ngpu==1gpu %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["scaling"], d["config"]["frames_per_gpu_per_step"], d["e2e"]["value"], d["checked"]["ok"], d["ngpu_equals_1gpu"]))
except Exception as e:
    print("failed", e); print(open("$OUT/n${n}_bench.log").read()[-1500:]); print(open("$OUT/n${n}_bench.err").read()[-1500:])
PY
done
