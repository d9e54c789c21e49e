#!/bin/bash
# One gpurun call: smoke, GPU tests, bench, and (only if the bench exited 0) an ncu launch list.
# usage: tools/gpu_round.sh [tag] [pytest -k expr]
TAG=${1:-r1}
KEXPR=${2:-}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.sm --format=csv > $OUT/${TAG}_gpu.txt 2>&1
nproc >> $OUT/${TAG}_gpu.txt; free -g >> $OUT/${TAG}_gpu.txt
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
tail -5 $OUT/${TAG}_smoke.log
echo "== pytest gpu"
if [ -n "$KEXPR" ]; then
  timeout 1500 python -m pytest tests -m gpu -q -k "$KEXPR" > $OUT/${TAG}_pytest.log 2>&1
else
  timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1
fi
echo "pytest rc=$?"; tail -40 $OUT/${TAG}_pytest.log
echo "== bench small"; timeout 300 python bench.py --workload small --steps 3 --warmup 3 --no-cpu > $OUT/${TAG}_bench_small.log 2>&1; echo "rc=$?"; tail -3 $OUT/${TAG}_bench_small.log
echo "== bench c2"; timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/${TAG}_bench_c2.log 2>&1; BRC=$?; echo "rc=$BRC"; tail -3 $OUT/${TAG}_bench_c2.log
