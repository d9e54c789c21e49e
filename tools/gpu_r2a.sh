#!/bin/bash
# round 2, first call: smoke, the new large-path tests, the rest of the GPU suite, c3 bench (short + full)
TAG=${1:-r2a}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,memory.used,clocks.max.sm,clocks.sm --format=csv > $OUT/${TAG}_gpu.txt 2>&1
nproc >> $OUT/${TAG}_gpu.txt; free -g >> $OUT/${TAG}_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/${TAG}_smoke.log
echo "== pytest large"; timeout 1200 python -m pytest tests/test_gpu_large.py -m gpu -q --timeout 600 > $OUT/${TAG}_pytest_large.log 2>&1; echo "rc=$?"; tail -60 $OUT/${TAG}_pytest_large.log | cut -c1-300
echo "== pytest rest"; timeout 900 python -m pytest tests -m gpu -q --timeout 600 --ignore tests/test_gpu_large.py > $OUT/${TAG}_pytest_rest.log 2>&1; echo "rc=$?"; tail -15 $OUT/${TAG}_pytest_rest.log | cut -c1-300
echo "== bench c3 short"; timeout 600 python bench.py --frames 1480 --steps 3 --warmup 3 --no-cpu > $OUT/${TAG}_bench_c3_short.log 2>&1; echo "rc=$?"; tail -c 6000 $OUT/${TAG}_bench_c3_short.log
echo "== bench c3 short, multi-kernel tail (PC_NO_TAIL)"; PC_NO_TAIL=1 timeout 600 python bench.py --frames 1480 --steps 3 --warmup 3 --no-cpu > $OUT/${TAG}_bench_c3_short_notail.log 2>&1; echo "rc=$?"; tail -c 1500 $OUT/${TAG}_bench_c3_short_notail.log
echo "== bench c3 full"; timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/${TAG}_bench_c3.log 2>&1; echo "rc=$?"; tail -c 8000 $OUT/${TAG}_bench_c3.log
