#!/bin/bash
# Full validation: smoke, all GPU tests, bench c2 (with CPU baseline), bench c3, reference arm.
TAG=${1:-full}; OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/${TAG}_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/${TAG}_pytest.log
timeout 900 python bench.py > $OUT/${TAG}_bench_c2.log 2>&1; echo "bench c2 rc=$?"; tail -1 $OUT/${TAG}_bench_c2.log | cut -c1-400
timeout 900 python bench.py --workload c3 --no-cpu > $OUT/${TAG}_bench_c3.log 2>&1; echo "bench c3 rc=$?"; tail -1 $OUT/${TAG}_bench_c3.log | cut -c1-400
