#!/bin/bash
# One gpurun call for the SASA / within kernels: parity tests, then timings.
TAG=${1:-sasa}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_sasa.py -q -x > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/${TAG}_pytest.log
timeout 300 python tools/gpu_time_sasa.py 1230 2000 > $OUT/${TAG}_time.log 2>&1; echo "time rc=$?"; tail -4 $OUT/${TAG}_time.log
timeout 300 python tools/gpu_time_sasa.py 100000 32 >> $OUT/${TAG}_time.log 2>&1; echo "time rc=$?"; tail -3 $OUT/${TAG}_time.log
