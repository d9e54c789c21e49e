#!/bin/bash
# usage: tools/gpu_k.sh TAG "pytest -k expr"
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x -k "$2" > $OUT/$1_pytest.log 2>&1; echo "pytest rc=$?"; tail -30 $OUT/$1_pytest.log
