#!/bin/bash
# One experiment on the small-frame kernel: plan timings, phase clocks, a parity subset.
# usage: tools/gpu_exp.sh TAG [plans] [pytest -k expr]
TAG=${1:-exp}; PLANS=${2:-0,1}; KEXPR=${3:-"golden or fixture or small"}
OUT=gpurun_out; mkdir -p $OUT
[ -f pycontact_b200/libpycontact_b200_base.so ] && { PYCONTACT_B200_LIB=pycontact_b200/libpycontact_b200_base.so timeout 250 python tools/gpu_time_small.py c2 2048 0 2>&1 | tail -1 | sed "s/^/BASE /"; }
timeout 250 python tools/gpu_time_small.py c2 2048 $PLANS > $OUT/${TAG}_time.log 2>&1; echo "time rc=$?"; tail -5 $OUT/${TAG}_time.log
PYCONTACT_B200_LIB=pycontact_b200/libpycontact_b200_prof.so timeout 250 python tools/phase_clocks.py c2 2048 > $OUT/${TAG}_phase.log 2>&1; echo "phase rc=$?"; head -14 $OUT/${TAG}_phase.log
timeout 900 python -m pytest tests -m gpu -q -x -k "$KEXPR" > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/${TAG}_pytest.log
