#!/bin/bash
TAG=${1:-final}; OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest.log
timeout 900 python bench.py > $OUT/${TAG}_bench_c2.log 2>&1; echo "bench c2 rc=$?"; tail -1 $OUT/${TAG}_bench_c2.log | cut -c1-200
CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
timeout 600 $CMD > $OUT/${TAG}_c2_plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_c2_launches.csv $CMD > $OUT/${TAG}_c2_ncu1.log 2>&1; echo "c2 launch list rc=$?"
PYCONTACT_B200_LIB=pycontact_b200/libpycontact_b200_prof.so timeout 250 python tools/phase_clocks.py c2 2048 > $OUT/${TAG}_phase.log 2>&1; echo "phase rc=$?"
