#!/bin/bash
# Round-end evidence in one call: tests, bench lines (c2 with CPU baseline, c3, c4, reference arm), launch lists of c2 / c3,
# SASA timings (each ncu pass only after the plain command exited 0).
TAG=${1:-final}; OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.sm --format=csv > $OUT/${TAG}_gpu.txt 2>&1; nproc >> $OUT/${TAG}_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest.log
timeout 900 python bench.py > $OUT/${TAG}_bench_c2.log 2>&1; echo "bench c2 rc=$?"; tail -1 $OUT/${TAG}_bench_c2.log | cut -c1-200
timeout 900 python bench.py --workload c3 > $OUT/${TAG}_bench_c3.log 2>&1; echo "bench c3 rc=$?"; tail -1 $OUT/${TAG}_bench_c3.log | cut -c1-200
timeout 900 python bench.py --workload c4 --steps 2 --no-cpu > $OUT/${TAG}_bench_c4.log 2>&1; echo "bench c4 rc=$?"; tail -1 $OUT/${TAG}_bench_c4.log | cut -c1-200
timeout 900 python bench.py --impl reference > $OUT/${TAG}_bench_ref.log 2>&1; echo "bench ref rc=$?"; tail -1 $OUT/${TAG}_bench_ref.log | cut -c1-200
timeout 300 python tools/gpu_time_sasa.py 1230 2000 > $OUT/${TAG}_sasa.log 2>&1; timeout 300 python tools/gpu_time_sasa.py 100000 32 >> $OUT/${TAG}_sasa.log 2>&1; echo "sasa rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
timeout 600 $CMD > $OUT/${TAG}_c2_plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_c2_launches.csv $CMD > $OUT/${TAG}_c2_ncu1.log 2>&1; echo "c2 launch list rc=$?"
CMD3="python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu"
timeout 600 $CMD3 > $OUT/${TAG}_c3_plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/${TAG}_c3_launches.csv $CMD3 > $OUT/${TAG}_c3_ncu1.log 2>&1; echo "c3 launch list rc=$?"
