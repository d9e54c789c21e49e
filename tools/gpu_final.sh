#!/bin/bash
# Round-end evidence in one call: tests, bench lines (c2 with CPU baseline, c3, reference arm), launch lists,
# one ncu --set full capture of the dominant c2 kernel (each ncu pass only after the plain command exited 0).
TAG=${1:-final}; OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.sm --format=csv > $OUT/${TAG}_gpu.txt 2>&1; nproc >> $OUT/${TAG}_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest.log
timeout 900 python bench.py > $OUT/${TAG}_bench_c2.log 2>&1; echo "bench c2 rc=$?"; tail -1 $OUT/${TAG}_bench_c2.log | cut -c1-300
timeout 900 python bench.py --workload c3 > $OUT/${TAG}_bench_c3.log 2>&1; echo "bench c3 rc=$?"; tail -1 $OUT/${TAG}_bench_c3.log | cut -c1-300
timeout 900 python bench.py --impl reference > $OUT/${TAG}_bench_ref.log 2>&1; echo "bench ref rc=$?"; tail -1 $OUT/${TAG}_bench_ref.log | cut -c1-300
CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
timeout 600 $CMD > $OUT/${TAG}_c2_plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_c2_launches.csv $CMD > $OUT/${TAG}_c2_ncu1.log 2>&1; echo "c2 launch list rc=$?"
CMD3="python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu"
timeout 600 $CMD3 > $OUT/${TAG}_c3_plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/${TAG}_c3_launches.csv $CMD3 > $OUT/${TAG}_c3_ncu1.log 2>&1; echo "c3 launch list rc=$?"
timeout 600 $CMD > $OUT/${TAG}_c2_plain2.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:pc_scan_small_kernel -s 8 -c 1 -o $OUT/${TAG}_c2_prof -f $CMD > $OUT/${TAG}_c2_ncu2.log 2>&1; echo "c2 full rc=$?"
