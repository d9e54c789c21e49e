#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
T=${1:-pt}
ARGS="--frames 592 --steps 1 --warmup 1 --quick --no-cpu"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pc_large_tail_inl -s 2 -c 1 -o $OUT/${T}_tail_prof -f python bench.py $ARGS > $OUT/${T}_ncu2.log 2>&1
echo "tail full rc=$?"
ncu -i $OUT/${T}_tail_prof.ncu-rep --page details > $OUT/${T}_tail_ncu_full.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pc_large_stream_kernel -s 5 -c 1 -o $OUT/${T}_bin_prof -f python bench.py $ARGS > $OUT/${T}_ncu3.log 2>&1
echo "bin full rc=$?"
ncu -i $OUT/${T}_bin_prof.ncu-rep --page details > $OUT/${T}_bin_ncu_full.txt 2>&1
grep -E "Duration|DRAM Throughput|Grid Size" $OUT/${T}_bin_ncu_full.txt | head -5
