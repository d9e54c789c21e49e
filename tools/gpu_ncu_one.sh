#!/bin/bash
# usage: tools/gpu_ncu_one.sh TAG KERNEL_REGEX SKIP "bench args"   (plain run first, then ncu --set full of one launch)
TAG=$1; KRE=$2; SKIP=$3; ARGS=$4
OUT=gpurun_out; mkdir -p $OUT
CMD="python bench.py $ARGS"
timeout 600 $CMD > $OUT/${TAG}_plain.log 2>&1 || { echo plain failed; tail -5 $OUT/${TAG}_plain.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE -s $SKIP -c 1 -o $OUT/${TAG}_prof -f $CMD > $OUT/${TAG}_ncu2.log 2>&1
echo "full rc=$?"; tail -c 600 $OUT/${TAG}_plain.log
