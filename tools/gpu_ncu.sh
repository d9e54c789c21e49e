#!/bin/bash
# usage: tools/gpu_ncu.sh TAG KERNEL_REGEX "bench args"
TAG=$1; KRE=$2; ARGS=$3
OUT=gpurun_out; mkdir -p $OUT
CMD="python bench.py $ARGS"
$CMD > $OUT/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu1.log 2>&1
echo "launch list rc=$?"
$CMD > $OUT/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$KRE -s 1 -c 2 -o $OUT/${TAG}_prof -f $CMD > $OUT/${TAG}_ncu2.log 2>&1
echo "full rc=$?"
tail -2 $OUT/${TAG}_plain.log
