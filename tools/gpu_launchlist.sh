#!/bin/bash
# usage: tools/gpu_launchlist.sh TAG "bench args"   (plain run, then the ncu launch list of the same command)
TAG=$1; ARGS=$2
OUT=gpurun_out; mkdir -p $OUT
CMD="python bench.py $ARGS"
timeout 900 $CMD > $OUT/${TAG}_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu1.log 2>&1
echo "rc=$?"; tail -c 3000 $OUT/${TAG}_plain.log
