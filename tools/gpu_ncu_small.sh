#!/bin/bash
# ncu --set full of the small-frame scan kernel (one launch), after the same command exited 0 without ncu.
TAG=${1:-ncu}; OUT=gpurun_out; mkdir -p $OUT
CMD="python tools/gpu_time_small.py c2 2048 0"
timeout 300 $CMD > $OUT/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/${TAG}_plain.log; exit 1; }
tail -1 $OUT/${TAG}_plain.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pc_scan_small_kernel -s 4 -c 1 -o $OUT/${TAG}_prof -f $CMD > $OUT/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 $OUT/${TAG}_ncu.log
