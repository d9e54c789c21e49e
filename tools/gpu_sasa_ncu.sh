#!/bin/bash
# launch list of the SASA / within kernels (after the plain run has exited 0)
TAG=${1:-sasa}; N=${2:-1230}; F=${3:-2000}
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python tools/gpu_time_sasa.py $N $F > $OUT/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/${TAG}_plain.log; exit 1; }
tail -2 $OUT/${TAG}_plain.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/${TAG}_launches.csv python tools/gpu_time_sasa.py $N $F > $OUT/${TAG}_ncu1.log 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py $OUT/${TAG}_launches.csv | head -20
