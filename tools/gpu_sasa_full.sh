#!/bin/bash
# ncu --set full of pc_sasa_atoms_kernel (after the plain run exited 0)
TAG=${1:-sasaf}; N=${2:-1230}; F=${3:-2000}
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python tools/gpu_time_sasa.py $N $F > $OUT/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/${TAG}_plain.log; exit 1; }
head -1 $OUT/${TAG}_plain.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pc_sasa_atoms -s 1 -c 1 -o $OUT/${TAG}_prof -f python tools/gpu_time_sasa.py $N $F > $OUT/${TAG}_ncu2.log 2>&1; echo "full rc=$?"
