#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
PC_DEBUG=1 PC_TAIL=2 timeout 600 python bench.py --frames 592 --steps 1 --warmup 1 --quick --no-cpu 2>&1 | grep -v "^{" | head -5
