#!/bin/bash
mkdir -p gpurun_out
HB_PHASES=1 timeout 300 python bench.py --no-cpu --steps 3 --warmup 3 2>&1 | grep -v "^{" | tail -4 > gpurun_out/exp14.txt
cat gpurun_out/exp14.txt
