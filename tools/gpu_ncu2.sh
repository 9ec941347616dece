#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 3 -c 1 -f -o gpurun_out/prof_full_v4 python bench.py --no-cpu --m 200000 --steps 1 --warmup 3 > gpurun_out/ncu4.log 2>&1
tail -3 gpurun_out/ncu4.log | cut -c1-200
