#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/exp4.txt
HB_DEBUG=32 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 3 -c 1 -f -o gpurun_out/prof_stream32_v3 python bench.py --no-cpu --m 200000 --steps 1 --warmup 3 > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log | cut -c1-300
cat gpurun_out/exp4.txt
