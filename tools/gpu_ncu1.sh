#!/bin/bash
mkdir -p gpurun_out
HB_DEBUG=32 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 3 -c 1 -f -o gpurun_out/prof_stream32 python bench.py --no-cpu --m 200000 --steps 1 --warmup 3 > gpurun_out/ncu1.log 2>&1
tail -5 gpurun_out/ncu1.log
ls -la gpurun_out/
