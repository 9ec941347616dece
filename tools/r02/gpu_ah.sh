#!/bin/bash
O=gpurun_out/ah; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 3 -c 1 -f -o $O/prof_cond python bench.py --no-cpu --no-product --m 100000 --steps 1 --warmup 3 > $O/ncu.log 2>&1
ls -la $O | tail -3
