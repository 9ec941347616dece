#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/exp11.txt; : > $out
HB_TRACE=gpurun_out/trace4.bin HB_PHASES=1 timeout 300 python bench.py --no-cpu --steps 3 --warmup 3 2>> $out | cut -c1-200 >> $out
python tools/trace_report.py gpurun_out/trace4.bin 4 >> $out 2>&1
HB_TRACE=gpurun_out/trace8.bin timeout 300 python bench.py --no-cpu --steps 3 --warmup 3 --lag 8 2>> $out | cut -c1-200 >> $out
python tools/trace_report.py gpurun_out/trace8.bin 8 >> $out 2>&1
cat $out
