#!/bin/bash
mkdir -p gpurun_out
HB_TRACE=gpurun_out/e23_trace.bin timeout 300 python bench.py --no-cpu --steps 6 --warmup 3 > gpurun_out/e23.json 2> gpurun_out/e23.err
python tools/trace_report.py gpurun_out/e23_trace.bin 5 | tee gpurun_out/e23_report.txt
python -c "
import json; d=json.loads(open('gpurun_out/e23.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
rm -f gpurun_out/e23_trace.bin
