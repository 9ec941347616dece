#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/exp17.txt
python bench.py --no-cpu --steps 10 --warmup 5 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('ms_per_step',round(d['ms_per_step'],3),'value',round(d['value']/1e6,2),'rounds',d['config']['scalar_rounds_per_sweep'])" >> gpurun_out/exp17.txt
cat gpurun_out/exp17.txt
