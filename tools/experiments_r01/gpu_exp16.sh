#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/exp16.txt; : > $out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 >> $out
HB_PHASES=1 timeout 300 python bench.py --no-cpu --steps 5 --warmup 3 2> gpurun_out/tmp.err | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('ms_per_step',d['ms_per_step'],'value',d['value'],'rounds',d['config'].get('scalar_rounds_per_sweep'))" >> $out
grep "re-spec" gpurun_out/tmp.err | tail -2 >> $out
cat $out
