#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/exp6.txt; : > $out
run() { echo "== $* $ARGS" >> $out; env "$@" HB_PHASES=1 timeout 300 python bench.py --no-cpu --steps 5 --warmup 3 ${ARGS} 2> gpurun_out/tmp.err | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('ms_per_step',d['ms_per_step'],'value',d['value'],'rounds',d['config'].get('scalar_rounds_per_sweep'),'changed',d['config']['changed_snps_per_sweep'],'layout',d['config']['layout'])" >> $out 2>&1; tail -2 gpurun_out/tmp.err >> $out; }
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 >> $out
ARGS="" run A=1
ARGS="" run HB_NG=8
ARGS="" run HB_NG=2
ARGS="--lag 8" run HB_NG=8
ARGS="" run HB_DEBUG=32
cat $out
