#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/exp5.txt; : > $out
run() { echo "== $* $ARGS" >> $out; env "$@" HB_PHASES=1 timeout 300 python bench.py --no-cpu --steps 5 --warmup 3 ${ARGS} 2> gpurun_out/tmp.err | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('ms_per_step',d['ms_per_step'],'value',d['value'],'rounds',d['config'].get('scalar_rounds_per_sweep'),'changed',d['config']['changed_snps_per_sweep'],'layout',d['config']['layout'])" >> $out 2>&1; tail -2 gpurun_out/tmp.err >> $out; }
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 >> $out
ARGS="" run HB_DEBUG=32
ARGS="" run A=1
HB_DEBUG=32 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 3 -c 1 -f -o gpurun_out/prof_stream32_v3b python bench.py --no-cpu --m 200000 --steps 1 --warmup 3 > gpurun_out/ncu3.log 2>&1
cat $out
