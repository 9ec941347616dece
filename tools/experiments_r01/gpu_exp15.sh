#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/exp15.txt; : > $out
run() { echo "== $* $ARGS" >> $out; env "$@" HB_PHASES=1 HB_TRACE=gpurun_out/tr.bin timeout 300 python bench.py --no-cpu --steps 4 --warmup 3 ${ARGS} 2> gpurun_out/tmp.err | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('ms_per_step',d['ms_per_step'],'value',d['value'],'rounds',d['config'].get('scalar_rounds_per_sweep'),'changed',d['config']['changed_snps_per_sweep'])" >> $out 2>&1; tail -1 gpurun_out/tmp.err >> $out; python tools/trace_report.py gpurun_out/tr.bin $LAG | grep -E "period|S done|hand-over arrived \(2\) ->" >> $out 2>&1; }
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 >> $out
LAG=4 ARGS="" run A=1
LAG=4 ARGS="" run HB_DEBUG=64
LAG=6 ARGS="--lag 6" run A=1
cat $out
