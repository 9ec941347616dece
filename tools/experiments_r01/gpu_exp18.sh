#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/exp18.txt
for FS in 8 64; do
HB_BENCH_FOLD_SCALE=$FS HB_PHASES=1 python bench.py --no-cpu --steps 6 --warmup 14 2> gpurun_out/tmp.err | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('fold x$FS ms_per_step',round(d['ms_per_step'],3),'value',round(d['value']/1e6,2),'rounds',d['config']['scalar_rounds_per_sweep'],'changed',d['config']['changed_snps_per_sweep'])" >> gpurun_out/exp18.txt
grep -o "re-speculated before the chain: [0-9]*, rounds [0-9]*" gpurun_out/tmp.err | tail -3 >> gpurun_out/exp18.txt
done
cat gpurun_out/exp18.txt
