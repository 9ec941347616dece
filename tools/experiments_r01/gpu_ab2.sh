#!/bin/bash
mkdir -p gpurun_out
cp hibayes_b200/libhibayes_b200.so /tmp/lib_cur.so
: > gpurun_out/ab2.txt
for v in new old; do
 for LAG in 5 6; do
  cp ab/lib_$v.so hibayes_b200/libhibayes_b200.so
  python bench.py --no-cpu --steps 10 --warmup 5 --lag $LAG 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$v lag $LAG ms_per_step',round(d['ms_per_step'],3),'kernel_ms',round(d['roofline']['kernel_ms'],3),'value',round(d['value']/1e6,2))"
 done
done >> gpurun_out/ab2.txt
cp /tmp/lib_cur.so hibayes_b200/libhibayes_b200.so
cat gpurun_out/ab2.txt
