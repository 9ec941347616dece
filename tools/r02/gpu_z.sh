#!/bin/bash
O=gpurun_out/z; mkdir -p $O
HB_PHASES=1 timeout 600 python bench.py --no-cpu --no-product --steps 6 --warmup 4 > $O/ph.json 2> $O/ph.err
grep "hb phases" $O/ph.err | tail -4 | cut -c1-700
