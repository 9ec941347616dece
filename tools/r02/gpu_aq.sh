#!/bin/bash
O=gpurun_out/aq; mkdir -p $O
for ng in 8 10 12 16; do
HB_NG=$ng timeout 300 python bench.py --no-cpu --no-product --steps 10 --warmup 5 > $O/bench_ng$ng.json 2> $O/bench_ng$ng.err; python -c "
import json; d=json.load(open('$O/bench_ng$ng.json')); print('ng$ng', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['config']['rounds_per_tile'], d['config']['layout'])"
done
HB_NG=12 HB_DEBUG=512 timeout 300 python bench.py --no-cpu --no-product --steps 10 --warmup 5 > $O/bench_ng12_old.json 2> $O/bench_ng12_old.err; python -c "
import json; d=json.load(open('$O/bench_ng12_old.json')); print('ng12 old', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
