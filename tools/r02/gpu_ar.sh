#!/bin/bash
O=gpurun_out/ar; mkdir -p $O
for lag in 8 10 12 16; do
HB_LAG=$lag timeout 300 python bench.py --no-cpu --no-product --steps 10 --warmup 5 > $O/bench_lag$lag.json 2> $O/bench_lag$lag.err; python -c "
import json; d=json.load(open('$O/bench_lag$lag.json')); print('lag$lag', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['config']['rounds_per_tile'], d['config']['layout']['lag_tiles'], d['config']['layout']['gram_bytes'])" || tail -3 $O/bench_lag$lag.err
done
HB_LAG=12 HB_DEBUG=512 timeout 300 python bench.py --no-cpu --no-product --steps 10 --warmup 5 > $O/bench_lag12_old.json 2> $O/bench_lag12_old.err; python -c "
import json; d=json.load(open('$O/bench_lag12_old.json')); print('lag12 old sums', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
HB_LAG=12 HB_PHASES=1 timeout 300 python bench.py --no-cpu --no-product --steps 2 --warmup 3 > $O/bench_phases.json 2> $O/bench_phases.err; tail -3 $O/bench_phases.err
