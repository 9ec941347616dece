#!/bin/bash
O=gpurun_out/ap; mkdir -p $O
timeout 300 python bench.py --no-cpu --no-product --steps 10 --warmup 5 > $O/bench_new.json 2> $O/bench_new.err; python -c "
import json; d=json.load(open('$O/bench_new.json')); print('new', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['config']['rounds_per_tile'])"
HB_DEBUG=512 timeout 300 python bench.py --no-cpu --no-product --steps 10 --warmup 5 > $O/bench_old.json 2> $O/bench_old.err; python -c "
import json; d=json.load(open('$O/bench_old.json')); print('old', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['config']['rounds_per_tile'])"
HB_PHASES=1 timeout 300 python bench.py --no-cpu --no-product --steps 2 --warmup 3 > $O/bench_phases.json 2> $O/bench_phases.err; tail -3 $O/bench_phases.err
HB_DEBUG=512 HB_PHASES=1 timeout 300 python bench.py --no-cpu --no-product --steps 2 --warmup 3 > $O/bench_phases_old.json 2> $O/bench_phases_old.err; tail -3 $O/bench_phases_old.err
