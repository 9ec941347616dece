#!/bin/bash
O=gpurun_out/ao; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_scalar_modes_gpu.py tests/test_reference_pin.py -m gpu -q -x > $O/pytest_1.log 2>&1; echo "rc=$?" >> $O/pytest_1.log; tail -6 $O/pytest_1.log
timeout 300 python bench.py --no-cpu --no-product --steps 10 --warmup 5 > $O/bench_new.json 2> $O/bench_new.err; python -c "
import json; d=json.load(open('$O/bench_new.json')); print('new', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['config']['rounds_per_tile'])"
HB_DEBUG=512 timeout 300 python bench.py --no-cpu --no-product --steps 10 --warmup 5 > $O/bench_old.json 2> $O/bench_old.err; python -c "
import json; d=json.load(open('$O/bench_old.json')); print('old', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['config']['rounds_per_tile'])"
HB_PHASES=1 timeout 300 python bench.py --no-cpu --no-product --steps 3 --warmup 3 > $O/bench_phases.json 2> $O/bench_phases.err; tail -12 $O/bench_phases.err
