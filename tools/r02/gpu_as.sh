#!/bin/bash
O=gpurun_out/as; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "loader or demo_data or ragged" > $O/pytest_1.log 2>&1; echo "rc=$?" >> $O/pytest_1.log; tail -8 $O/pytest_1.log
timeout 600 python bench.py --no-cpu --steps 5 --warmup 3 > $O/bench_e2e.json 2> $O/bench_e2e.err; python -c "
import json; d=json.load(open('$O/bench_e2e.json')); print(d['ms_per_step'], d['e2e'])" || tail -5 $O/bench_e2e.err
