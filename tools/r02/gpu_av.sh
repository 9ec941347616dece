#!/bin/bash
O=gpurun_out/av; mkdir -p $O
timeout 600 python -m pytest tests/test_sbayes.py tests/test_reference_pin.py tests/test_dropin_rcpp.py -m gpu -q -x -k "sbayes" > $O/pytest_1.log 2>&1; echo "rc=$?" >> $O/pytest_1.log; tail -5 $O/pytest_1.log
timeout 600 python bench.py --config c4 --m 60000 > $O/bench_c4.json 2> $O/bench_c4.err; python -c "
import json; d=json.load(open('$O/bench_c4.json')); print('c4', d['ms_per_step'], d['value'], d['roofline']['frac'], d['config']['rounds_per_tile'], d['config']['columns_per_sweep'])" || tail -5 $O/bench_c4.err
