#!/bin/bash
# what the driver does at round end on one B200: GPU tests, smoke(), both bench arms
O=gpurun_out/final3; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log; tail -5 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log; tail -4 $O/smoke.log
timeout 600 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; python -c "
import json; d=json.load(open('$O/bench_reference.json')); print('ref', d['value'], d['cpu_baseline']['kind'], d['config']['variants'])"
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; python -c "
import json; d=json.load(open('$O/bench_n1.json')); print('ours', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['x_load_s'], d['cpu_baseline']['value'])"
