#!/bin/bash
# both bench arms once more with the final library
O=gpurun_out/final5; mkdir -p $O
timeout 300 python bench.py --impl reference --steps 3 > $O/bench_reference.json 2> $O/bench_reference.err; python -c "
import json; d=json.load(open('$O/bench_reference.json')); print('ref', d['value'], d['cpu_baseline']['kind'])"
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; python -c "
import json; d=json.load(open('$O/bench_n1.json')); print('ours', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['x_load_s'], d['cpu_baseline']['value'], d['clocks'])"
