#!/bin/bash
# what the driver does at round end, on one box: build check is done at home; here GPU tests, smoke, both bench arms
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/final_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.txt 2>&1
python bench.py --impl reference > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err
python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 15 -c 40 --csv --log-file gpurun_out/final_launches.csv python bench.py --no-cpu --steps 4 --warmup 3 > /dev/null 2>&1
cat gpurun_out/final_pytest.txt gpurun_out/final_smoke.txt
python -c "
import json
for f in ('final_bench_ref','final_bench_n1'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, d.get('value'), d.get('ms_per_step'), d.get('roofline',{}).get('frac'), d.get('e2e'), d.get('cpu_baseline',{}).get('value'), d.get('clocks'))"
