#!/bin/bash
# what the driver does at round end, on one box (GPU tests, smoke, both bench arms) + the ncu evidence for profiles/
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/final2_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final2_smoke.txt 2>&1
python bench.py --impl reference > gpurun_out/final2_bench_ref.json 2> gpurun_out/final2_bench_ref.err
python bench.py > gpurun_out/final2_bench_n1.json 2> gpurun_out/final2_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 15 -c 40 --csv --log-file gpurun_out/final2_launches.csv python bench.py --no-cpu --steps 4 --warmup 3 > /dev/null 2>&1
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_sweep -s 3 -c 1 --csv --log-file gpurun_out/final2_traffic_full.csv python bench.py --no-cpu --steps 1 --warmup 3 > /dev/null 2>&1
cat gpurun_out/final2_pytest.txt gpurun_out/final2_smoke.txt
python -c "
import json
for f in ('final2_bench_ref','final2_bench_n1'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, d.get('value'), d.get('ms_per_step'), d.get('roofline',{}).get('frac'), d.get('e2e'), d.get('cpu_baseline',{}).get('value'), d.get('clocks'))"
tail -3 gpurun_out/final2_traffic_full.csv | cut -c1-300
timeout 420 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 3 -c 1 -f -o gpurun_out/prof_full_v5 python bench.py --no-cpu --m 200000 --steps 1 --warmup 3 > gpurun_out/ncu5.log 2>&1
tail -2 gpurun_out/ncu5.log | cut -c1-200
ls -la gpurun_out/prof_full_v5.ncu-rep 2>/dev/null
