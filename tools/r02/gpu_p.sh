#!/bin/bash
mkdir -p gpurun_out/p
O=gpurun_out/p
free -g > $O/host.txt; nproc >> $O/host.txt; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv >> $O/host.txt; cat $O/host.txt
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1
echo "rc=$?" >> $O/pytest_gpu.log; grep -E "passed|failed|FAILED|rc=|error" $O/pytest_gpu.log | tail -20
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -c 3000 $O/bench.json; tail -5 $O/bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"; tail -c 1200 $O/bench_ref.json
