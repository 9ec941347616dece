#!/bin/bash
O=gpurun_out/am; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log; tail -6 $O/pytest_gpu.log
timeout 600 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; tail -c 1500 $O/bench_reference.json
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 3000 $O/bench_n1.json
