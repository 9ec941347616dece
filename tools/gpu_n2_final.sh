#!/bin/bash
# two GPUs: sharded parity tests and the N=2 bench line of the final kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sharded.py -x -q -m gpu 2>&1 | tail -3 > gpurun_out/n2_pytest.txt
cat gpurun_out/n2_pytest.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 5 > gpurun_out/final2_bench_n2.json 2> gpurun_out/final2_bench_n2.err
python -c "
import json
d=json.loads(open('gpurun_out/final2_bench_n2.json').read().strip().splitlines()[-1]); print('N',d['n_gpus'],'value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'], d['config'].get('scalar_rounds_per_sweep'))"
tail -2 gpurun_out/final2_bench_n2.err
