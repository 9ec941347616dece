#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29731 bench.py --gpus 8 > gpurun_out/final_bench_n8.json 2> gpurun_out/final_bench_n8.err
python -c "
import json
d=json.loads(open('gpurun_out/final_bench_n8.json').read().strip().splitlines()[-1]); print('N',d['n_gpus'],'value',d['value'],'ms',d['ms_per_step'],'kernel_ms',d['roofline']['kernel_ms'],'rounds',d['config']['scalar_rounds_per_sweep'],'e2e',d['e2e']['value'])"
