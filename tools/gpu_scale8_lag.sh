#!/bin/bash
mkdir -p gpurun_out
for LAG in 8 6; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29700+LAG)) bench.py --gpus 8 --steps 10 --warmup 5 --lag $LAG > gpurun_out/scale_n8_lag$LAG.json 2> gpurun_out/scale_n8_lag$LAG.err
python -c "
import json
d=json.loads(open('gpurun_out/scale_n8_lag$LAG.json').read().strip().splitlines()[-1]); print('lag',$LAG,'N',d['n_gpus'],'value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'])"
done
