#!/bin/bash
# 8 GPUs: the metric shape weak (50k rows per GPU), BASELINE configs[2] (BayesB n=200k rows over 8 GPUs), strong scaling (50k rows over 8)
O=gpurun_out/v; mkdir -p $O
run() { name=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 8 "$@" > $O/$name.json 2> $O/$name.err; echo "$name rc=$?"; tail -c 900 $O/$name.json; PORT=$((PORT+1)); }
PORT=29700
run weak8 --steps 20 --warmup 10 --no-cpu
run c3 --config c3 --steps 20 --warmup 10 --no-cpu
run strong8 --scaling strong --steps 20 --warmup 10 --no-cpu
HB_PHASES=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29720 bench.py --gpus 8 --steps 6 --warmup 24 --no-cpu > $O/weak8_ph.json 2> $O/weak8_ph.err
grep "hb phases" $O/weak8_ph.err | tail -4 | cut -c1-300
