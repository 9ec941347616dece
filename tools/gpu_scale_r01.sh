#!/bin/bash
# scaling pass: bench.py at N = 1, 2, 4, 8 ranks (weak scaling over individuals) + DRAM traffic of one full-size sweep
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
for N in 8 4 2 1; do
  if [ "$NG" -ge "$N" ]; then
    if [ "$N" -eq 1 ]; then
      python bench.py --steps 10 --warmup 5 > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
    else
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 10 --warmup 5 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
    fi
    echo "N=$N rc=$?"; python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/scale_n$N.json').read().strip().splitlines()[-1]); print('N',d['n_gpus'],'value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'])
except Exception as e: print('parse failed',e)"
  fi
done
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_sweep -s 3 -c 1 --csv --log-file gpurun_out/traffic_full.csv python bench.py --no-cpu --steps 1 --warmup 3 > gpurun_out/traffic_full.log 2>&1
tail -4 gpurun_out/traffic_full.csv | cut -c1-300
