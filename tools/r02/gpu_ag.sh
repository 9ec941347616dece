#!/bin/bash
# 8 GPUs: the path-independence fix -- config 3 and the weak run again (labels must be equal across ranks), plus the new test on one GPU
O=gpurun_out/ag; mkdir -p $O
timeout 900 python -m pytest tests/test_scalar_modes_gpu.py -m gpu -q > $O/pytest_1.log 2>&1; echo "rc=$?" >> $O/pytest_1.log; tail -4 $O/pytest_1.log
PORT=29900
run() { name=$1; shift; envs=$1; shift; env $envs timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 8 "$@" > $O/$name.json 2> $O/$name.err; PORT=$((PORT+1)); python - $name <<'PY'
import json,sys
try:
    j=json.loads(open('gpurun_out/ag/%s.json'%sys.argv[1]).read().strip().splitlines()[-1]); c=j['config']
    print(sys.argv[1],'value=%.1fM ms=%.2f snp/s=%.2fM rpt=%.3f'%(j['value']/1e6,j['ms_per_step'],c['snp_updates_per_s']/1e6,c['rounds_per_tile']), j.get('parity_check'))
except Exception as e: print(sys.argv[1],'ERR',e)
PY
}
run c3 HB_X=0 --config c3 --steps 20 --warmup 10 --no-cpu
run weak8 HB_X=0 --steps 20 --warmup 10 --no-cpu
