#!/bin/bash
# 8 GPUs, final kernel of round 2: weak (50k rows per GPU), BASELINE configs[2], strong; then the row-gathering margin in the flip-heavy regime
O=gpurun_out/final8; mkdir -p $O
PORT=29800
run() { name=$1; shift; envs=$1; shift; env $envs timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 8 "$@" > $O/$name.json 2> $O/$name.err; PORT=$((PORT+1)); python - $name <<'PY'
import json,sys
try:
    j=json.loads(open('gpurun_out/final8/%s.json'%sys.argv[1]).read().strip().splitlines()[-1]); c=j['config']
    print(sys.argv[1],'value=%.1fM ms=%.2f snp/s=%.2fM rpt=%.3f'%(j['value']/1e6,j['ms_per_step'],c['snp_updates_per_s']/1e6,c['rounds_per_tile']), j.get('parity_check',{}).get('small_sharded_vs_oracle'), j.get('parity_check',{}).get('timed_run_labels_equal_across_ranks'))
except Exception as e: print(sys.argv[1],'ERR',e)
PY
}
run weak8 HB_X=0 --steps 20 --warmup 10 --no-cpu
run c3 HB_X=0 --config c3 --steps 20 --warmup 10 --no-cpu
run strong8 HB_X=0 --scaling strong --steps 20 --warmup 10 --no-cpu
run weak8_near25 HB_NEAR=0.25 --steps 20 --warmup 10 --no-cpu
run weak8_near10 HB_NEAR=0.1 --steps 20 --warmup 10 --no-cpu
run c3_near10 HB_NEAR=0.1 --config c3 --steps 20 --warmup 10 --no-cpu
