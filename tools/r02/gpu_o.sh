#!/bin/bash
mkdir -p gpurun_out/o
O=gpurun_out/o
run() { # name, env..., -- args
  local name=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python bench.py --no-cpu "$@" > $O/$name.json 2> $O/$name.err
  python - "$name" <<'PY'
import json,sys
nm=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/o/%s.json'%nm).read().strip().splitlines()[-1])
    print(nm, 'ms/step %.3f kernel_ms %.3f rounds %.1f changed %.0f' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['config']['scalar_rounds_per_sweep'], d['config']['changed_snps_per_sweep']))
except Exception as e:
    print(nm, 'FAILED', e)
PY
  grep -h "error\|Error" $O/$name.err | tail -2
}
run ring HB_RING=1 -- --steps 10 --warmup 5 --lag 5
run ring_x HB_RING=1 HB_XEVICT=1 -- --steps 10 --warmup 5 --lag 5
run ring_x_d8 HB_RING=1 HB_XEVICT=1 -- --steps 10 --warmup 5 --lag 8
run ring_x_d7 HB_RING=1 HB_XEVICT=1 -- --steps 10 --warmup 5 --lag 7
run ring_x_d6 HB_RING=1 HB_XEVICT=1 -- --steps 10 --warmup 5 --lag 6
run ring_x_d8_limbs HB_RING=1 HB_XEVICT=1 HB_LIMBS=1 -- --steps 10 --warmup 5 --lag 8
run ring_x_d8_ng12 HB_RING=1 HB_XEVICT=1 HB_NG=12 -- --steps 10 --warmup 5 --lag 8
run ring_x_d8_ns3 HB_RING=1 HB_XEVICT=1 HB_NS=3 -- --steps 10 --warmup 5 --lag 8
run ring_x_d8_steady HB_RING=1 HB_XEVICT=1 -- --steps 20 --warmup 300 --lag 8
run serial_x_steady HB_XEVICT=1 -- --steps 20 --warmup 300
