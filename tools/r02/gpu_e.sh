#!/bin/bash
mkdir -p gpurun_out/e
O=gpurun_out/e
run() { # name, env..., -- args
  local name=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python bench.py --no-cpu "$@" > $O/$name.json 2> $O/$name.err
  python - "$name" <<'PY'
import json,sys
nm=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/e/%s.json'%nm).read().strip().splitlines()[-1])
    print(nm, 'ms/step %.3f kernel_ms %.3f rounds %.1f changed %.0f' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['config']['scalar_rounds_per_sweep'], d['config']['changed_snps_per_sweep']))
except Exception as e:
    print(nm, 'FAILED', e)
PY
  grep -h "serial CTA\|phases serial\|error\|Error" $O/$name.err | tail -2
}
run abl768 HB_DEBUG=768 -- --steps 5 --warmup 3
run abl768ph HB_DEBUG=768 HB_PHASES=1 -- --steps 5 --warmup 3
run abl768ph_stream HB_DEBUG=770 HB_PHASES=1 -- --steps 5 --warmup 3
run abl768_lag5 HB_DEBUG=768 HB_PHASES=1 -- --steps 5 --warmup 3 --lag 5
