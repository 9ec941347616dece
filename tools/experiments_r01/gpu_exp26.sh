#!/bin/bash
mkdir -p gpurun_out
run() { local name=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu --steps 10 --warmup 5 > gpurun_out/e26_$name.json 2> gpurun_out/e26_$name.err
  echo "== $name"
  python -c "
import json; d=json.loads(open('gpurun_out/e26_$name.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['config']['scalar_rounds_per_sweep'])"
}
run d4 HB_DEBUG=0
run d8 HB_DEBUG=256
run d4b HB_DEBUG=0
run d8b HB_DEBUG=256
