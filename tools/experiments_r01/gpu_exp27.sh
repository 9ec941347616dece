#!/bin/bash
# experimental: scalar workers in clusters of 2, hand-over through distributed shared memory (HB_CLUSTER=1)
mkdir -p gpurun_out
HB_CLUSTER=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3 > gpurun_out/e27_pytest.txt
cat gpurun_out/e27_pytest.txt
run() { local name=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu --steps 10 --warmup 5 > gpurun_out/e27_$name.json 2> gpurun_out/e27_$name.err
  echo "== $name"
  python -c "
import json; d=json.loads(open('gpurun_out/e27_$name.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['config']['scalar_rounds_per_sweep'], d['config']['changed_snps_per_sweep'])"
  grep -h "\[hb\]" gpurun_out/e27_$name.err | head -3
}
run cluster HB_CLUSTER=1
run plain HB_CLUSTER=0
