#!/bin/bash
O=gpurun_out/ai; mkdir -p $O
b() { name=$1; w=$2; shift; shift; env "$@" timeout 600 python bench.py --no-cpu --no-product --steps 20 --warmup $w > $O/$name.json 2> $O/$name.err; python - $name <<'PY'
import json,sys
try:
    j=json.loads(open('gpurun_out/ai/%s.json'%sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1],'ms=%.2f kms=%.2f rpt=%.3f changed=%.0f'%(j['ms_per_step'], j['roofline']['kernel_ms'], j['config']['rounds_per_tile'], j['config']['changed_snps_per_sweep']))
except Exception as e: print(sys.argv[1],'ERR',e)
PY
}
b cond1 5 HB_X=0
b cond0 5 HB_COND=0
b fold64_cond1 10 HB_BENCH_FOLD_SCALE=64
b fold64_cond0 10 HB_BENCH_FOLD_SCALE=64 HB_COND=0
HB_PHASES=1 timeout 600 python bench.py --no-cpu --no-product --steps 6 --warmup 4 > $O/ph.json 2> $O/ph.err
grep "hb phases" $O/ph.err | tail -1 | cut -c1-400
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_scalar_modes_gpu.py -m gpu -q > $O/pytest_1.log 2>&1; echo "rc=$?" >> $O/pytest_1.log; tail -5 $O/pytest_1.log
