#!/bin/bash
O=gpurun_out/aa; mkdir -p $O
b() { name=$1; shift; env "$@" timeout 600 python bench.py --no-cpu --no-product --steps 20 --warmup 5 > $O/$name.json 2> $O/$name.err; python - $name <<'PY'
import json,sys
try:
    j=json.loads(open('gpurun_out/aa/%s.json'%sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1],'ms=%.2f kms=%.2f rpt=%.3f'%(j['ms_per_step'], j['roofline']['kernel_ms'], j['config']['rounds_per_tile']))
except Exception as e: print(sys.argv[1],'ERR',e)
PY
}
b base HB_X=0
b matrix HB_DEBUG=256
b matrix_ng12 HB_DEBUG=256 HB_NG=12
b ng6 HB_NG=6
b ng16 HB_NG=16

