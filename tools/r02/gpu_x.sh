#!/bin/bash
O=gpurun_out/x; mkdir -p $O
b() { name=$1; shift; env "$@" timeout 600 python bench.py --no-cpu --no-product --steps 20 --warmup 5 > $O/$name.json 2> $O/$name.err; python - $name <<'PY'
import json,sys
try:
    j=json.loads(open('gpurun_out/x/%s.json'%sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1],'ms=%.2f kms=%.2f rpt=%.3f'%(j['ms_per_step'], j['roofline']['kernel_ms'], j['config']['rounds_per_tile']))
except Exception as e: print(sys.argv[1],'ERR',e)
PY
}
b base HB_X=0
b limbs HB_LIMBS=1
b xevict0 HB_XEVICT=0
b limbs_xevict0 HB_LIMBS=1 HB_XEVICT=0
b ns3 HB_NS=3
b ns6 HB_NS=6
b ng12 HB_NG=12
b stream32 HB_DEBUG=32
b stream32_limbs HB_DEBUG=32 HB_LIMBS=1
b noaxpy HB_DEBUG=1
b limbs_ng12 HB_LIMBS=1 HB_NG=12
