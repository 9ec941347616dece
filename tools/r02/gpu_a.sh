#!/bin/bash
# Round 2, GPU visit A: what paces the round-1 kernel? (sub-stages in flight, lag, workers, speculation lead,
# streaming floor fp64 vs integer limbs, steady state of a long chain)
mkdir -p gpurun_out/a
O=gpurun_out/a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/smi.txt 2>&1
run() { # name, env..., -- args
  local name=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python bench.py --no-cpu "$@" > $O/$name.json 2> $O/$name.err
  python - "$name" <<'PY'
import json,sys
nm=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/a/%s.json'%nm).read().strip().splitlines()[-1])
    print(nm, 'ms/step %.3f kernel_ms %.3f rounds %.1f changed %.0f' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['config']['scalar_rounds_per_sweep'], d['config']['changed_snps_per_sweep']))
except Exception as e:
    print(nm, 'FAILED', e)
PY
  grep -h "re-speculated\|phases worker 0" $O/$name.err | tail -2
}
run base HB_PHASES=1 -- --steps 10 --warmup 5
run ns2 HB_NS=2 -- --steps 10 --warmup 5
run ns3 HB_NS=3 -- --steps 10 --warmup 5
run ns6 HB_NS=6 -- --steps 10 --warmup 5
run lag3 HB_X=0 -- --steps 10 --warmup 5 --lag 3
run lag8 HB_X=0 -- --steps 10 --warmup 5 --lag 8
run ng4 HB_NG=4 -- --steps 10 --warmup 5
run ng12 HB_NG=12 -- --steps 10 --warmup 5
run ng16 HB_NG=16 -- --steps 10 --warmup 5
run ng16lag8 HB_NG=16 -- --steps 10 --warmup 5 --lag 8
run lead1 HB_LEAD=1 HB_PHASES=1 -- --steps 5 --warmup 25
run lead2 HB_LEAD=2 HB_PHASES=1 -- --steps 5 --warmup 25
run lead4 HB_LEAD=4 HB_PHASES=1 -- --steps 5 --warmup 25
run lead4lag8 HB_LEAD=4 HB_PHASES=1 -- --steps 5 --warmup 25 --lag 8
run stream_fp64 HB_DEBUG=32 -- --steps 5 --warmup 3
run stream_limbs HB_DEBUG=32 HB_LIMBS=1 -- --steps 5 --warmup 3
run limbs HB_LIMBS=1 -- --steps 10 --warmup 5
run stream_fp64_ns6 HB_DEBUG=32 HB_NS=6 -- --steps 5 --warmup 3
run stream_nodot HB_DEBUG=34 -- --steps 5 --warmup 3
run steady HB_PHASES=1 -- --steps 20 --warmup 300
run trace HB_TRACE=$O/trace.bin -- --steps 3 --warmup 5
python tools/trace_report.py $O/trace.bin 5 > $O/trace_report.txt 2>&1; cat $O/trace_report.txt
rm -f $O/trace.bin
