#!/bin/bash
mkdir -p gpurun_out/h
O=gpurun_out/h
run() { # name, env..., -- args
  local name=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python bench.py --no-cpu "$@" > $O/$name.json 2> $O/$name.err
  python - "$name" <<'PY'
import json,sys
nm=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/h/%s.json'%nm).read().strip().splitlines()[-1])
    print(nm, 'ms/step %.3f kernel_ms %.3f rounds %.1f changed %.0f' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['config']['scalar_rounds_per_sweep'], d['config']['changed_snps_per_sweep']))
except Exception as e:
    print(nm, 'FAILED', e)
PY
  grep -h "serial CTA\|phases serial\|error\|Error" $O/$name.err | tail -2
}
run ns2 HB_NS=2 HB_PHASES=1 -- --steps 5 --warmup 3
run ns3 HB_NS=3 HB_PHASES=1 -- --steps 5 --warmup 3
run ns2_stream HB_NS=2 HB_DEBUG=32 -- --steps 5 --warmup 3
run ns3_stream HB_NS=3 HB_DEBUG=32 -- --steps 5 --warmup 3
run ns2_trace HB_NS=2 HB_TRACE=$O/trace.bin -- --steps 3 --warmup 5
python tools/trace_report.py $O/trace.bin 8 > $O/trace_report.txt 2>&1; tail -10 $O/trace_report.txt
rm -f $O/trace.bin
