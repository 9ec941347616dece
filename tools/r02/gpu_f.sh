#!/bin/bash
mkdir -p gpurun_out/g
O=gpurun_out/g
timeout 900 python -m pytest tests/test_serial_gpu.py -q -k "every_lag or mixture or agree or many_candidates or flip_heavy" > $O/pytest_serial.log 2>&1
echo "rc=$?" >> $O/pytest_serial.log; grep -E "passed|failed|FAILED|rc=" $O/pytest_serial.log | tail -20
run() { # name, env..., -- args
  local name=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python bench.py --no-cpu "$@" > $O/$name.json 2> $O/$name.err
  python - "$name" <<'PY'
import json,sys
nm=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/g/%s.json'%nm).read().strip().splitlines()[-1])
    print(nm, 'ms/step %.3f kernel_ms %.3f rounds %.1f changed %.0f' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['config']['scalar_rounds_per_sweep'], d['config']['changed_snps_per_sweep']))
except Exception as e:
    print(nm, 'FAILED', e)
PY
  grep -h "serial CTA\|phases serial\|error\|Error" $O/$name.err | tail -2
}
run serial HB_X=1 -- --steps 10 --warmup 5
run serial_ph HB_PHASES=1 -- --steps 10 --warmup 5
run abl256 HB_DEBUG=256 HB_PHASES=1 -- --steps 5 --warmup 3
run abl768 HB_DEBUG=768 HB_PHASES=1 -- --steps 5 --warmup 3
run serial_trace HB_TRACE=$O/trace.bin -- --steps 3 --warmup 5
python tools/trace_report.py $O/trace.bin 8 > $O/trace_report.txt 2>&1; tail -24 $O/trace_report.txt
rm -f $O/trace.bin
