#!/bin/bash
# chain as a precomputed matrix-vector product; pre-check only after recent misses
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/e25_pytest.txt
cat gpurun_out/e25_pytest.txt
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu --steps 8 --warmup 4 > gpurun_out/e25_$name.json 2> gpurun_out/e25_$name.err
  echo "== $name"
  python -c "
import json; d=json.loads(open('gpurun_out/e25_$name.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['config']['scalar_rounds_per_sweep'])"
  grep -h "phases worker 0\|re-spec" gpurun_out/e25_$name.err | tail -2
}
run plain HB_X=0
run phases HB_PHASES=1
run nomat HB_DEBUG=128 HB_PHASES=1
run hard HB_BENCH_FOLD_SCALE=64
run hard_nomat HB_BENCH_FOLD_SCALE=64 HB_DEBUG=128
HB_TRACE=gpurun_out/e25_trace.bin timeout 300 python bench.py --no-cpu --steps 6 --warmup 3 > /dev/null 2>&1
python tools/trace_report.py gpurun_out/e25_trace.bin 5 | tee gpurun_out/e25_report.txt
rm -f gpurun_out/e25_trace.bin
