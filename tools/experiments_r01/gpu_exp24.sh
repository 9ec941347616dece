#!/bin/bash
# loop-latency reductions (async gather, direct dot polls, one-trip update fetch, 16-wide AXPY loads) + lag sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/e24_pytest.txt
cat gpurun_out/e24_pytest.txt
for LAG in 5 6 7 8; do
  HB_TRACE=gpurun_out/e24_trace.bin timeout 300 python bench.py --no-cpu --steps 8 --warmup 4 --lag $LAG > gpurun_out/e24_lag$LAG.json 2> gpurun_out/e24_lag$LAG.err
  echo "== lag $LAG"
  python -c "
import json; d=json.loads(open('gpurun_out/e24_lag$LAG.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['config']['scalar_rounds_per_sweep'])"
  python tools/trace_report.py gpurun_out/e24_trace.bin $LAG > gpurun_out/e24_report_lag$LAG.txt
  grep -h "period\|P done\|hand-over\|published\|dots" gpurun_out/e24_report_lag$LAG.txt
  rm -f gpurun_out/e24_trace.bin
done
