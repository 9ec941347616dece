#!/bin/bash
# far corrections with 32 loads in flight; with and without cluster hand-over
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/e30_pytest.txt
cat gpurun_out/e30_pytest.txt
run() { local name=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu --steps 10 --warmup 5 > gpurun_out/e30_$name.json 2> gpurun_out/e30_$name.err
  echo "== $name"
  python -c "
import json; d=json.loads(open('gpurun_out/e30_$name.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['config']['scalar_rounds_per_sweep'], d['config']['changed_snps_per_sweep'])"
  grep -h "\[hb\]" gpurun_out/e30_$name.err | head -3
}
run plain HB_CLUSTER=0
run cluster HB_CLUSTER=1
HB_CLUSTER=1 HB_TRACE=gpurun_out/e30_trace.bin timeout 300 python bench.py --no-cpu --steps 6 --warmup 3 > /dev/null 2>&1
python tools/trace_report.py gpurun_out/e30_trace.bin 5 > gpurun_out/e30_report.txt; grep "period\|P later\|hand-over\|published" gpurun_out/e30_report.txt
rm -f gpurun_out/e30_trace.bin
