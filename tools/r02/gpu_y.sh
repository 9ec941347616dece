#!/bin/bash
# integer dots by default: whole GPU suite, bench, trace, ncu; LD tcgen05 v2 timing
O=gpurun_out/y; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
tail -8 $O/pytest_gpu.log
timeout 600 python bench.py --no-cpu --no-product --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; python - <<'PY'
import json
j=json.loads(open('gpurun_out/y/bench.json').read().strip().splitlines()[-1]); print('bench ms=%.2f kms=%.2f rpt=%.3f frac=%.3f'%(j['ms_per_step'], j['roofline']['kernel_ms'], j['config']['rounds_per_tile'], j['roofline']['frac']))
PY
HB_TRACE=$O/trace.bin timeout 600 python bench.py --no-cpu --no-product --steps 5 --warmup 5 > $O/trace.json 2> $O/trace.err
python tools/trace_report.py $O/trace.bin 8 > $O/trace_report.txt 2>&1; rm -f $O/trace.bin; head -22 $O/trace_report.txt
HB_DEBUG=1 timeout 600 python bench.py --no-cpu --no-product --steps 20 --warmup 5 > $O/noaxpy.json 2>/dev/null; python - <<'PY'
import json
j=json.loads(open('gpurun_out/y/noaxpy.json').read().strip().splitlines()[-1]); print('noaxpy ms=%.2f kms=%.2f'%(j['ms_per_step'], j['roofline']['kernel_ms']))
PY
timeout 600 python tools/bench_ldmat.py --n 5000 --m 30000 --cpu-m 0 2>/dev/null | grep '"ldmat"' | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 3 -c 1 -f -o $O/prof_full python bench.py --no-cpu --no-product --m 200000 --steps 1 --warmup 3 > $O/ncu_full.log 2>&1
ls -la $O | tail -5
