#!/bin/bash
# rows2 / dt=2 correction from shared memory + sequential chain default: parity, then timing + trace; tcgen05 LD kernel tests + timing
O=gpurun_out/w; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_scalar_modes_gpu.py tests/test_sbayes.py -m gpu -q -x > $O/pytest_1.log 2>&1; echo "rc=$?" >> $O/pytest_1.log
tail -6 $O/pytest_1.log
timeout 600 python bench.py --no-cpu --no-product --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; python - <<'PY'
import json
j=json.loads(open('gpurun_out/w/bench.json').read().strip().splitlines()[-1]); print('bench ms=%.2f kms=%.2f rpt=%.3f'%(j['ms_per_step'], j['roofline']['kernel_ms'], j['config']['rounds_per_tile']))
PY
HB_DEBUG=256 timeout 600 python bench.py --no-cpu --no-product --steps 20 --warmup 5 > $O/bench_mat.json 2> $O/bench_mat.err; tail -c 300 $O/bench_mat.json | head -c 10; python - <<'PY'
import json
j=json.loads(open('gpurun_out/w/bench_mat.json').read().strip().splitlines()[-1]); print('matrix chain ms=%.2f kms=%.2f'%(j['ms_per_step'], j['roofline']['kernel_ms']))
PY
for L in 5 6; do timeout 600 python bench.py --no-cpu --no-product --steps 20 --warmup 5 --lag $L > $O/bench_d$L.json 2> $O/bench_d$L.err; python - $L <<'PY'
import json,sys
j=json.loads(open('gpurun_out/w/bench_d%s.json'%sys.argv[1]).read().strip().splitlines()[-1]); print('lag',sys.argv[1],'ms=%.2f kms=%.2f'%(j['ms_per_step'], j['roofline']['kernel_ms']))
PY
done
HB_TRACE=$O/trace.bin timeout 600 python bench.py --no-cpu --no-product --steps 5 --warmup 5 > $O/trace.json 2> $O/trace.err
python tools/trace_report.py $O/trace.bin 8 > $O/trace_report.txt 2>&1; rm -f $O/trace.bin; head -22 $O/trace_report.txt
HB_BENCH_FOLD_SCALE=64 timeout 600 python bench.py --no-cpu --no-product --steps 20 --warmup 10 > $O/bench_fold64.json 2> $O/bench_fold64.err; python - <<'PY'
import json
j=json.loads(open('gpurun_out/w/bench_fold64.json').read().strip().splitlines()[-1]); print('fold64 ms=%.2f kms=%.2f rpt=%.3f'%(j['ms_per_step'], j['roofline']['kernel_ms'], j['config']['rounds_per_tile']))
PY
timeout 900 python -m pytest tests/test_ldmat_bed_gpu.py -m gpu -q -x > $O/pytest_ld.log 2>&1; echo "rc=$?" >> $O/pytest_ld.log; tail -8 $O/pytest_ld.log
timeout 600 python tools/bench_ldmat.py --n 5000 --m 30000 --cpu-m 0 > $O/ld_tc.json 2> $O/ld_tc.err; tail -1 $O/ld_tc.json
HB_LD_MMA_SYNC=1 timeout 600 python tools/bench_ldmat.py --n 5000 --m 30000 --cpu-m 0 > $O/ld_mmasync.json 2> $O/ld_mmasync.err; tail -1 $O/ld_mmasync.json
