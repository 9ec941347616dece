#!/bin/bash
# First GPU visit of round 2: the kernels added at the end of round 1 (.bed decoder, LD builder, predict_samples;
# DESIGN.md 4b) have never run on hardware.  1) the verified suite, 2) the new tests with their xfail marker
# ignored, 3) memcheck of the new kernels on a small case, 4) their timings, 5) the usual bench line.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_zz_ldmat_bed_gpu.py > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python -m pytest tests/test_zz_ldmat_bed_gpu.py -m gpu -q --runxfail > gpurun_out/pytest_new.log 2>&1
echo "pytest(new) rc=$?" >> gpurun_out/pytest_new.log; tail -25 gpurun_out/pytest_new.log
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_zz_ldmat_bed_gpu.py -m gpu -q --runxfail \
  -k "ragged or missing_genotypes or imputes_over" > gpurun_out/memcheck_new.log 2>&1
tail -15 gpurun_out/memcheck_new.log
timeout 600 python tools/bench_ldmat.py --n 5000 --m 30000 > gpurun_out/bench_ldmat_dense.json 2> gpurun_out/bench_ldmat.err
timeout 600 python tools/bench_ldmat.py --n 5000 --m 30000 --chisq 3.84 > gpurun_out/bench_ldmat_sparse.json 2>> gpurun_out/bench_ldmat.err
cat gpurun_out/bench_ldmat_dense.json gpurun_out/bench_ldmat_sparse.json; tail -5 gpurun_out/bench_ldmat.err
# integer-dot variant of the streaming CTAs (HB_LIMBS=1): streaming side alone (HB_DEBUG=32), then coupled
HB_DEBUG=32 timeout 600 python bench.py --no-cpu --steps 5 --warmup 3 > gpurun_out/bench_stream_fp64.json 2>> gpurun_out/bench.err
HB_LIMBS=1 HB_DEBUG=32 timeout 600 python bench.py --no-cpu --steps 5 --warmup 3 > gpurun_out/bench_stream_limbs.json 2>> gpurun_out/bench.err
HB_LIMBS=1 timeout 600 python bench.py --no-cpu --steps 10 --warmup 5 > gpurun_out/bench_limbs.json 2>> gpurun_out/bench.err
for f in bench_stream_fp64 bench_stream_limbs bench_limbs; do python -c "import json,sys; d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]); print('$f', d['ms_per_step'], d['value'])"; done
# how complete is a candidate list built `lead` tiles early?: rebuilt lists and rounds per sweep
for lead in 0 1 2 4; do
  HB_LEAD=$lead HB_PHASES=1 timeout 600 python bench.py --no-cpu --steps 5 --warmup 25 > gpurun_out/bench_lead$lead.json 2> gpurun_out/bench_lead$lead.err
  echo "lead $lead: $(grep -c 're-speculated' gpurun_out/bench_lead$lead.err) sweeps; last: $(grep 're-speculated' gpurun_out/bench_lead$lead.err | tail -1)"
  python -c "import json; d=json.loads(open('gpurun_out/bench_lead$lead.json').read().strip().splitlines()[-1]); print('  ms/sweep', d['ms_per_step'], 'rounds/sweep', d['config'].get('scalar_rounds_per_sweep'))"
done
timeout 900 python bench.py --steps 10 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/bench.json
