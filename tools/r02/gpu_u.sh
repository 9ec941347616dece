#!/bin/bash
O=gpurun_out/u; mkdir -p $O
timeout 900 python -m pytest tests/test_single_step.py tests/test_abi_compiled.py tests/test_ldmat_bed_gpu.py -m gpu -q > $O/pytest_1.log 2>&1; echo "rc=$?" >> $O/pytest_1.log
tail -12 $O/pytest_1.log
timeout 600 python tools/bench_ldmat.py --n 50000 --m 200000 --gebv 64 > $O/gebv.json 2> $O/gebv.err; tail -2 $O/gebv.json; tail -3 $O/gebv.err
timeout 1200 python bench.py --config c4 --m 60000 --niter 30 > $O/c4_m60k.json 2> $O/c4.err; tail -c 1800 $O/c4_m60k.json; tail -3 $O/c4.err
