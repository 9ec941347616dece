#!/bin/bash
O=gpurun_out/ak; mkdir -p $O
timeout 600 python -m pytest tests/test_ldmat_bed_gpu.py tests/test_sbayes.py -m gpu -q > $O/pytest_1.log 2>&1; echo "rc=$?" >> $O/pytest_1.log; tail -8 $O/pytest_1.log
timeout 300 python tools/bench_ldmat.py --n 5000 --m 30000 --cpu-m 0 > $O/ld_sym.log 2>&1; tail -3 $O/ld_sym.log
HB_LD_FULL=0 timeout 300 python tools/bench_ldmat.py --n 5000 --m 30000 --cpu-m 0 > $O/ld_panels.log 2>&1; tail -3 $O/ld_panels.log
