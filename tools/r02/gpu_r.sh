#!/bin/bash
# device-side non-SNP effects: the single-step and covariate/random-effect tests (1 GPU), then the sharded tests on 2 GPUs
O=gpurun_out/r; mkdir -p $O
timeout 900 python -m pytest tests/test_single_step.py tests/test_gpu_parity.py -m gpu -q -x > $O/pytest_1.log 2>&1; echo "rc=$?" >> $O/pytest_1.log
tail -15 $O/pytest_1.log
timeout 900 python -m pytest tests/test_sharded.py -m gpu -q > $O/pytest_2.log 2>&1; echo "rc=$?" >> $O/pytest_2.log
tail -25 $O/pytest_2.log
