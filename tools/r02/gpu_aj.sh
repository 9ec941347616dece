#!/bin/bash
O=gpurun_out/aj; mkdir -p $O
timeout 900 python -m pytest tests/test_single_step.py tests/test_abi_compiled.py tests/test_gpu_parity.py -m gpu -q > $O/pytest_1.log 2>&1; echo "rc=$?" >> $O/pytest_1.log; tail -15 $O/pytest_1.log
