#!/bin/bash
O=gpurun_out/at; mkdir -p $O
timeout 600 python -m pytest tests/test_class_decisions_gpu.py -q -s > $O/pytest_1.log 2>&1; echo "rc=$?" >> $O/pytest_1.log; tail -6 $O/pytest_1.log
timeout 600 python -m pytest tests/test_dropin_rcpp.py tests/test_gpu_parity.py -m gpu -q -x > $O/pytest_2.log 2>&1; echo "rc=$?" >> $O/pytest_2.log; tail -4 $O/pytest_2.log
