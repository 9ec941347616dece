#!/bin/bash
O=gpurun_out/s; mkdir -p $O
timeout 900 python -m pytest tests/test_sbayes.py tests/test_single_step.py -m gpu -q > $O/pytest_1.log 2>&1; echo "rc=$?" >> $O/pytest_1.log
tail -30 $O/pytest_1.log
