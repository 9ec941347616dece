#!/bin/bash
O=gpurun_out/an; mkdir -p $O
timeout 900 python -m pytest tests/test_dropin_rcpp.py -q -x > $O/pytest_1.log 2>&1; echo "rc=$?" >> $O/pytest_1.log; tail -40 $O/pytest_1.log
