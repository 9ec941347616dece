#!/bin/bash
O=gpurun_out/al; mkdir -p $O
timeout 900 python -m pytest tests/test_reference_pin.py -q > $O/pytest_1.log 2>&1; echo "rc=$?" >> $O/pytest_1.log; tail -15 $O/pytest_1.log
