#!/bin/bash
# the row-sharded tests on 2 GPUs with the final tree (loader through pinned bounce buffers)
O=gpurun_out/au; mkdir -p $O
timeout 600 python -m pytest tests/test_sharded.py -m gpu -q > $O/pytest_2gpu.log 2>&1; echo "rc=$?" >> $O/pytest_2gpu.log; tail -6 $O/pytest_2gpu.log
