#!/bin/bash
# last 1-GPU pass: the GPU suite and smoke on the final library, the default bench line once more, config-5 surrogate
O=gpurun_out/final2; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log; tail -4 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; tail -1 $O/smoke.txt
timeout 900 python bench.py --no-cpu > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -c 1200 $O/bench.json
timeout 900 python bench.py --config c5 > $O/c5.json 2> $O/c5.err; echo "c5 rc=$?"; tail -c 1500 $O/c5.json; tail -3 $O/c5.err
