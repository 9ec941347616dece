#!/bin/bash
# r02 baseline: GPU suite, both bench arms, ncu launch list, DRAM traffic of k_sweep at the metric shape, --set full capture (m=200k), tile trace
O=gpurun_out/q; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
grep -E "passed|failed|FAILED|rc=|error" $O/pytest_gpu.log | tail -20
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -c 1500 $O/bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 15 -c 40 --csv --log-file $O/launches.csv python bench.py --no-cpu --no-product --steps 4 --warmup 3 > $O/ncu_launch.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_sweep -s 3 -c 1 --csv --log-file $O/traffic_full.csv python bench.py --no-cpu --no-product --steps 1 --warmup 3 > $O/ncu_traffic.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 3 -c 1 -f -o $O/prof_full python bench.py --no-cpu --no-product --m 200000 --steps 1 --warmup 3 > $O/ncu_full.log 2>&1
HB_TRACE=$O/trace.bin timeout 600 python bench.py --no-cpu --no-product --steps 5 --warmup 5 > $O/trace.json 2> $O/trace.err
python tools/trace_report.py $O/trace.bin 8 > $O/trace_report.txt 2>&1; rm -f $O/trace.bin
tail -3 $O/traffic_full.csv | cut -c1-300; ls -la $O
