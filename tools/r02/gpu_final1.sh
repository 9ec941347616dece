#!/bin/bash
# final 1-GPU pass of round 2: what the driver runs (GPU suite, smoke, both bench arms) + launch list, DRAM traffic, c2, c4
O=gpurun_out/final1; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log; tail -5 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; tail -2 $O/smoke.txt
timeout 600 python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err; tail -c 400 $O/bench_ref.json
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -c 2500 $O/bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 15 -c 48 --csv --log-file $O/launches.csv python bench.py --no-cpu --no-product --steps 4 --warmup 3 > $O/ncu_launch.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_sweep -s 3 -c 1 --csv --log-file $O/traffic_full.csv python bench.py --no-cpu --no-product --steps 1 --warmup 3 > $O/ncu_traffic.log 2>&1
tail -3 $O/traffic_full.csv | cut -c1-250
timeout 1500 python bench.py --config c2 > $O/c2.json 2> $O/c2.err; echo "c2 rc=$?"; tail -c 1500 $O/c2.json
timeout 900 python bench.py --config c4 --m 60000 --niter 30 > $O/c4_m60k.json 2> $O/c4.err; echo "c4 rc=$?"; tail -c 900 $O/c4_m60k.json
