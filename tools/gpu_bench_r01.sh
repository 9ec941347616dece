#!/bin/bash
# Round-1 measurement pass: bench at 1 and 2 GPUs, reference arm, ncu launch list and one full capture of k_sweep.
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT 2>/dev/null || true
python bench.py --steps 10 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "n1 rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "n2 rc=$?"
fi
ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --no-cpu --steps 3 --warmup 3 > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 3 -c 1 -f -o gpurun_out/prof_sweep_r01 python bench.py --no-cpu --m 200000 --steps 1 --warmup 3 > gpurun_out/ncu_full.log 2>&1
tail -c 1500 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
tail -c 600 gpurun_out/bench_ref.json
tail -c 1500 gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
