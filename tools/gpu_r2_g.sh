#!/usr/bin/env bash
# round 2, GPU call G (4 GPUs): BASELINE config 5 strong scaling at N = 2 and 4 (8M matrices), weak scaling at N = 4 with the
# NUMA-bound e2e leg, and the two-devices-in-one-process parity test that needs > 1 GPU
mkdir -p gpurun_out/r2g
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n bench.py --gpus $n "$@"; }
run 2 --batch 8388608 --steps 6 --warmup 3 > gpurun_out/r2g/bench_8M_2gpu.json 2> gpurun_out/r2g/bench_8M_2gpu.err; cut -c1-200 gpurun_out/r2g/bench_8M_2gpu.json
run 4 --batch 8388608 --steps 6 --warmup 3 > gpurun_out/r2g/bench_8M_4gpu.json 2> gpurun_out/r2g/bench_8M_4gpu.err; cut -c1-200 gpurun_out/r2g/bench_8M_4gpu.json
run 4 --steps 20 --warmup 3 > gpurun_out/r2g/bench_weak_4gpu.json 2> gpurun_out/r2g/bench_weak_4gpu.err; cut -c1-200 gpurun_out/r2g/bench_weak_4gpu.json
run 2 --steps 20 --warmup 3 > gpurun_out/r2g/bench_weak_2gpu.json 2> gpurun_out/r2g/bench_weak_2gpu.err; cut -c1-200 gpurun_out/r2g/bench_weak_2gpu.json
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k two_devices > gpurun_out/r2g/pytest_2dev.log 2>&1; tail -2 gpurun_out/r2g/pytest_2dev.log
nvidia-smi topo -m > gpurun_out/r2g/topo.txt 2>&1; lscpu | grep -i "numa\|socket\|model name" > gpurun_out/r2g/lscpu.txt
