#!/bin/bash
# round-2 GPU session G (2 GPUs): one-process multi-GPU overlap test, torchrun bench at N=2 with e2e_single_process, gather
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L > gpurun_out/g_smi_$N.txt
timeout 900 python -m pytest tests/test_gpu_host_paths.py -m gpu -x -q > gpurun_out/g_pytest_$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/g_pytest_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/g_bench_n$N.json 2> gpurun_out/g_bench_n$N.err
timeout 600 python tools/bench_gather.py $N > gpurun_out/g_gather_n$N.log 2>&1
