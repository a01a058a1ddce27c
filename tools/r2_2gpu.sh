#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=index,name --format=csv
echo "=== 2-GPU tests"
timeout 600 python -m pytest tests -x -q -m gpu -k "two_rank or two_gpu or ddp or dp" -p no:cacheprovider 2>&1 | tail -n 6 | tee gpurun_out/pytest_2gpu.log
echo "=== bench --gpus 2 (weak scaling)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 5 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "exit $?"; tail -c 300 gpurun_out/bench_2gpu.err; head -c 600 gpurun_out/bench_2gpu.json; echo
echo "=== reference arm under torchrun"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus 2 --steps 5 --warmup 1 --cpu-batch 2 2>/dev/null | tail -n 1 | head -c 300; echo
