#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== targeted tests"
timeout 900 python -m pytest tests/test_gan_gpu.py tests/test_model_gpu.py tests/test_bench_config_gpu.py tests/test_kernels_gpu.py -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_s6.log 2>&1; echo "exit $?"; tail -n 8 gpurun_out/pytest_s6.log
echo "=== timeline"
timeout 500 python tools/graph_timeline.py gen_loss 512 gpurun_out/timeline_gen.csv 2>&1 | grep -v -i "warn\|shards\|Loading" | head -n 16 | tee gpurun_out/timeline_gen.txt
