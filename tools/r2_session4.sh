#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== timeline"
timeout 600 python tools/graph_timeline.py gen_loss 512 gpurun_out/timeline_gen.csv 2>&1 | grep -v Warning | tail -n 40 | tee gpurun_out/timeline_gen.txt
echo "=== targeted tests"
timeout 900 python -m pytest tests/test_gan_gpu.py tests/test_model_gpu.py tests/test_bench_config_gpu.py tests/test_graphs_gpu.py tests/test_bert_gpu.py tests/test_relattn_gpu.py tests/test_kernels_gpu.py -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_s3.log 2>&1; echo "exit $?"; tail -n 12 gpurun_out/pytest_s3.log
