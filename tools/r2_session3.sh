#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== targeted tests"
timeout 900 python -m pytest tests/test_gan_gpu.py tests/test_model_gpu.py tests/test_bench_config_gpu.py tests/test_graphs_gpu.py tests/test_bert_gpu.py tests/test_relattn_gpu.py tests/test_kernels_gpu.py -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_s3.log 2>&1; echo "exit $?"; tail -n 25 gpurun_out/pytest_s3.log
echo "=== phase timing"
for cfg in "8192" "0"; do
  timeout 300 python tools/phase_time.py $cfg 2>&1 | grep -E "side_rows=|Error|error" | tail -n 4
done > gpurun_out/phase_time.txt 2>&1
cat gpurun_out/phase_time.txt
