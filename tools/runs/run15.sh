set -x
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_dp_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/r2/t15_2gpu.txt 2>&1; grep -n "graphs=\|DP_WORKER\|passed\|failed" gpurun_out/r2/t15_2gpu.txt | head
