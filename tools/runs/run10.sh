set -x
mkdir -p gpurun_out/r2
timeout 300 python -m pytest tests/test_relattn_gpu.py tests/test_gan_gpu.py tests/test_abi_cpu.py -q -m "gpu or not gpu" -p no:cacheprovider > gpurun_out/r2/t10.txt 2>&1; tail -5 gpurun_out/r2/t10.txt
timeout 120 python tools/decode_bench.py 512 64 10 > gpurun_out/r2/decode_bench2.txt 2>&1; cat gpurun_out/r2/decode_bench2.txt
timeout 120 python tools/decode_bench.py 512 127 10 >> gpurun_out/r2/decode_bench2.txt 2>&1; tail -2 gpurun_out/r2/decode_bench2.txt
