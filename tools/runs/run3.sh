set -x
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_relattn_gpu.py tests/test_model_gpu.py tests/test_gan_gpu.py tests/test_boundary_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/r2/tests_dec.txt 2>&1; tail -30 gpurun_out/r2/tests_dec.txt
python tools/gan_bench.py 512 2 1 > gpurun_out/r2/gan512_dec.log 2>&1; cat gpurun_out/r2/gan512_dec.log
