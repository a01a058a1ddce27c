set -x
mkdir -p gpurun_out/r2
timeout 300 python -m pytest tests/test_relattn_gpu.py tests/test_gan_gpu.py tests/test_model_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/r2/t7.txt 2>&1; tail -8 gpurun_out/r2/t7.txt
timeout 60 tools/mma_rate > gpurun_out/r2/mma_rate.txt 2>&1; cat gpurun_out/r2/mma_rate.txt
timeout 400 python bench.py --extras-only > gpurun_out/r2/extras_c.json 2> gpurun_out/r2/extras_c.err; tail -c 1500 gpurun_out/r2/extras_c.json; tail -3 gpurun_out/r2/extras_c.err
