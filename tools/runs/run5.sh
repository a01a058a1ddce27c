set -x
mkdir -p gpurun_out/r2
timeout 300 python -m pytest tests/test_relattn_gpu.py tests/test_sampling_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/r2/t5_relattn.txt 2>&1; tail -15 gpurun_out/r2/t5_relattn.txt
timeout 300 python -m pytest tests/test_bert_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/r2/t5_bert.txt 2>&1; tail -30 gpurun_out/r2/t5_bert.txt
timeout 400 python -m pytest tests/test_model_gpu.py tests/test_gan_gpu.py tests/test_boundary_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/r2/t5_model.txt 2>&1; tail -15 gpurun_out/r2/t5_model.txt
timeout 300 python tools/gan_bench.py 512 2 1 > gpurun_out/r2/gan512_dec.log 2>&1; cat gpurun_out/r2/gan512_dec.log | tail -5
timeout 600 python bench.py --steps 5 --warmup 5 --no-cpu-baseline > gpurun_out/r2/bench_b.json 2> gpurun_out/r2/bench_b.err; tail -c 2500 gpurun_out/r2/bench_b.json; tail -5 gpurun_out/r2/bench_b.err
