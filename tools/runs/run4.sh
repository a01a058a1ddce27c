set -x
mkdir -p gpurun_out/r2
timeout 300 python -m pytest tests/test_relattn_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/r2/t4_relattn.txt 2>&1; tail -15 gpurun_out/r2/t4_relattn.txt
timeout 120 python tools/attn_bench.py 512 5 0.1 > gpurun_out/r2/attn_bench_a.log 2>&1; cat gpurun_out/r2/attn_bench_a.log
timeout 300 python -m pytest tests/test_bert_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/r2/t4_bert.txt 2>&1; tail -40 gpurun_out/r2/t4_bert.txt
timeout 400 python -m pytest tests/test_model_gpu.py tests/test_gan_gpu.py tests/test_boundary_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/r2/t4_model.txt 2>&1; tail -15 gpurun_out/r2/t4_model.txt
timeout 300 python tools/gan_bench.py 512 2 1 > gpurun_out/r2/gan512_dec.log 2>&1; cat gpurun_out/r2/gan512_dec.log | tail -5
