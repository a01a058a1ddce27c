set -x
mkdir -p gpurun_out/r2
timeout 300 python -m pytest tests/test_relattn_gpu.py tests/test_gan_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/r2/t11.txt 2>&1; tail -4 gpurun_out/r2/t11.txt
timeout 120 python tools/decode_bench.py 512 64 2 > gpurun_out/r2/decode_bench3.txt 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:relattn_dec_bwd -s 6 -c 2 -o gpurun_out/r2/decb_prof python tools/decode_bench.py 512 64 2 > gpurun_out/r2/ncu_decb.log 2>&1
tail -3 gpurun_out/r2/ncu_decb.log
