set -x
mkdir -p gpurun_out/r2
timeout 120 python tools/decode_bench.py 512 64 10 > gpurun_out/r2/decode_bench.txt 2>&1 && cat gpurun_out/r2/decode_bench.txt && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:relattn_dec -s 12 -c 4 -o gpurun_out/r2/dec_prof python tools/decode_bench.py 512 64 2 > gpurun_out/r2/ncu_dec.log 2>&1
tail -3 gpurun_out/r2/ncu_dec.log
