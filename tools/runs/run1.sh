set -x
mkdir -p gpurun_out/r2
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2/smi.txt
bash run_gpu_tests.sh > gpurun_out/r2/tests_a.txt 2>&1
timeout 900 python -m pytest tests/test_gan_gpu.py tests/test_graphs_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/r2/tests_b.txt 2>&1; tail -3 gpurun_out/r2/tests_b.txt
python tools/gan_bench.py 512 2 1 > gpurun_out/r2/gan512_graphs.log 2>&1; cat gpurun_out/r2/gan512_graphs.log
python tools/gan_bench.py 512 1 0 gen_loss > gpurun_out/r2/gan512_gen_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 22000 -c 30000 --csv --log-file gpurun_out/r2/launches_gen512.csv python tools/gan_bench.py 512 1 0 gen_loss > gpurun_out/r2/ncu_gen.log 2>&1
tail -2 gpurun_out/r2/ncu_gen.log
