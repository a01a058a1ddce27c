set -x
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_bench_config_gpu.py tests/test_boundary_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/r2/tests_new.txt 2>&1; tail -30 gpurun_out/r2/tests_new.txt
timeout 900 python -m pytest tests/test_gan_gpu.py tests/test_graphs_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/r2/tests_gan.txt 2>&1; tail -30 gpurun_out/r2/tests_gan.txt
timeout 900 python bench.py --steps 10 --warmup 5 > gpurun_out/r2/bench_a.json 2> gpurun_out/r2/bench_a.err; tail -c 3000 gpurun_out/r2/bench_a.json; tail -5 gpurun_out/r2/bench_a.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2/bench_ref.json 2> gpurun_out/r2/bench_ref.err; tail -c 1500 gpurun_out/r2/bench_ref.json; tail -5 gpurun_out/r2/bench_ref.err
