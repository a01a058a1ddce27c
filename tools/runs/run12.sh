set -x
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/r2/t12_all.txt 2>&1; tail -6 gpurun_out/r2/t12_all.txt
timeout 700 python bench.py --steps 10 --warmup 5 > gpurun_out/r2/bench_c.json 2> gpurun_out/r2/bench_c.err; tail -c 2200 gpurun_out/r2/bench_c.json; tail -3 gpurun_out/r2/bench_c.err
