set -x
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_dropin_scripts_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2/t13_dropin.txt 2>&1; tail -60 gpurun_out/r2/t13_dropin.txt
