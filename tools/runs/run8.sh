set -x
mkdir -p gpurun_out/r2
timeout 300 python -m pytest tests/test_relattn_gpu.py tests/test_gan_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/r2/t8.txt 2>&1; tail -5 gpurun_out/r2/t8.txt
timeout 300 python tools/cycle_prof.py gen_loss 512 > gpurun_out/r2/prof_gen2.txt 2>&1; head -32 gpurun_out/r2/prof_gen2.txt | tail -28
timeout 300 python tools/cycle_prof.py mle 512 > gpurun_out/r2/prof_mle.txt 2>&1; head -40 gpurun_out/r2/prof_mle.txt | tail -36
