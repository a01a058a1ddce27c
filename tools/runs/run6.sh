set -x
mkdir -p gpurun_out/r2
timeout 600 python bench.py --steps 5 --warmup 5 --no-cpu-baseline > gpurun_out/r2/bench_b.json 2> gpurun_out/r2/bench_b.err; tail -c 1800 gpurun_out/r2/bench_b.json; tail -3 bench_b.err
timeout 300 python tools/cycle_prof.py gen_loss 512 > gpurun_out/r2/prof_gen.txt 2>&1; head -45 gpurun_out/r2/prof_gen.txt
timeout 300 python tools/cycle_prof.py dis_loss 512 > gpurun_out/r2/prof_dis.txt 2>&1; head -45 gpurun_out/r2/prof_dis.txt
TGAN_B200_LIB=transformer-gan_b200/tgan_b200/libtgan_b200_s4.so timeout 120 python tools/attn_bench.py 512 5 0.1 > gpurun_out/r2/attn_bench_s4.log 2>&1; cat gpurun_out/r2/attn_bench_s4.log
TGAN_B200_LIB=transformer-gan_b200/tgan_b200/libtgan_b200_prof.so timeout 120 python tools/bwd_phase_prof.py 512 > gpurun_out/r2/bwd_phase_r2a.log 2>&1; cat gpurun_out/r2/bwd_phase_r2a.log
