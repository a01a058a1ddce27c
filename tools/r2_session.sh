#!/bin/bash
# Round-2 GPU session: bench line, ncu launch lists of the cycle, ncu --set full of the attention kernels, GPU tests.
# Everything lands in gpurun_out/ as it goes (most important first).
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/smi.txt 2>&1
echo "=== bench"; timeout 900 python bench.py > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; echo "exit $?"; tail -c 600 gpurun_out/bench_r2.json
echo "=== attn_bench"; timeout 120 python tools/attn_bench.py 512 5 > gpurun_out/attn_bench.txt 2>&1; cat gpurun_out/attn_bench.txt
echo "=== ncu launch list: MLE step"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/r2_launches_mle.csv python tools/prof_cycle.py 512 mle > gpurun_out/ncu_mle.log 2>&1; echo "exit $?"
echo "=== ncu launch list: dis + gen update"
timeout 1500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/r2_launches_gan.csv python tools/prof_cycle.py 512 gan > gpurun_out/ncu_gan.log 2>&1; echo "exit $?"
echo "=== ncu --set full: attention"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:relattn_.*_tc -s 1 -c 2 \
  -o gpurun_out/r2_attn_full -f python tools/attn_bench.py 512 1 > gpurun_out/ncu_attn.log 2>&1; echo "exit $?"
echo "=== pytest -m gpu"
timeout 1500 python -m pytest tests -x -q -m gpu --durations=25 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?"; tail -n 40 gpurun_out/pytest_gpu.log
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 3
