#!/bin/bash
# Round-2 final validation on one B200: GPU tests, smoke, the bench line, ncu evidence, graph timelines.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/smi.txt 2>&1
echo "=== pytest -m gpu"
timeout 1500 python -m pytest tests -x -q -m gpu --durations=12 -p no:cacheprovider > gpurun_out/pytest_gpu_final.log 2>&1; echo "exit $?"; tail -n 18 gpurun_out/pytest_gpu_final.log
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
echo "=== bench"; timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "exit $?"; head -c 400 gpurun_out/bench_final.json; echo
echo "=== ncu launch list: MLE step"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/r2_launches_mle.csv python tools/prof_cycle.py 512 mle > gpurun_out/ncu_mle.log 2>&1; echo "exit $?"
echo "=== ncu --set full: attention + single-token attention"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:relattn_.*_tc -s 1 -c 2 \
  -o gpurun_out/r2_attn_full -f python tools/attn_bench.py 512 1 > gpurun_out/ncu_attn.log 2>&1; echo "exit $?"
echo "=== timelines"
timeout 500 python tools/graph_timeline.py gen_loss 512 2>&1 | grep -v -i "warn\|shards\|Loading" | head -n 30 > gpurun_out/timeline_gen.txt
timeout 500 python tools/graph_timeline.py dis_loss 512 2>&1 | grep -v -i "warn\|shards\|Loading" | head -n 30 > gpurun_out/timeline_dis.txt
head -n 3 gpurun_out/timeline_gen.txt gpurun_out/timeline_dis.txt
