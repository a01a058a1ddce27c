#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/r2_launches_gan.csv python tools/prof_cycle.py 512 gan > gpurun_out/ncu_gan.log 2>&1; echo "exit $?"
tail -n 2 gpurun_out/ncu_gan.log
