set -x
mkdir -p gpurun_out/r2
nvidia-smi -L
timeout 600 python -m pytest tests/test_dp_gpu.py "tests/test_dropin_scripts_gpu.py::test_train_py_under_two_rank_ddp" -q -m gpu -p no:cacheprovider > gpurun_out/r2/t14_2gpu.txt 2>&1; tail -40 gpurun_out/r2/t14_2gpu.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 5 --warmup 5 > gpurun_out/r2/bench_n2.json 2> gpurun_out/r2/bench_n2.err; tail -c 1500 gpurun_out/r2/bench_n2.json; tail -5 gpurun_out/r2/bench_n2.err
