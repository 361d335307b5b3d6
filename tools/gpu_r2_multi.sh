#!/bin/bash
# 2 GPUs of one box: the driver's torchrun launch of bench.py (both arms) and the 2-GPU test cases
O=gpurun_out/m2; mkdir -p $O
nvidia-smi -L > $O/smi.txt
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 ) > $O/bench_2gpu.json 2> $O/bench_2gpu.err
tail -c 400 $O/bench_2gpu.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 1 ) > $O/bench_2gpu_ref.json 2> $O/bench_2gpu_ref.err
tail -c 200 $O/bench_2gpu_ref.err
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "two_gpus or 2gpu or two_gpu" > $O/pytest_2gpu.log 2>&1
tail -3 $O/pytest_2gpu.log
