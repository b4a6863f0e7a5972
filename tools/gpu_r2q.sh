#!/bin/bash
# round 2, GPU call K (8 GPUs): the scaling points N = 8 and 4 with the window / NCCL exchange, fountain config on 8 GPUs
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi topo -m > $O/r2q_topo.txt 2>&1
T8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
T4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522"
timeout 600 $T8 bench.py --gpus 8 --steps 10 --warmup 3 --gather window --no-cpu-baseline > $O/r2q_n8_window.json 2> $O/r2q_n8_window.err
timeout 600 $T8 bench.py --gpus 8 --steps 10 --warmup 3 --gather window-direct --no-cpu-baseline --no-e2e > $O/r2q_n8_window_direct.json 2> $O/r2q_n8_window_direct.err
timeout 600 $T8 bench.py --gpus 8 --fountain --steps 3 --warmup 1 > $O/r2q_fountain_n8.json 2> $O/r2q_fountain_n8.err
echo done
