#!/bin/bash
# round 2, GPU call R (4 GPUs): N = 4 with the push exchange, the two-rank exchange tests, single-frame walk latency
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
T4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541"
timeout 600 $T4 bench.py --gpus 4 --steps 10 --warmup 3 --gather window --no-cpu-baseline > $O/r2r_n4_window.json 2> $O/r2r_n4_window.err
timeout 900 python -m pytest tests/test_gpu_exchange.py -m gpu -q > $O/r2r_pytest_exchange.log 2>&1; echo "pytest rc=$?" >> $O/r2r_pytest_exchange.log
timeout 600 python bench.py --no-cpu-baseline > $O/r2r_n1.json 2> $O/r2r_n1.err
echo done
