#!/bin/bash
# round 2, GPU call G (2 GPUs): the chunk-record exchange (NVLink window / NCCL / torch) and the fountain config
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi topo -m > $O/r2g_topo.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > $O/r2g_pytest.log 2>&1; prc=$?; echo "pytest rc=$prc" >> $O/r2g_pytest.log
if [ $prc -ne 0 ]; then   # the frame-pair K2 is new in this call: fall back to the per-warp kernel for the rest if it broke anything
  export CB200_K2_FRAMES=0
  timeout 900 python -m pytest tests -m gpu -q -x > $O/r2g_pytest_k2old.log 2>&1; echo "pytest rc=$?" >> $O/r2g_pytest_k2old.log
fi
# K2 A/B on one GPU
for w in clean errors1pct; do
  CB200_K2_FRAMES=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --workload $w > $O/r2g_k2old_$w.json 2> $O/r2g_k2old_$w.err
  timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --workload $w > $O/r2g_k2new_$w.json 2> $O/r2g_k2new_$w.err
done
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r2g_n1.json 2> $O/r2g_n1.err
for g in window nccl torch; do
  timeout 600 $T bench.py --gpus 2 --steps 10 --warmup 3 --gather $g --no-cpu-baseline > $O/r2g_n2_$g.json 2> $O/r2g_n2_$g.err
done
timeout 600 $T bench.py --gpus 2 --fountain --steps 3 --warmup 1 > $O/r2g_fountain_n2.json 2> $O/r2g_fountain_n2.err
timeout 600 $T bench.py --gpus 2 --fountain --steps 3 --warmup 1 --gather nccl > $O/r2g_fountain_n2_nccl.json 2> $O/r2g_fountain_n2_nccl.err
timeout 600 python bench.py --fountain --steps 5 --warmup 1 > $O/r2g_fountain_n1.json 2> $O/r2g_fountain_n1.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --workload noise1pct --frames 9472 > $O/r2g_noise.json 2> $O/r2g_noise.err
echo done
