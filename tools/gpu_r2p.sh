#!/bin/bash
# round 2, GPU call P (2 GPUs): the copy-engine push into the window vs the direct stores; K2 table syndromes
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $O/r2p_pytest.log 2>&1; prc=$?; echo "pytest rc=$prc" >> $O/r2p_pytest.log
if [ $prc -ne 0 ]; then export CB200_K2_FRAMES=0; fi
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --workload errors1pct > $O/r2p_errors1pct.json 2> $O/r2p_errors1pct.err
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531"
for g in window window-direct nccl; do
  timeout 600 $T bench.py --gpus 2 --steps 10 --warmup 3 --gather $g --no-cpu-baseline --no-e2e > $O/r2p_n2_$g.json 2> $O/r2p_n2_$g.err
done
timeout 600 $T bench.py --gpus 2 --fountain --steps 3 --warmup 1 > $O/r2p_fountain_n2.json 2> $O/r2p_fountain_n2.err
timeout 600 $T bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --workload errors1pct > $O/r2p_n2_errors.json 2> $O/r2p_n2_errors.err
echo done
