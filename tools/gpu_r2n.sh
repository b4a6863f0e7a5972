#!/bin/bash
# round 2, GPU call N (1 GPU): K1 with the front/back stage split, K2 with one warp per unit (30 warps)
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $O/r2n_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2n_pytest.log
SWEEP_FRAMES=10000 timeout 900 python tools/k1_sweep.py CB200_K1_L2_AHEAD 0,0,4096 > $O/r2n_k1.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r2n_clean.json 2> $O/r2n_clean.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --workload errors1pct > $O/r2n_errors1pct.json 2> $O/r2n_errors1pct.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --mode 67 > $O/r2n_mode67.json 2> $O/r2n_mode67.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --mode 66 > $O/r2n_mode66.json 2> $O/r2n_mode66.err
echo done
