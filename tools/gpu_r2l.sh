#!/bin/bash
# round 2, GPU call L (1 GPU): K1 with a single-stage ring, 5 CTAs/SM
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $O/r2l_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2l_pytest.log
SWEEP_FRAMES=10000 timeout 900 python tools/k1_sweep.py CB200_K1_CTAS_PER_SM 5,4,5,6 > $O/r2l_k1_ctas.log 2>&1
SWEEP_FRAMES=10000 timeout 900 python tools/k1_sweep.py CB200_K1_L2_AHEAD 0,1,2,4096 > $O/r2l_k1_l2.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r2l_clean.json 2> $O/r2l_clean.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --mode 67 > $O/r2l_mode67.json 2> $O/r2l_mode67.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --color-correction 2 > $O/r2l_cc2.json 2> $O/r2l_cc2.err
echo done
