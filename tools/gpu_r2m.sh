#!/bin/bash
# round 2, GPU call M (1 GPU): K1 single-stage-ring layout compiled for 4 CTAs/SM (128 registers) vs 5 (96)
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
CB200_NVCC_EXTRA="-DCB200_K1_MIN_CTAS=4" python -m libcimbar_b200.build --force > $O/r2m_build4.log 2>&1
SWEEP_FRAMES=10000 timeout 900 python tools/k1_sweep.py CB200_K1_CTAS_PER_SM 4,5,4 > $O/r2m_k1_min4.log 2>&1
python -m libcimbar_b200.build --force > $O/r2m_build5.log 2>&1
SWEEP_FRAMES=10000 timeout 900 python tools/k1_sweep.py CB200_K1_CTAS_PER_SM 5 > $O/r2m_k1_min5.log 2>&1
echo done
