#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
SWEEP_FRAMES=10000 timeout 900 python tools/k1_sweep.py CB200_K1_L2_AHEAD 0,0,1,2,3,8192,8194,4096,4097,4098 > $O/r2j2_k1_sweep.log 2>&1
SWEEP_FRAMES=10000 CB200_K1_L2_AHEAD=4096 timeout 600 python tools/k1_sweep.py CB200_K1_CTAS_PER_SM 4,3,2,1 > $O/r2j2_k1_loadonly_ctas.log 2>&1
echo done
