#!/bin/bash
# round 2, GPU call J (1 GPU): K1 cursor split + load-only ceiling, K2 lane-parallel Forney / packed BM
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $O/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2j_pytest.log
# l2_ahead sweep with the split cursors, the round-1 placement (0x2000 + n), and the copy-only ceiling (0x1000 + n)
SWEEP_FRAMES=6000 timeout 600 python tools/k1_sweep.py CB200_K1_L2_AHEAD 4,8196,2,6,8,12,4100,4104,4112 > $O/r2j_k1_sweep.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2j_clean.json 2> $O/r2j_clean.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --workload errors1pct > $O/r2j_errors1pct.json 2> $O/r2j_errors1pct.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --mode 4 --workload errors1pct > $O/r2j_mode4_errors.json 2> $O/r2j_mode4_errors.err
echo done
