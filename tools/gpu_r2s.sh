#!/bin/bash
# round 2, GPU call S (1 GPU): frame-pair RS kernel for the legacy coupled 6-bit stream (mode 4C)
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $O/r2s_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2s_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --mode 4 > $O/r2s_mode4.json 2> $O/r2s_mode4.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --mode 4 --workload errors1pct > $O/r2s_mode4_errors.json 2> $O/r2s_mode4_errors.err
CB200_K2_FRAMES=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --mode 4 > $O/r2s_mode4_k2old.json 2> $O/r2s_mode4_k2old.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2s_clean.json 2> $O/r2s_clean.err
echo done
