#!/bin/bash
# round 2, GPU call T (1 GPU): sharpen inside K1 -- parity tests, bench --sharpen, headline sanity
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sharpen" > $O/r2t_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2t_pytest.log
timeout 240 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --sharpen > $O/r2t_sharpen.json 2> $O/r2t_sharpen.err
CB200_K1_SHARPEN=0 timeout 240 python bench.py --steps 2 --warmup 3 --frames 2048 --no-cpu-baseline --no-e2e --sharpen > $O/r2t_sharpen_old_route.json 2> $O/r2t_sharpen_old_route.err
timeout 240 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2t_clean.json 2> $O/r2t_clean.err
tail -3 $O/r2t_pytest.log
echo done
