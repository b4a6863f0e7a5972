#!/bin/bash
# round 2, GPU call V (1 GPU): final state -- the whole GPU suite, the headline line, sharpen lines (K1 and the exact walk), launch list
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1100 python -m pytest tests -m gpu -q -x > $O/r2v_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2v_pytest.log
timeout 400 python bench.py --steps 10 --warmup 3 > $O/r2v_n1.json 2> $O/r2v_n1.err
timeout 200 python bench.py --steps 3 --warmup 3 --frames 3552 --no-cpu-baseline --no-e2e --workload noise1pct --sharpen > $O/r2v_noise_sharpen.json 2> $O/r2v_noise_sharpen.err
CB200_K1X_SHARPEN_RASTER=0 timeout 200 python bench.py --steps 2 --warmup 3 --frames 1776 --no-cpu-baseline --no-e2e --workload noise1pct --sharpen > $O/r2v_noise_sharpen_old_raster.json 2> $O/r2v_noise_sharpen_old_raster.err
timeout 200 python bench.py --steps 3 --warmup 3 --frames 3552 --no-cpu-baseline --no-e2e --workload noise1pct > $O/r2v_noise.json 2> $O/r2v_noise.err
timeout 300 python bench.py --camera --steps 3 --warmup 3 > $O/r2v_camera.json 2> $O/r2v_camera.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/r2v_launches_sharpen.csv python bench.py --sharpen --frames 4000 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2v_sharpen_ncu.log 2>&1
tail -3 $O/r2v_pytest.log
echo done
