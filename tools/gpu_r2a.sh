#!/bin/bash
# round 2, GPU call A: baseline health + K1x knob sweep + fresh ncu captures of the current exact-walk kernels
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/r2a_gpus.txt 2>&1
nproc > $O/r2a_host.txt; grep -m1 "model name" /proc/cpuinfo >> $O/r2a_host.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2a_pytest.log
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 300 $B --workload noise1pct --frames 3552 > $O/r2a_noise_default.json 2> $O/r2a_noise_default.err
for hs in 2047 1023 511; do
  CB200_K1X_HEAP_SMEM=$hs timeout 300 $B --workload noise1pct --frames 7104 > $O/r2a_noise_hs$hs.json 2> $O/r2a_noise_hs$hs.err
done
CB200_K1X_HEAP_SMEM=4095 timeout 300 $B --workload noise1pct --frames 7104 > $O/r2a_noise_hs4095_7104.json 2> $O/r2a_noise_hs4095.err
timeout 300 $B --workload errors1pct > $O/r2a_errors1pct.json 2> $O/r2a_errors1pct.err
timeout 300 $B > $O/r2a_clean.json 2> $O/r2a_clean.err
# fresh counters for the current exact-walk kernels (one wave of walks)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_flood_walk -s 3 -c 1 -o $O/r2a_walk -f \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --workload noise1pct --frames 1776 > $O/r2a_ncu_walk.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_flood_raster -s 3 -c 1 -o $O/r2a_raster -f \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --workload noise1pct --frames 1776 > $O/r2a_ncu_raster.log 2>&1
echo done
