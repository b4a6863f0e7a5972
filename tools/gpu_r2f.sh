#!/bin/bash
# round 2, GPU call D: K1x v3.4 + K2 (remainder + weighted-sum syndromes) -- tests, benches, ncu
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2f_pytest.log
CB200_K1X_SERIAL_ABOVE=0 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "golden or trace or camera or several_chunks" > $O/r2f_pytest_serial.log 2>&1; echo "rc=$?" >> $O/r2f_pytest_serial.log
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
for hs in 511 255 1023; do
  CB200_K1X_HEAP_SMEM=$hs timeout 300 $B --workload noise1pct --frames 9472 > $O/r2f_noise_hs$hs.json 2> $O/r2f_noise_hs$hs.err
done
timeout 300 $B --workload noise1pct --frames 3552 > $O/r2f_noise_3552.json 2> $O/r2f_noise_3552.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_flood_walk -s 3 -c 1 -o $O/r2f_walk -f \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --workload noise1pct --frames 4736 > $O/r2f_ncu_walk.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/r2f_clean.json 2> $O/r2f_clean.err
echo done
