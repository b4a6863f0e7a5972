#!/bin/bash
# round 2, GPU call B: K1x v3.2 -- tests, knob sweep, fresh ncu of the walk, config benches
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2b_pytest.log
# the literal one-level pop on real frames (camera golden + trace test)
CB200_K1X_SERIAL_ABOVE=0 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or trace or camera or several_chunks" > $O/r2b_pytest_serial.log 2>&1; echo "rc=$?" >> $O/r2b_pytest_serial.log
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
for hs in 1023 2047 511 4095; do
  CB200_K1X_HEAP_SMEM=$hs timeout 300 $B --workload noise1pct --frames 9472 > $O/r2b_noise_hs$hs.json 2> $O/r2b_noise_hs$hs.err
done
for w in 16 24; do
  CB200_K1X_WALKS_PER_SM=$w timeout 300 $B --workload noise1pct --frames 9472 > $O/r2b_noise_w$w.json 2> $O/r2b_noise_w$w.err
done
timeout 300 $B --workload noise1pct --frames 3552 > $O/r2b_noise_3552.json 2> $O/r2b_noise_3552.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_flood_walk -s 3 -c 1 -o $O/r2b_walk -f \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --workload noise1pct --frames 4736 > $O/r2b_ncu_walk.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --workload errors1pct > $O/r2b_errors1pct.json 2> $O/r2b_errors1pct.err
timeout 600 python bench.py --steps 10 --warmup 3 --mode 4 > $O/r2b_mode4.json 2> $O/r2b_mode4.err
timeout 600 python bench.py --steps 10 --warmup 3 --mode 4 --workload errors1pct --no-cpu-baseline > $O/r2b_mode4_errors.json 2> $O/r2b_mode4_errors.err
timeout 600 python bench.py --steps 20 --warmup 3 > $O/r2b_clean.json 2> $O/r2b_clean.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2b_reference.json 2> $O/r2b_reference.err
echo done
