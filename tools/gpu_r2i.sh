#!/bin/bash
# round 2, GPU call I (1 GPU): K1 with the multiply-high bit gather, K2 with conflict-free tables; ncu of K1 / K2, launch list
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $O/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2i_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r2i_clean.json 2> $O/r2i_clean.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --workload errors1pct > $O/r2i_errors1pct.json 2> $O/r2i_errors1pct.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --workload noise1pct --frames 9472 > $O/r2i_noise.json 2> $O/r2i_noise.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --workload noise1pct --frames 3552 > $O/r2i_noise_3552.json 2> $O/r2i_noise_3552.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1_decode_kernel -s 4 -c 1 -o $O/r2i_k1 -f \
  python bench.py --steps 2 --warmup 3 --frames 4144 --no-e2e --no-cpu-baseline > $O/r2i_ncu_k1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rs_frames -s 4 -c 1 -o $O/r2i_k2 -f \
  python bench.py --steps 2 --warmup 3 --frames 4144 --no-e2e --no-cpu-baseline > $O/r2i_ncu_k2.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2i_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $O/r2i_launches.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2i_launches_noise.csv \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --workload noise1pct --frames 4736 > $O/r2i_launches_noise.log 2>&1
SANITIZE_CAMERA=0 timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all python tools/sanitize_small.py > $O/r2i_racecheck_no_k1x.log 2>&1; echo "rc=$?" >> $O/r2i_racecheck_no_k1x.log
echo done
