#!/bin/bash
# round 2, final evidence call (1 GPU): tests, every bench line, ncu captures of the final kernels, launch lists, sanitizer
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/r2o_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2o_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2o_smoke.log 2>&1
timeout 600 python bench.py > $O/r2o_bench_n1.json 2> $O/r2o_bench_n1.err
timeout 600 python bench.py --impl reference > $O/r2o_reference.json 2> $O/r2o_reference.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload errors1pct > $O/r2o_errors1pct.json 2> $O/r2o_errors1pct.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --color-correction 2 > $O/r2o_cc2.json 2> $O/r2o_cc2.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --color-correction 1 > $O/r2o_cc1.json 2> $O/r2o_cc1.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --mode 4 > $O/r2o_mode4.json 2> $O/r2o_mode4.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --mode 4 --workload errors1pct > $O/r2o_mode4_errors.json 2> $O/r2o_mode4_errors.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --mode 67 > $O/r2o_mode67.json 2> $O/r2o_mode67.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --workload noise1pct --frames 9472 > $O/r2o_noise.json 2> $O/r2o_noise.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --workload noise1pct --frames 3552 > $O/r2o_noise_3552.json 2> $O/r2o_noise_3552.err
timeout 300 python bench.py --fountain --steps 5 --warmup 1 > $O/r2o_fountain_n1.json 2> $O/r2o_fountain_n1.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1_decode_kernel -s 4 -c 1 -o $O/r2o_k1 -f \
  python bench.py --steps 2 --warmup 3 --frames 4144 --no-e2e --no-cpu-baseline > $O/r2o_ncu_k1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rs_frames -s 4 -c 1 -o $O/r2o_k2 -f \
  python bench.py --steps 2 --warmup 3 --frames 4144 --no-e2e --no-cpu-baseline > $O/r2o_ncu_k2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rs_frames -s 4 -c 1 -o $O/r2o_k2_err -f \
  python bench.py --steps 2 --warmup 3 --frames 4144 --no-e2e --no-cpu-baseline --workload errors1pct > $O/r2o_ncu_k2_err.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_flood_walk -s 3 -c 1 -o $O/r2o_walk -f \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --workload noise1pct --frames 4736 > $O/r2o_ncu_walk.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_flood_raster -s 3 -c 1 -o $O/r2o_raster -f \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --workload noise1pct --frames 1776 > $O/r2o_ncu_raster.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2o_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $O/r2o_launches.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2o_launches_noise.csv \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --workload noise1pct --frames 4736 > $O/r2o_launches_noise.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py > $O/r2o_memcheck.log 2>&1; echo "rc=$?" >> $O/r2o_memcheck.log
SANITIZE_CAMERA=0 timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all python tools/sanitize_small.py > $O/r2o_racecheck_no_k1x.log 2>&1; echo "rc=$?" >> $O/r2o_racecheck_no_k1x.log
echo done
