#!/bin/bash
# round 2, GPU call H (1 GPU): facade test, frame-pair K2 under ncu and the sanitizer, the remaining BASELINE configs
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2h_pytest.log
timeout 600 python bench.py > $O/r2h_bench_n1.json 2> $O/r2h_bench_n1.err
timeout 600 python bench.py --impl reference > $O/r2h_reference.json 2> $O/r2h_reference.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload errors1pct > $O/r2h_errors1pct.json 2> $O/r2h_errors1pct.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --color-correction 2 > $O/r2h_cc2.json 2> $O/r2h_cc2.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --color-correction 1 > $O/r2h_cc1.json 2> $O/r2h_cc1.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --mode 4 > $O/r2h_mode4.json 2> $O/r2h_mode4.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --mode 4 --workload errors1pct > $O/r2h_mode4_errors.json 2> $O/r2h_mode4_errors.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --mode 67 > $O/r2h_mode67.json 2> $O/r2h_mode67.err
timeout 300 python bench.py --fountain --steps 5 --warmup 1 > $O/r2h_fountain_n1.json 2> $O/r2h_fountain_n1.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rs_frames -s 4 -c 1 -o $O/r2h_k2 -f \
  python bench.py --steps 2 --warmup 3 --frames 4144 --no-e2e --no-cpu-baseline > $O/r2h_ncu_k2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rs_frames -s 4 -c 1 -o $O/r2h_k2_err -f \
  python bench.py --steps 2 --warmup 3 --frames 4144 --no-e2e --no-cpu-baseline --workload errors1pct > $O/r2h_ncu_k2_err.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py > $O/r2h_memcheck.log 2>&1; echo "rc=$?" >> $O/r2h_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all python tools/sanitize_small.py > $O/r2h_racecheck.log 2>&1; echo "rc=$?" >> $O/r2h_racecheck.log
echo done
