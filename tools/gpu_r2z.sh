#!/bin/bash
# round 2, GPU call Z (1 GPU): compute-sanitizer over the kernels of the second half of the round
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 230 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py > $O/r2z_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/r2z_memcheck.log
SANITIZE_CAMERA=0 timeout 110 compute-sanitizer --tool racecheck --racecheck-report all python tools/sanitize_small.py > $O/r2z_racecheck_no_k1x.log 2>&1; echo "racecheck rc=$?" >> $O/r2z_racecheck_no_k1x.log
tail -4 $O/r2z_memcheck.log; grep -E "RACECHECK SUMMARY|rc=" $O/r2z_racecheck_no_k1x.log | tail -3
echo done
