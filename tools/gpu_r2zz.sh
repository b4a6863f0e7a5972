#!/bin/bash
# round 2, GPU call ZZ (1 GPU): why the sharpen decode differed under racecheck
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 60 python tools/diag_sharpen.py > $O/r2zz_plain.log 2>&1
timeout 100 compute-sanitizer --tool racecheck --racecheck-report all python tools/diag_sharpen.py > $O/r2zz_racecheck.log 2>&1
CB200_K1_SHARPEN=0 timeout 100 compute-sanitizer --tool racecheck python tools/diag_sharpen.py > $O/r2zz_racecheck_old_route.log 2>&1
tail -5 $O/r2zz_plain.log; grep -E "rep|one frame|SUMMARY" $O/r2zz_racecheck.log | tail -8
echo done
