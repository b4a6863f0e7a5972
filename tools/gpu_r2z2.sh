#!/bin/bash
# round 2, GPU call Z2 (1 GPU): racecheck of everything but the exact walk, with the script's CCM reset fixed
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
SANITIZE_CAMERA=0 timeout 120 compute-sanitizer --tool racecheck --racecheck-report all python tools/sanitize_small.py > $O/r2z2_racecheck_no_k1x.log 2>&1; echo "racecheck rc=$?" >> $O/r2z2_racecheck_no_k1x.log
grep -E "RACECHECK SUMMARY|rc=|sanitize_small" $O/r2z2_racecheck_no_k1x.log | tail -4
echo done
