#!/bin/bash
# round 2, GPU call X (1 GPU): deskew with aligned word loads -- the whole GPU suite, smoke(), camera lines, one capture of k_deskew
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1100 python -m pytest tests -m gpu -q > $O/r2x_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2x_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2x_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r2x_smoke.log
timeout 300 python bench.py --camera --steps 3 --warmup 3 > $O/r2x_camera.json 2> $O/r2x_camera.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/r2x_n1.json 2> $O/r2x_n1.err
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:k_deskew' -c 1 -o $O/r2x_deskew python bench.py --camera --frames 256 --steps 1 --warmup 3 --no-cpu-baseline > $O/r2x_deskew_ncu.log 2>&1
tail -3 $O/r2x_pytest.log
echo done
