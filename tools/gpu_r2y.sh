#!/bin/bash
# round 2, GPU call Y (1 GPU): k_scan_blur4 -- the whole GPU suite, smoke(), camera line with the CPU pipeline beside it, headline, one capture of the blur kernel
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1100 python -m pytest tests -m gpu -q > $O/r2y_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2y_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2y_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r2y_smoke.log
timeout 300 python bench.py --camera --steps 3 --warmup 3 > $O/r2y_camera.json 2> $O/r2y_camera.err
timeout 400 python bench.py --steps 10 --warmup 3 > $O/r2y_n1.json 2> $O/r2y_n1.err
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:k_scan_blur' -c 1 -o $O/r2y_blur python bench.py --camera --frames 256 --steps 1 --warmup 3 --no-cpu-baseline > $O/r2y_blur_ncu.log 2>&1
tail -3 $O/r2y_pytest.log
echo done
