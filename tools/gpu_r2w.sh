#!/bin/bash
# round 2, GPU call W (1 GPU): the build with the whole-heap-in-shared-memory walks for small batches and the branch-free deskew --
# the whole GPU suite, smoke(), latency A/B of the heap placement, camera lines, headline sanity
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1100 python -m pytest tests -m gpu -q > $O/r2w_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2w_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2w_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r2w_smoke.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/r2w_n1.json 2> $O/r2w_n1.err
CB200_K1X_HEAP_SMEM=1023 timeout 300 python bench.py --steps 3 --warmup 3 --frames 2000 --no-cpu-baseline > $O/r2w_n1_latency_heap1023.json 2> $O/r2w_n1_latency_heap1023.err
timeout 300 python bench.py --camera --steps 3 --warmup 3 --no-cpu-baseline > $O/r2w_camera.json 2> $O/r2w_camera.err
timeout 200 python bench.py --camera --frames 512 --steps 3 --warmup 3 --no-cpu-baseline > $O/r2w_camera512.json 2> $O/r2w_camera512.err
CB200_K1X_HEAP_SMEM=1023 timeout 200 python bench.py --camera --frames 512 --steps 3 --warmup 3 --no-cpu-baseline > $O/r2w_camera512_heap1023.json 2> $O/r2w_camera512_heap1023.err
timeout 200 python bench.py --steps 3 --warmup 3 --frames 592 --no-cpu-baseline --no-e2e --workload noise1pct > $O/r2w_noise592.json 2> $O/r2w_noise592.err
CB200_K1X_HEAP_SMEM=1023 timeout 200 python bench.py --steps 3 --warmup 3 --frames 592 --no-cpu-baseline --no-e2e --workload noise1pct > $O/r2w_noise592_heap1023.json 2> $O/r2w_noise592_heap1023.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r2w_launches_camera.csv python bench.py --camera --frames 256 --steps 1 --warmup 3 --no-cpu-baseline > $O/r2w_camera_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:k_scan_blur|k_scan_anchors|k_deskew' -c 3 -o $O/r2w_camera_kernels python bench.py --camera --frames 256 --steps 1 --warmup 3 --no-cpu-baseline > $O/r2w_camera_kernels_ncu.log 2>&1
tail -3 $O/r2w_pytest.log
echo done
