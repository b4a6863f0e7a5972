#!/bin/bash
# round 2, GPU call W (1 GPU): whole heap in shared memory for small batches of walks -- parity of every walk test, latency A/B
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scan.py -m gpu -q -x -k "golden or camera or trace or walk or mixing or noise or degenerate or sharpen or extract or ccm or color" > $O/r2w_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2w_pytest.log
timeout 200 python bench.py --steps 3 --warmup 3 --frames 2000 --no-cpu-baseline > $O/r2w_n1_latency.json 2> $O/r2w_n1_latency.err
CB200_K1X_HEAP_SMEM=1023 timeout 200 python bench.py --steps 3 --warmup 3 --frames 2000 --no-cpu-baseline > $O/r2w_n1_latency_heap1023.json 2> $O/r2w_n1_latency_heap1023.err
timeout 200 python bench.py --camera --frames 512 --steps 3 --warmup 3 --no-cpu-baseline > $O/r2w_camera512.json 2> $O/r2w_camera512.err
CB200_K1X_HEAP_SMEM=1023 timeout 200 python bench.py --camera --frames 512 --steps 3 --warmup 3 --no-cpu-baseline > $O/r2w_camera512_heap1023.json 2> $O/r2w_camera512_heap1023.err
timeout 200 python bench.py --steps 3 --warmup 3 --frames 592 --no-cpu-baseline --no-e2e --workload noise1pct > $O/r2w_noise592.json 2> $O/r2w_noise592.err
CB200_K1X_HEAP_SMEM=1023 timeout 200 python bench.py --steps 3 --warmup 3 --frames 592 --no-cpu-baseline --no-e2e --workload noise1pct > $O/r2w_noise592_heap1023.json 2> $O/r2w_noise592_heap1023.err
tail -3 $O/r2w_pytest.log
echo done
