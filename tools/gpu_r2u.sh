#!/bin/bash
# round 2, GPU call U (1 GPU): the device anchor scan -- parity tests, facade, camera bench, launch list, one full capture of K1 (sharpen)
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 420 python -m pytest tests/test_gpu_scan.py tests/test_cimbard_facade.py -m gpu -q -x > $O/r2u_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2u_pytest.log
timeout 240 python bench.py --camera --frames 128 --steps 5 --warmup 3 > $O/r2u_camera.json 2> $O/r2u_camera.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r2u_launches_camera.csv python bench.py --camera --frames 64 --steps 1 --warmup 3 --no-cpu-baseline > $O/r2u_camera_ncu.log 2>&1
timeout 240 ncu --set full --clock-control none --import-source on -k regex:k1_decode_kernel -s 3 -c 1 -o $O/r2u_k1_sharpen python bench.py --sharpen --frames 4000 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2u_k1_sharpen_ncu.log 2>&1
tail -3 $O/r2u_pytest.log
echo done
