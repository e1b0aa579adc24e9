#!/bin/bash
# round-2 GPU check A: PatchMatch parity + schedules + SGM front kernel parity + timings + bench
mkdir -p gpurun_out
echo "== sgm tests"
timeout 600 python -m pytest tests/test_sgm_parity_gpu.py -m gpu -x -q 2>&1 | tail -8
echo "== sgm variants"
timeout 300 python scripts/profile_sgm.py 128 2>&1 | tail -12 | tee gpurun_out/sgm_variants.txt
echo "== pm tests"
timeout 1200 python -m pytest tests/test_pm_parity_gpu.py tests/test_image_prep_gpu.py tests/test_real_fixture.py tests/test_cpp_adapter.py -m gpu -x -q 2>&1 | tail -15
echo "== schedules"
for cfg in "6 2 1 0" "6 2 0 0" "6 0 0 2" "6 0 1 2" "6 1 1 0"; do timeout 120 python scripts/profile_sweep.py $cfg 2>&1 | tail -3; done | tee gpurun_out/pm_schedules.txt
echo "== bench"
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; tail -c 1500 gpurun_out/bench_a.json; tail -3 gpurun_out/bench_a.err
