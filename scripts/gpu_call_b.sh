#!/bin/bash
# round-2 GPU check B: remaining parity tests, SGM variants after the fixes, ncu captures of the shipped kernels, bench
mkdir -p gpurun_out
echo "== sgm tests"
timeout 600 python -m pytest tests/test_sgm_parity_gpu.py -m gpu -x -q 2>&1 | tail -6
echo "== sgm variants"
timeout 400 python scripts/profile_sgm.py 128 2>&1 | tail -40 | tee gpurun_out/sgm_variants.txt
echo "== pm + other tests"
timeout 1500 python -m pytest tests/test_pm_parity_gpu.py tests/test_image_prep_gpu.py tests/test_real_fixture.py tests/test_cpp_adapter.py tests/test_filter_parity_gpu.py -m gpu -q 2>&1 | tail -40
echo "== ncu: wave-front kernel (default layout, both passes), tensor-core cost kernel"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sgm_front_kernel -s 2 -c 2 -o gpurun_out/sgm_front -f python scripts/profile_sgm.py 128 default > gpurun_out/ncu_front.log 2>&1
for i in 0 1; do timeout 60 python scripts/ncu_summary.py gpurun_out/sgm_front.ncu-rep $i > gpurun_out/ncu_sgm_front_pass$i.txt 2>&1; done; head -30 gpurun_out/ncu_sgm_front_pass0.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sgm_cost_tc_kernel -s 1 -c 1 -o gpurun_out/sgm_cost_tc -f python scripts/profile_sgm.py 128 tc > gpurun_out/ncu_tc.log 2>&1
timeout 60 python scripts/ncu_summary.py gpurun_out/sgm_cost_tc.ncu-rep > gpurun_out/ncu_sgm_cost_tc.txt 2>&1; head -30 gpurun_out/ncu_sgm_cost_tc.txt
echo "== ncu: shipped sweep kernel (launch 12 of a run from random initialisation)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pm_sweep -s 12 -c 1 -o gpurun_out/pm_sweep -f python scripts/profile_sweep.py 6 > gpurun_out/ncu_pm.log 2>&1
timeout 60 python scripts/ncu_summary.py gpurun_out/pm_sweep.ncu-rep > gpurun_out/ncu_pm_sweep.txt 2>&1; head -12 gpurun_out/ncu_pm_sweep.txt
echo "== 4 CTAs/SM variant"
timeout 120 python scripts/profile_sweep.py 6 2 1 0 1 2>&1 | tail -3 | tee gpurun_out/pm_4ctas.txt
echo "== bench"
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; tail -c 2500 gpurun_out/bench_b.json; tail -3 gpurun_out/bench_b.err
