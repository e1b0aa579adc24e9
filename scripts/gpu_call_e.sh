#!/bin/bash
# round-2 GPU check E: wave-front kernel after the convergence fix, tensor-core cost kernel with the issuing warp,
# sweep kernel with one propagation+refinement loop (evaluation cap 0 / 7)
mkdir -p gpurun_out
echo "== sgm tests"
timeout 900 python -m pytest tests/test_sgm_parity_gpu.py -m gpu -x -q 2>&1 | tail -5
echo "== sgm variants"
timeout 600 python scripts/profile_sgm.py 128 2>&1 | tail -70 | tee gpurun_out/sgm_variants.txt
echo "== D=64 / D=256 defaults"
timeout 200 python scripts/profile_sgm.py 64 default 2>&1 | tail -2 | tee -a gpurun_out/sgm_variants.txt
timeout 200 python scripts/profile_sgm.py 256 default 2>&1 | tail -2 | tee -a gpurun_out/sgm_variants.txt
echo "== sweep kernel: cap 0 / 7 / 6"
for cap in 0 7 6; do timeout 200 python scripts/profile_sweep.py 6 2 1 0 0 $cap 2>&1 | tail -2; done | tee gpurun_out/pm_cap.txt
echo "== pm parity tests (cap 0: bit-identical results expected)"
timeout 1500 python -m pytest tests/test_pm_parity_gpu.py tests/test_real_fixture.py tests/test_cpp_adapter.py -m gpu -q -x 2>&1 | tail -15
echo "== ncu: wave-front kernel (default), tensor-core cost kernel"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sgm_front_kernel -s 1 -c 1 -o gpurun_out/sgm_front -f python scripts/profile_sgm.py 128 default > gpurun_out/ncu_front.log 2>&1
timeout 60 python scripts/ncu_summary.py gpurun_out/sgm_front.ncu-rep 0 > gpurun_out/ncu_sgm_front.txt 2>&1; head -32 gpurun_out/ncu_sgm_front.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sgm_cost_tc_kernel -s 1 -c 1 -o gpurun_out/sgm_cost_tc -f python scripts/profile_sgm.py 128 tc > gpurun_out/ncu_tc.log 2>&1
timeout 60 python scripts/ncu_summary.py gpurun_out/sgm_cost_tc.ncu-rep > gpurun_out/ncu_sgm_cost_tc.txt 2>&1; head -24 gpurun_out/ncu_sgm_cost_tc.txt
