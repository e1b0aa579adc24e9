#!/bin/bash
# round-2 GPU check N: tensor-core cost kernel with the staging window prefetched into registers
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sgm_parity_gpu.py -m gpu -x -q -k "cost or host_api or pair or full_size" 2>&1 | tail -4
timeout 200 python scripts/profile_sgm.py 128 default 2>&1 | tail -1 | tee gpurun_out/sgm_tc5.txt
timeout 200 python scripts/profile_sgm.py 64 default 2>&1 | tail -1 | tee -a gpurun_out/sgm_tc5.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sgm_cost_tc_kernel -s 1 -c 1 -o gpurun_out/sgm_cost_tc -f python scripts/profile_sgm.py 128 default > gpurun_out/ncu_tc.log 2>&1
timeout 60 python scripts/ncu_summary.py gpurun_out/sgm_cost_tc.ncu-rep > gpurun_out/ncu_sgm_cost_tc.txt 2>&1; head -12 gpurun_out/ncu_sgm_cost_tc.txt
