#!/bin/bash
# round-2 GPU check D: wave-front kernel after the convergence fix: parity, variant grid, ncu capture
mkdir -p gpurun_out
echo "== sgm front tests"
timeout 900 python -m pytest tests/test_sgm_parity_gpu.py -m gpu -x -q -k "wave_front or variants" 2>&1 | tail -5
echo "== sgm variants"
timeout 600 python scripts/profile_sgm.py 128 2>&1 | tail -70 | tee gpurun_out/sgm_variants.txt
echo "== D=64 / D=256 defaults"
timeout 200 python scripts/profile_sgm.py 64 default 2>&1 | tail -2 | tee -a gpurun_out/sgm_variants.txt
timeout 200 python scripts/profile_sgm.py 256 default 2>&1 | tail -2 | tee -a gpurun_out/sgm_variants.txt
echo "== ncu: wave-front kernel (default)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sgm_front_kernel -s 1 -c 1 -o gpurun_out/sgm_front -f python scripts/profile_sgm.py 128 default > gpurun_out/ncu_front.log 2>&1
timeout 60 python scripts/ncu_summary.py gpurun_out/sgm_front.ncu-rep 0 > gpurun_out/ncu_sgm_front.txt 2>&1; head -32 gpurun_out/ncu_sgm_front.txt
