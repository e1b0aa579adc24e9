#!/bin/bash
# round-2 GPU check G: wave-front kernel with sub-cell phase dependencies (lag 0): parity, grid, DRAM traffic
mkdir -p gpurun_out
echo "== sgm tests"
timeout 900 python -m pytest tests/test_sgm_parity_gpu.py tests/test_dmap_io.py -m gpu -x -q 2>&1 | tail -5
echo "== sgm variants"
timeout 600 python scripts/profile_sgm.py 128 2>&1 | tail -50 | tee gpurun_out/sgm_variants.txt
echo "== D=64 / D=256 defaults"
timeout 200 python scripts/profile_sgm.py 64 default 2>&1 | tail -2 | tee -a gpurun_out/sgm_variants.txt
timeout 200 python scripts/profile_sgm.py 256 default 2>&1 | tail -2 | tee -a gpurun_out/sgm_variants.txt
echo "== ncu: one SGM Match, every kernel: time and DRAM bytes"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:sgm -c 40 --csv --log-file gpurun_out/launches_sgm.csv python scripts/profile_sgm.py 128 default > gpurun_out/ncu_sgm.log 2>&1; grep "sgm_front\|wta" gpurun_out/launches_sgm.csv | tail -6 | cut -c1-40,150-330
echo "== ncu --set full: wave-front kernel"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sgm_front_kernel -s 1 -c 1 -o gpurun_out/sgm_front -f python scripts/profile_sgm.py 128 default > gpurun_out/ncu_front.log 2>&1
timeout 60 python scripts/ncu_summary.py gpurun_out/sgm_front.ncu-rep 0 > gpurun_out/ncu_sgm_front.txt 2>&1; head -32 gpurun_out/ncu_sgm_front.txt
