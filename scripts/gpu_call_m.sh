#!/bin/bash
# round-2 GPU check M: wave-front kernel holding two claimed items per warp (runs whichever is ready)
mkdir -p gpurun_out
echo "== sgm tests"
timeout 900 python -m pytest tests/test_sgm_parity_gpu.py -m gpu -x -q 2>&1 | tail -4
echo "== grid"
timeout 600 python scripts/profile_sgm.py 128 order 2>&1 | tail -32 | tee gpurun_out/sgm_defer.txt
echo "== D=64 / D=256 defaults"
timeout 200 python scripts/profile_sgm.py 64 default 2>&1 | tail -1 | tee -a gpurun_out/sgm_defer.txt
timeout 200 python scripts/profile_sgm.py 256 default 2>&1 | tail -1 | tee -a gpurun_out/sgm_defer.txt
echo "== ncu: DRAM bytes of the wave-front kernel: default, lag 0, lag 2"
for cfg in "default" "lag0" "lag2"; do
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:sgm_front -c 3 --csv --log-file gpurun_out/front_$cfg.csv python scripts/profile_sgm.py 128 $cfg > /dev/null 2>&1
grep "sgm_front" gpurun_out/front_$cfg.csv | tail -3 | awk -F'","' '{print "'$cfg'", $(NF-2), $(NF)}'
done
