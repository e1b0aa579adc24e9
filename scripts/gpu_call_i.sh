#!/bin/bash
# round-2 GPU check I: ragged (tSGM) aggregation with the eight directions side by side; traffic of the wave-front defaults
mkdir -p gpurun_out
echo "== sgm tests"
timeout 900 python -m pytest tests/test_sgm_parity_gpu.py -m gpu -x -q 2>&1 | tail -5
echo "== sanitizer on the ragged path (memcheck + racecheck)"
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python scripts/sanitize_small.py 2>&1 | tail -6
echo "== bench sgm block"
timeout 600 python - <<'PY' 2>&1 | tail -40 | tee gpurun_out/sgm_block.json
import json, torch, bench
dev = torch.device("cuda", 0)
print(json.dumps(bench.sgm_block(dev, 6585.4, 16), indent=1))
PY
echo "== ncu: DRAM bytes of the wave-front kernel: default (FB 32, lag 2), lag 1, FB 64"
for cfg in "default" "lag1" "fb64"; do
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:sgm_front -c 3 --csv --log-file gpurun_out/front_$cfg.csv python scripts/profile_sgm.py 128 $cfg > /dev/null 2>&1
grep "sgm_front" gpurun_out/front_$cfg.csv | tail -3 | awk -F'","' '{print "'$cfg'", $(NF-2), $(NF)}'
done
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:sgm -c 40 --csv --log-file gpurun_out/launches_sgm.csv python scripts/profile_sgm.py 128 default > gpurun_out/ncu_sgm.log 2>&1
