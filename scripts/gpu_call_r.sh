#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sgm_aggregate_kernel -s 8 -c 2 -o gpurun_out/sgm_ragged -f python - <<'PY' > gpurun_out/ncu_ragged.log 2>&1
import json, torch, bench
dev = torch.device("cuda", 0)
print(json.dumps(bench.sgm_block(dev, 6585.4, 16)["tsgm_ragged"]))
PY
tail -2 gpurun_out/ncu_ragged.log | cut -c1-300
for i in 0 1; do timeout 60 python scripts/ncu_summary.py gpurun_out/sgm_ragged.ncu-rep $i > gpurun_out/ncu_sgm_ragged_$i.txt 2>&1; head -24 gpurun_out/ncu_sgm_ragged_$i.txt; done
