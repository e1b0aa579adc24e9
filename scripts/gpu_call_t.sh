#!/bin/bash
mkdir -p gpurun_out
timeout 600 python - <<'PY' 2>&1 | tail -30 | tee gpurun_out/sgm_block.json
import json, torch, bench
dev = torch.device("cuda", 0)
b = bench.sgm_block(dev, 6585.4, 16)
print(json.dumps({k: b[k] for k in ("pair_D128",)}, indent=1)); print(b["fixed_range_D128"]["ms_match"], b["fixed_range_D128"]["ms_host_api"], b["tsgm_ragged"]["ms_match"])
PY
