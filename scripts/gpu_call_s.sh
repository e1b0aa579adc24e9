#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sgm_parity_gpu.py -m gpu -q -x -k "bit_exact or host_api or subpixel" 2>&1 | tail -4
timeout 600 python - <<'PY' 2>&1 | tail -12
import json, torch, bench
dev = torch.device("cuda", 0)
print(json.dumps(bench.sgm_block(dev, 6585.4, 16)["tsgm_ragged"], indent=1))
PY
