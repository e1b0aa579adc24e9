#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/profile_sweep.py 6 2 1 0 0 0 2>&1 | tail -2 | tee gpurun_out/pm_block.txt
timeout 900 python -m pytest tests/test_pm_parity_gpu.py -m gpu -q -x -k "full_estimate or bench_configuration or multi_scale or geometric_consistency or textureless or strided" 2>&1 | tail -4
