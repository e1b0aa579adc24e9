#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/profile_sgm_pair.py 128 2>&1 | tail -4 | tee gpurun_out/sgm_pair.txt
timeout 300 python scripts/profile_sgm_pair.py 64 2>&1 | tail -4 | tee -a gpurun_out/sgm_pair.txt
timeout 300 python -m pytest tests/test_sgm_parity_gpu.py -m gpu -q -k pair 2>&1 | tail -3
