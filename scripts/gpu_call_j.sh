#!/bin/bash
# round-2 GPU check J: wave-front queue ordered along the front in every phase; lag / sub-cell width grid
mkdir -p gpurun_out
echo "== front tests"
timeout 900 python -m pytest tests/test_sgm_parity_gpu.py -m gpu -x -q -k "wave_front or variants" 2>&1 | tail -4
echo "== grid"
timeout 600 python scripts/profile_sgm.py 128 order 2>&1 | tail -40 | tee gpurun_out/sgm_order.txt
echo "== smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
