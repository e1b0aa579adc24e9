#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sgm_parity_gpu.py -m gpu -q -x -k "wide_image or pair" 2>&1 | tail -6
