#!/bin/bash
# round-2 GPU check L: sweep kernel with the conversion-free floor
mkdir -p gpurun_out
timeout 200 python scripts/profile_sweep.py 6 2 1 0 0 0 2>&1 | tail -2 | tee gpurun_out/pm_floor.txt
timeout 1500 python -m pytest tests/test_pm_parity_gpu.py tests/test_real_fixture.py tests/test_cpp_adapter.py -m gpu -q -x 2>&1 | tail -8
