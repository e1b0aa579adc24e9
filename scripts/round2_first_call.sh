#!/bin/bash
# One gpurun call that measures everything round 1 prepared but could not run (GPU budget spent):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash scripts/round2_first_call.sh'
# Outputs land in gpurun_out/ (copy what is worth keeping to profiles/).
mkdir -p gpurun_out
echo "== experimental variant checks (DPX step, concurrent directions, packed taps in geometric passes, toGray)"
B200MVS_TEST_EXPERIMENTAL=1 timeout 240 python -m pytest tests/test_experimental_gpu.py -m gpu -q 2>&1 | tail -8
echo "== SGM aggregation variants: time + identity"
timeout 120 python scripts/profile_sgm.py 128 2>&1 | tail -6 | tee gpurun_out/sgm_variants.txt
echo "== sweep kernel variants: time + identity"
timeout 120 python scripts/profile_variants.py 3 2>&1 | tail -6 | tee gpurun_out/pm_variants.txt
echo "== ncu --set full: packed sweep kernel (converged launch) and one ring aggregation launch per direction class"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:pm_sweep -s 30 -c 1 -o gpurun_out/pm_sweep_packed -f python scripts/profile_sweep.py 6 > gpurun_out/ncu_pm.log 2>&1
timeout 60 python scripts/ncu_summary.py gpurun_out/pm_sweep_packed.ncu-rep > gpurun_out/ncu_pm_sweep_packed.txt 2>&1; head -12 gpurun_out/ncu_pm_sweep_packed.txt
timeout 200 ncu --set full --clock-control none --import-source on -k regex:sgm_aggregate_uniform_ring -s 8 -c 2 -o gpurun_out/sgm_ring -f python scripts/profile_sgm.py 128 default > gpurun_out/ncu_sgm.log 2>&1
timeout 60 python scripts/ncu_summary.py gpurun_out/sgm_ring.ncu-rep 0 > gpurun_out/ncu_sgm_ring_down.txt 2>&1
timeout 60 python scripts/ncu_summary.py gpurun_out/sgm_ring.ncu-rep 1 > gpurun_out/ncu_sgm_ring_right.txt 2>&1; head -12 gpurun_out/ncu_sgm_ring_right.txt
