#!/bin/bash
# round-2 GPU check F: the whole -m gpu suite on the final kernels, sanitizer on the new SGM kernels, bench line, ncu evidence
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
echo "== compute-sanitizer memcheck (SGM: ring, wave fronts, tensor-core cost)"
SANITIZE_ONLY=sgm timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python scripts/sanitize_small.py 2>&1 | tail -12 | tee gpurun_out/sanitizer_sgm.txt
SANITIZE_ONLY=sgm timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python scripts/sanitize_small.py 2>&1 | tail -8 | tee -a gpurun_out/sanitizer_sgm.txt
echo "== sweep kernel: 3 vs 4 CTAs per SM"
timeout 200 python scripts/profile_sweep.py 6 2 1 0 0 0 2>&1 | tail -2 | tee gpurun_out/pm_ctas.txt
timeout 200 python scripts/profile_sweep.py 6 2 1 0 1 0 2>&1 | tail -2 | tee -a gpurun_out/pm_ctas.txt
echo "== bench"
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err; tail -c 3000 gpurun_out/bench_f.json; tail -3 gpurun_out/bench_f.err
echo "== ncu launch list of the bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | cut -c1-300
echo "== ncu: one SGM Match, every kernel: time and DRAM bytes"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:sgm -c 40 --csv --log-file gpurun_out/launches_sgm.csv python scripts/profile_sgm.py 128 default > gpurun_out/ncu_sgm.log 2>&1; tail -12 gpurun_out/launches_sgm.csv | cut -c1-260
echo "== ncu --set full: shipped sweep kernel (launch 12), wave-front kernel, tensor-core cost kernel"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pm_sweep -s 12 -c 1 -o gpurun_out/pm_sweep -f python scripts/profile_sweep.py 6 > gpurun_out/ncu_pm.log 2>&1
timeout 60 python scripts/ncu_summary.py gpurun_out/pm_sweep.ncu-rep > gpurun_out/ncu_pm_sweep.txt 2>&1; head -12 gpurun_out/ncu_pm_sweep.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sgm_front_kernel -s 1 -c 1 -o gpurun_out/sgm_front -f python scripts/profile_sgm.py 128 default > gpurun_out/ncu_front.log 2>&1
timeout 60 python scripts/ncu_summary.py gpurun_out/sgm_front.ncu-rep 0 > gpurun_out/ncu_sgm_front.txt 2>&1; head -8 gpurun_out/ncu_sgm_front.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sgm_cost_tc_kernel -s 1 -c 1 -o gpurun_out/sgm_cost_tc -f python scripts/profile_sgm.py 128 default > gpurun_out/ncu_tc.log 2>&1
timeout 60 python scripts/ncu_summary.py gpurun_out/sgm_cost_tc.ncu-rep > gpurun_out/ncu_sgm_cost_tc.txt 2>&1; head -8 gpurun_out/ncu_sgm_cost_tc.txt
