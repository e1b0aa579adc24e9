#!/bin/bash
# round-2 final GPU check: the whole -m gpu suite, smoke(), sanitizer, the bench line, SGM traffic and ncu evidence of the shipped kernels
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -12
echo "== smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== compute-sanitizer memcheck (PatchMatch single/multi scale/geometric, SGM ragged/ring/fronts/tensor-core cost, post-processing)"
timeout 1200 compute-sanitizer --tool memcheck --print-limit 5 python scripts/sanitize_small.py 2>&1 | tail -14 | tee gpurun_out/sanitizer_all.txt
echo "== bench"
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 1500 gpurun_out/bench_final.json; tail -3 gpurun_out/bench_final.err
echo "== reference arm (bounded CPU sample)"
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 700 gpurun_out/bench_reference.json
echo "== ncu: one SGM Match, every kernel: time and DRAM bytes; D = 64 / 256 timings"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:sgm -c 40 --csv --log-file gpurun_out/launches_sgm.csv python scripts/profile_sgm.py 128 default > gpurun_out/ncu_sgm.log 2>&1; grep "sgm_front\|wta\|cost_tc" gpurun_out/launches_sgm.csv | tail -9 | awk -F'","' '{print substr($5,1,40), $(NF-2), $(NF)}'
timeout 200 python scripts/profile_sgm.py 128 default 2>&1 | tail -1 | tee gpurun_out/sgm_defaults.txt
timeout 200 python scripts/profile_sgm.py 64 default 2>&1 | tail -1 | tee -a gpurun_out/sgm_defaults.txt
timeout 200 python scripts/profile_sgm.py 256 default 2>&1 | tail -1 | tee -a gpurun_out/sgm_defaults.txt
echo "== ncu --set full: wave-front kernel (shipped defaults)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sgm_front_kernel -s 1 -c 1 -o gpurun_out/sgm_front -f python scripts/profile_sgm.py 128 default > gpurun_out/ncu_front.log 2>&1
timeout 60 python scripts/ncu_summary.py gpurun_out/sgm_front.ncu-rep 0 > gpurun_out/ncu_sgm_front.txt 2>&1; head -8 gpurun_out/ncu_sgm_front.txt
