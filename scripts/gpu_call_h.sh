#!/bin/bash
# round-2 GPU check H (8 GPUs of one box): BASELINE configs[3] (C4) and configs[4] (C5) sharded over the ranks, C2 with the gather in e2e
mkdir -p gpurun_out
N=${1:-8}
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
echo "== C4 on $N GPUs"
run --workload c4 --steps 3 --warmup 3 2> gpurun_out/c4_n$N.err | tail -1 > gpurun_out/scale_c4_n$N.json; cut -c1-900 gpurun_out/scale_c4_n$N.json; tail -2 gpurun_out/c4_n$N.err
echo "== C5 on $N GPUs"
run --workload c5 --steps 2 --warmup 3 2> gpurun_out/c5_n$N.err | tail -1 > gpurun_out/scale_c5_n$N.json; cut -c1-1200 gpurun_out/scale_c5_n$N.json; tail -2 gpurun_out/c5_n$N.err
echo "== C2 on $N GPUs (weak scaling, gather inside e2e)"
run --steps 3 --warmup 3 --no-sgm 2> gpurun_out/c2_n$N.err | tail -1 > gpurun_out/scale_c2_n$N.json; cut -c1-700 gpurun_out/scale_c2_n$N.json; tail -2 gpurun_out/c2_n$N.err
