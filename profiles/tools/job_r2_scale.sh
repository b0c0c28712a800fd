#!/bin/bash
# Round-2 multi-GPU record:  gpurun --gpus N --timeout 900 -- bash profiles/tools/job_r2_scale.sh N
set -u
N=${1:-8}
O=gpurun_out
mkdir -p $O
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 \
    bench.py --gpus $N --steps 10 --warmup 3 > $O/r2i_bench_n$N.json 2> $O/r2i_bench_n$N.err
cut -c1-300 $O/r2i_bench_n$N.json; tail -n 3 $O/r2i_bench_n$N.err
