#!/bin/bash
# gpurun --gpus N --timeout 900 -- bash profiles/tools/job_r2_scale_probe.sh N
N=${1:-4}
O=gpurun_out
mkdir -p $O
run() {  # tag, extra args...; NCCL_* from the environment
  tag=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      profiles/tools/gossip_scale_probe.py --tag $tag "$@" 2> $O/probe_$tag.err | grep '^{' | tee -a $O/r2h_scale_probe_n$N.jsonl
  tail -n 2 $O/probe_$tag.err | cut -c1-200
}
run default
NCCL_MAX_CTAS=4 run maxctas4
NCCL_MAX_CTAS=4 run maxctas4_qg8 --query-group 8
