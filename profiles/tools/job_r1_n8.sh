set -x
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 profiles/tools/config5.py --chunks 2 --gossip-steps 2 > gpurun_out/config5_10m_n$N.json 2> gpurun_out/config5_10m_n$N.err
cat gpurun_out/config5_10m_n$N.json; tail -n 3 gpurun_out/config5_10m_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
cat gpurun_out/bench_n$N.json | cut -c1-400; tail -n 3 gpurun_out/bench_n$N.err
