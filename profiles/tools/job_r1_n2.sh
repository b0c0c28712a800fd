set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
cat gpurun_out/bench_n2.json; tail -n 3 gpurun_out/bench_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 profiles/tools/config5.py --chunks 2 --gossip-steps 2 > gpurun_out/config5_10m_n2.json 2> gpurun_out/config5_10m_n2.err
cat gpurun_out/config5_10m_n2.json; tail -n 3 gpurun_out/config5_10m_n2.err
