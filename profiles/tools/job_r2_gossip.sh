set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gossip_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python profiles/tools/gossip_time.py > gpurun_out/r2h_gossip_time.json 2> gpurun_out/r2h_gossip_time.err
cat gpurun_out/r2h_gossip_time.json; tail -n 5 gpurun_out/r2h_gossip_time.err
