set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gossip_gpu.py tests/test_workload_gpu.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gossip.log
cat gpurun_out/pytest_gossip.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err
grep -o '"gossip": {.*' gpurun_out/bench_g.json | cut -c1-900; tail -n 3 gpurun_out/bench_g.err
timeout 300 python profiles/tools/gossip_phase_profile.py 2>&1 | tail -8
