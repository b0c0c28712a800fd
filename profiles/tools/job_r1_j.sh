set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-gossip > gpurun_out/bench_nog.json 2> gpurun_out/bench_nog.err
cat gpurun_out/bench_nog.json; tail -n 3 gpurun_out/bench_nog.err
timeout 300 python profiles/tools/host_overhead.py > gpurun_out/host.txt 2>&1; cat gpurun_out/host.txt
