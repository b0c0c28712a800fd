set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_shmp_gpu.py tests/test_workload_gpu.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_shmp.log
cat gpurun_out/pytest_shmp.log
timeout 600 python bench.py --no-gossip > gpurun_out/bench_nog.json 2> gpurun_out/bench_nog.err
cat gpurun_out/bench_nog.json; tail -n 3 gpurun_out/bench_nog.err
timeout 300 python profiles/tools/fused_phase_profile.py > gpurun_out/fused_phase.txt 2>&1; cat gpurun_out/fused_phase.txt
