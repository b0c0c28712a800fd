set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 python profiles/tools/fused_phase_profile.py > gpurun_out/phase.txt 2>&1
timeout 300 python profiles/tools/host_overhead.py > gpurun_out/host.txt 2>&1
cat gpurun_out/pytest_gpu.log gpurun_out/bench.json gpurun_out/phase.txt gpurun_out/host.txt; tail -5 gpurun_out/bench.err
