set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r2b_pytest_gpu.log
cat gpurun_out/r2b_pytest_gpu.log
timeout 300 python profiles/tools/fused_phase_profile.py > gpurun_out/r2b_fused_phase.txt 2>&1
cat gpurun_out/r2b_fused_phase.txt
