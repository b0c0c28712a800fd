set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_shmp_gpu.py tests/test_pipeline_gpu.py tests/test_workload_gpu.py -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r2c_pytest_shmp.log
cat gpurun_out/r2c_pytest_shmp.log
timeout 300 python profiles/tools/fused_phase_profile.py > gpurun_out/r2c_fused_phase.txt 2>&1
cat gpurun_out/r2c_fused_phase.txt
timeout 600 python bench.py --no-gossip --no-config5 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
python - <<'PY'
import json
for l in open('gpurun_out/r2c_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['stage_ms_per_step'], d['roofline']['frac'], d['parity'])
PY
tail -n 5 gpurun_out/r2c_bench.err
