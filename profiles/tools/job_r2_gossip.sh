set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gossip_gpu.py tests/test_gossip_train_gpu.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 5 --warmup 3 --no-config5 > gpurun_out/r2e_bench_g.json 2> gpurun_out/r2e_bench_g.err
python - <<'PY'
import json
for l in open('gpurun_out/r2e_bench_g.json'):
    if l.startswith('{'):
        d=json.loads(l); g=d['gossip']; print(g['value'], g['ms_per_step'], g['stage_ms'])
PY
tail -n 3 gpurun_out/r2e_bench_g.err
