set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 python profiles/tools/config3_train.py > gpurun_out/config3_train.json 2> gpurun_out/config3_train.err
timeout 600 python profiles/tools/config5.py --nodes 1000000 --edges 10000000 > gpurun_out/config5_1m.json 2> gpurun_out/config5_1m.err
timeout 500 python profiles/tools/config5.py --chunk 1024 --chunks 1 --gossip-steps 2 > gpurun_out/config5_10m_n1.json 2> gpurun_out/config5_10m_n1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_config5_1m.csv python profiles/tools/config5.py --nodes 1000000 --edges 10000000 --chunks 1 --gossip-steps 1 --no-shmp > gpurun_out/c5_ncu.log 2>&1
cp MEASURED_PEAKS.json gpurun_out/ 2>/dev/null
cat gpurun_out/pytest_gpu.log gpurun_out/bench.json gpurun_out/config3_train.json gpurun_out/config5_1m.json gpurun_out/config5_10m_n1.json; tail -5 gpurun_out/*.err
