set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_partition_gpu.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_part.log
cat gpurun_out/pytest_part.log
timeout 600 python profiles/tools/config5.py --nodes 1000000 --edges 10000000 --gossip-steps 1 --no-shmp > gpurun_out/config5_1m.json 2> gpurun_out/config5_1m.err
timeout 500 python profiles/tools/config5.py --chunks 2 --gossip-steps 1 --no-shmp > gpurun_out/config5_10m_n1.json 2> gpurun_out/config5_10m_n1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_config5_1m.csv python profiles/tools/config5.py --nodes 1000000 --edges 10000000 --chunks 1 --gossip-steps 1 --no-shmp > gpurun_out/c5_ncu.log 2>&1
cat gpurun_out/config5_1m.json gpurun_out/config5_10m_n1.json; tail -n 5 gpurun_out/*.err
