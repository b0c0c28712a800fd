set -x
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_config5_1m_4chunks.csv python profiles/tools/config5.py --nodes 1000000 --edges 10000000 --gossip-steps 1 --no-shmp > gpurun_out/c5_ncu.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_config5_10m_2chunks.csv python profiles/tools/config5.py --chunks 2 --gossip-steps 1 --no-shmp > gpurun_out/c5_ncu10.log 2>&1
