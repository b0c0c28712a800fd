set -x
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:partition_small_kernel -c 1 -s 5 -o gpurun_out/partition_small_v9 -f python bench.py --no-gossip --steps 3 --warmup 3 > gpurun_out/ncu_p1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:dense_tc_kernel -c 1 -s 3 -o gpurun_out/dense_tc_v9 -f python bench.py --no-gossip --steps 3 --warmup 3 > gpurun_out/ncu_p2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:partition_sparse_kernel -c 1 -s 3 -o gpurun_out/partition_sparse_v9 -f python profiles/tools/config5.py --nodes 1000000 --edges 10000000 --chunks 1 --gossip-steps 1 --no-shmp > gpurun_out/ncu_p3.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:partition_team_kernel -c 1 -s 3 -o gpurun_out/partition_team_v9 -f python profiles/tools/config5.py --nodes 1000000 --edges 10000000 --chunks 1 --gossip-steps 1 --no-shmp > gpurun_out/ncu_p4.log 2>&1
tail -n 2 gpurun_out/ncu_p*.log
