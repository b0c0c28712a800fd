set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()"
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 300 python profiles/tools/fused_phase_profile.py > gpurun_out/fused_phase.txt 2>&1
timeout 300 python profiles/tools/gossip_phase_profile.py > gpurun_out/gossip_phase.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_launches.log 2>&1
timeout 600 python profiles/tools/config3_train.py > gpurun_out/config3_train.json 2> gpurun_out/config3_train.err
timeout 600 python profiles/tools/config5.py --chunks 2 --gossip-steps 2 > gpurun_out/config5_10m_n1.json 2> gpurun_out/config5_10m_n1.err; cat gpurun_out/config5_10m_n1.json
timeout 600 python profiles/tools/config5.py --nodes 1000000 --edges 10000000 --gossip-steps 2 > gpurun_out/config5_1m.json 2> gpurun_out/config5_1m.err
timeout 600 python profiles/tools/configs_small.py > gpurun_out/configs_small.jsonl 2> gpurun_out/configs_small.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:shmp_fused_kernel -c 1 -s 3 -o gpurun_out/fused_v10 -f python bench.py --no-gossip --steps 3 --warmup 1 > gpurun_out/ncu_fused.log 2>&1
