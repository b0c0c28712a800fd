set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cat gpurun_out/bench_ref.json
timeout 300 python profiles/tools/fused_phase_profile.py > gpurun_out/fused_phase.txt 2>&1; cat gpurun_out/fused_phase.txt
timeout 300 python profiles/tools/host_overhead.py > gpurun_out/host.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 5 --warmup 3 --no-gossip > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:shmp_fused_kernel -c 1 -s 3 -o gpurun_out/fused_v7 -f python bench.py --no-gossip --steps 3 --warmup 1 > gpurun_out/ncu_fused.log 2>&1
tail -n 2 gpurun_out/ncu_fused.log
timeout 600 python profiles/tools/config3_train.py > gpurun_out/config3_train.json 2> gpurun_out/config3_train.err; cat gpurun_out/config3_train.json
