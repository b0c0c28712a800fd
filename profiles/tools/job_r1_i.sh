set -x
mkdir -p gpurun_out
timeout 600 python bench.py --no-gossip > gpurun_out/bench_nog.json 2> gpurun_out/bench_nog.err
cat gpurun_out/bench_nog.json; tail -n 3 gpurun_out/bench_nog.err
timeout 300 python profiles/tools/fused_phase_profile.py > gpurun_out/fused_phase.txt 2>&1; cat gpurun_out/fused_phase.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:shmp_fused_kernel -c 1 -s 3 -o gpurun_out/fused_v6 -f python bench.py --no-gossip --steps 3 --warmup 1 > gpurun_out/ncu_fused.log 2>&1
tail -n 3 gpurun_out/ncu_fused.log
