#!/bin/bash
# Round-2 record run (one B200):  gpurun --timeout 2400 -- bash profiles/tools/job_r2_final.sh
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv
timeout 1500 python -m pytest tests -m gpu -q > $O/r2f_pytest_gpu.log 2>&1; tail -3 $O/r2f_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2f_smoke.log 2>&1; tail -4 $O/r2f_smoke.log
timeout 900 python bench.py > $O/r2f_bench_n1.json 2> $O/r2f_bench_n1.err; tail -c 300 $O/r2f_bench_n1.err
timeout 600 python bench.py --impl reference > $O/r2f_bench_ref_n1.json 2> $O/r2f_bench_ref_n1.err
timeout 300 python profiles/tools/fused_phase_profile.py > $O/r2f_fused_phase.txt 2>&1
timeout 600 python profiles/tools/configs_small.py > $O/r2f_configs_small.jsonl 2> $O/r2f_configs_small.err
timeout 600 python profiles/tools/config3_train.py > $O/r2f_config3_train.json 2> $O/r2f_config3_train.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2f_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-config5 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:shmp_fused_kernel -s 4 -c 1 -o $O/r2f_fused -f \
    python bench.py --steps 2 --warmup 3 --no-gossip --no-config5 --eager-step > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gossip_gather_kernel -s 6 -c 1 -o $O/r2f_gossip_gather -f \
    python bench.py --steps 1 --warmup 3 --no-config5 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gossip_chain_kernel -s 6 -c 1 -o $O/r2f_gossip_chain -f \
    python bench.py --steps 1 --warmup 3 --no-config5 > /dev/null 2>&1
cat $O/r2f_fused_phase.txt; head -c 600 $O/r2f_bench_n1.json; tail -n 3 $O/*.err | tail -30
