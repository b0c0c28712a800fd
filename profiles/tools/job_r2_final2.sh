#!/bin/bash
# Round-2 record run after the tensor-path gossip gather (one B200):  gpurun --timeout 2400 -- bash profiles/tools/job_r2_final2.sh
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv
timeout 1500 python -m pytest tests -m gpu -q > $O/r2i_pytest_gpu.log 2>&1; tail -3 $O/r2i_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2i_smoke.log 2>&1; tail -4 $O/r2i_smoke.log
timeout 900 python bench.py > $O/r2i_bench_n1.json 2> $O/r2i_bench_n1.err; tail -c 300 $O/r2i_bench_n1.err
timeout 600 python bench.py --impl reference > $O/r2i_bench_ref_n1.json 2> $O/r2i_bench_ref_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2i_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-config5 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gossip_gather_kernel -s 6 -c 1 -o $O/r2i_gossip_gather -f \
    python bench.py --steps 1 --warmup 3 --no-config5 > /dev/null 2>&1
head -c 400 $O/r2i_bench_n1.json; tail -n 3 $O/*.err | tail -20
