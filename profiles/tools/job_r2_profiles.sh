#!/bin/bash
# Round-2 measurement job (one B200):  gpurun --timeout 2400 -- bash profiles/tools/job_r2_profiles.sh
# Everything lands in gpurun_out/; the summaries are copied to profiles/ by hand afterwards.
set -u
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r2_pytest_gpu.log 2>&1; tail -3 $O/r2_pytest_gpu.log
python bench.py --impl reference > $O/r2_bench_ref_n1.json 2> $O/r2_bench_ref_n1.err
python bench.py > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err; tail -c 300 $O/r2_bench_n1.err
# launch list of the bench command (per-launch times are cold-cache and serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-config5 > /dev/null 2>&1
# full captures of the dominant kernels at HEAD
ncu --set full --clock-control none --import-source on -k regex:shmp_fused_kernel -s 4 -c 1 -o $O/r2_fused -f \
    python bench.py --steps 2 --warmup 3 --no-gossip --no-config5 --eager-step > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:dense_tc_kernel -s 4 -c 1 -o $O/r2_dense_tc -f \
    python bench.py --steps 2 --warmup 3 --no-gossip --no-config5 --eager-step > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:shmp_mt_layer_kernel -s 9 -c 1 -o $O/r2_shmp_mt -f \
    python profiles/tools/config5_shmp.py --reps 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:partition_team_kernel -c 1 -o $O/r2_partition_team -f \
    python profiles/tools/config5_shmp.py --reps 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:partition_sparse_kernel -c 1 -o $O/r2_partition_sparse -f \
    python profiles/tools/config5_shmp.py --reps 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:gossip_layer0_kernel -c 1 -o $O/r2_gossip_layer0 -f \
    python bench.py --steps 3 --warmup 3 --no-config5 > /dev/null 2>&1
ls -la $O/*.ncu-rep | tail -8
