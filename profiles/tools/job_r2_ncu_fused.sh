set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:shmp_fused_kernel -c 1 -s 3 -o gpurun_out/r2d_fused -f python bench.py --no-gossip --no-config5 --steps 3 --warmup 3 > gpurun_out/r2d_ncu_fused.log 2>&1
tail -n 3 gpurun_out/r2d_ncu_fused.log
