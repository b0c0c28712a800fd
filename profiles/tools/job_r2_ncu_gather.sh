mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gossip_gather_kernel -s 4 -c 1 -o gpurun_out/r2h_gossip_gather -f python profiles/tools/gossip_time.py --steps 1 --no-fp32 > gpurun_out/r2h_ncu.log 2>&1
tail -n 3 gpurun_out/r2h_ncu.log
