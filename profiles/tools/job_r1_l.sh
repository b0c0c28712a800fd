set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_config3.csv python profiles/tools/config3_train.py --steps 6 --warmup 1 --cpu-steps 0 > gpurun_out/c3_ncu.log 2>&1
tail -n 2 gpurun_out/c3_ncu.log
