set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
python profiles/tools/fused_phase_profile.py > gpurun_out/phase.txt 2>&1
python profiles/tools/host_overhead.py > gpurun_out/host.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:shmp_fused -s 3 -c 1 -o gpurun_out/prof_fused python bench.py --steps 2 --warmup 3 > gpurun_out/b_ncu2.log 2>&1
cat gpurun_out/pytest_gpu.log gpurun_out/bench.json gpurun_out/phase.txt gpurun_out/host.txt
