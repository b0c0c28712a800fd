set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
timeout 600 python profiles/tools/config5.py --nodes 1000000 --edges 10000000 > gpurun_out/config5_1m.json 2> gpurun_out/config5_1m.err
timeout 900 python profiles/tools/config5.py > gpurun_out/config5_10m_n1.json 2> gpurun_out/config5_10m_n1.err
cat gpurun_out/pytest_gpu.log gpurun_out/config5_1m.json gpurun_out/config5_10m_n1.json; tail -5 gpurun_out/config5_1m.err gpurun_out/config5_10m_n1.err
