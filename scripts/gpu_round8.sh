set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 2 --no-cpu > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "rc=$?" >> gpurun_out/bench_full.err
