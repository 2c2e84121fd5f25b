set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_splendor.json 2> gpurun_out/bench_splendor.err; echo "rc=$?" >> gpurun_out/bench_splendor.err
tail -3 gpurun_out/bench_splendor.err
