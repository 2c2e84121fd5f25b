set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_splendor.json 2> gpurun_out/bench_splendor.err; echo "rc=$?" >> gpurun_out/bench_splendor.err
tail -5 gpurun_out/pytest_gpu.log
