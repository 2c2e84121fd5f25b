set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --game abalone --no-cpu --no-e2e --steps 2 --warmup 3 > gpurun_out/bench_abalone_x.json 2> gpurun_out/bench_abalone_x.err
timeout 900 python bench.py --no-cpu --no-e2e --steps 2 > gpurun_out/bench_splendor_x.json 2> gpurun_out/bench_splendor_x.err
