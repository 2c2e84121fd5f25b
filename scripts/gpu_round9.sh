set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --game abalone --steps 2 --warmup 2 > gpurun_out/bench_abalone.json 2> gpurun_out/bench_abalone.err; echo "rc=$?" >> gpurun_out/bench_abalone.err
