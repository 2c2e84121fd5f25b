set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_abalone.py -m gpu -x -q > gpurun_out/pytest_aba.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_aba.log
timeout 900 python bench.py --game abalone --steps 2 --warmup 3 > gpurun_out/bench_abalone.json 2> gpurun_out/bench_abalone.err; echo "rc=$?" >> gpurun_out/bench_abalone.err
