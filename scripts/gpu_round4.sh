set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 1500 python bench.py --no-cpu > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "rc=$?" >> gpurun_out/bench_full.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_v80 -s 2000 -c 1 -o gpurun_out/prof_net2 python bench.py --steps 1 --warmup 2 --no-e2e --no-cpu > gpurun_out/ncu_net.log 2>&1
ls -la gpurun_out
