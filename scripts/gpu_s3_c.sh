set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_splendor.json 2> gpurun_out/bench_splendor.err; echo "rc=$?" >> gpurun_out/bench_splendor.err
AZG_V80_KERNEL=fp32 timeout 900 python bench.py --no-e2e --no-cpu --steps 2 > gpurun_out/bench_splendor_fp32net.json 2> gpurun_out/bench_splendor_fp32net.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_v80_tc -s 2000 -c 1 -o gpurun_out/prof_v80tc python bench.py --steps 1 --warmup 2 --no-e2e --no-cpu > gpurun_out/ncu_v80tc.log 2>&1
