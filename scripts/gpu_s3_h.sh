set -x
mkdir -p gpurun_out
timeout 900 python bench.py --no-cpu --no-e2e --steps 2 > gpurun_out/bench_splendor_x.json 2> gpurun_out/bench_splendor_x.err; echo "rc=$?" >> gpurun_out/bench_splendor_x.err
