set -x
mkdir -p gpurun_out
AZG_V80_KERNEL=fp32 timeout 900 python bench.py --no-cpu --steps 3 > gpurun_out/bench_splendor_fp32net.json 2> gpurun_out/bench_splendor_fp32net.err
AZG_TREE_REPLAY=0 timeout 900 python bench.py --no-cpu --no-e2e --steps 3 > gpurun_out/bench_splendor_noreplay.json 2> gpurun_out/bench_splendor_noreplay.err
