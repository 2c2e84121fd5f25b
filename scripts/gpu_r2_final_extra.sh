set -x
mkdir -p gpurun_out
timeout 900 python bench.py --game santorini > gpurun_out/r02_bench_santorini_4096x800.json 2> gpurun_out/bench_santorini.err
timeout 900 python bench.py --game azul > gpurun_out/r02_bench_azul_8192x800.json 2> gpurun_out/bench_azul.err
timeout 900 env AZG_V89_KERNEL=fp32 python bench.py --game santorini --steps 3 --warmup 3 --no-e2e --no-cpu --no-pcr --no-iteration > gpurun_out/r02_bench_santorini_fp32_net_kernel.json 2>> gpurun_out/bench_santorini.err
wc -c gpurun_out/r02_bench_*.json
