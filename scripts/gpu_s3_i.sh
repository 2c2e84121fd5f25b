set -x
mkdir -p gpurun_out
for nc in 1400 2560; do
timeout 900 python bench.py --no-cpu --no-e2e --steps 2 --node-cap $nc > gpurun_out/bench_nc$nc.json 2> gpurun_out/bench_nc$nc.err
done
timeout 900 python bench.py --no-cpu --no-e2e --steps 2 --games 4096 > gpurun_out/bench_g4096.json 2> gpurun_out/bench_g4096.err
