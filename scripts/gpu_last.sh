set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_splendor.json 2> gpurun_out/bench_splendor.err; echo "rc=$?" >> gpurun_out/bench_splendor.err
timeout 900 python bench.py --game santorini --no-cpu --no-e2e --steps 2 --warmup 3 > gpurun_out/bench_santorini_x.json 2> gpurun_out/bench_santorini_x.err
