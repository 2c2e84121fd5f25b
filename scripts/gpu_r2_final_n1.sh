# Round-2 evidence, 1 GPU: GPU tests, smoke, the four bench lines (default K / W), the reference arm, the launch list with the final kernels.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/r02_bench_splendor_16384x800.json 2> gpurun_out/bench_splendor.err
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r02_bench_splendor_reference_arm.json 2>> gpurun_out/bench_splendor.err
timeout 900 python bench.py --game santorini > gpurun_out/r02_bench_santorini_4096x800.json 2> gpurun_out/bench_santorini.err
timeout 900 python bench.py --game abalone > gpurun_out/r02_bench_abalone_2048x1600.json 2> gpurun_out/bench_abalone.err
timeout 900 python bench.py --game azul > gpurun_out/r02_bench_azul_8192x800.json 2> gpurun_out/bench_azul.err
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-pcr --no-iteration"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 300 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/ncu_launch_run.log 2>&1
tail -3 gpurun_out/pytest_gpu.log gpurun_out/smoke.log; wc -c gpurun_out/r02_bench_*.json
