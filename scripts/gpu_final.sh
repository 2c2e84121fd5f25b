# Round-end evidence run: GPU tests, the three benches (default = BASELINE configs[2]), the reference arm, ncu launch list + full captures.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_splendor.json 2> gpurun_out/bench_splendor.err; echo "rc=$?" >> gpurun_out/bench_splendor.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_splendor_ref.json 2> gpurun_out/bench_splendor_ref.err
timeout 900 python bench.py --game santorini --steps 2 --warmup 3 > gpurun_out/bench_santorini.json 2> gpurun_out/bench_santorini.err
timeout 900 python bench.py --game abalone --steps 2 --warmup 3 > gpurun_out/bench_abalone.json 2> gpurun_out/bench_abalone.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_select -s 1500 -c 1 -o gpurun_out/prof_select_final python bench.py --steps 1 --warmup 2 --no-e2e --no-cpu > gpurun_out/ncu_select.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_v80_tc -s 1500 -c 1 -o gpurun_out/prof_net_final python bench.py --steps 1 --warmup 2 --no-e2e --no-cpu > gpurun_out/ncu_net.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_backup -s 1500 -c 1 -o gpurun_out/prof_backup_final python bench.py --steps 1 --warmup 2 --no-e2e --no-cpu > gpurun_out/ncu_backup.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
