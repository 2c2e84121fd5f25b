set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py --games 2048 --sims 200 --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "rc=$?" >> gpurun_out/bench_small.err
timeout 1500 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "rc=$?" >> gpurun_out/bench_full.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_select -s 1000 -c 2 -o gpurun_out/prof_select python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_select.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_v80 -s 1000 -c 2 -o gpurun_out/prof_net python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_net.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_backup -s 1000 -c 2 -o gpurun_out/prof_backup python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_backup.log 2>&1
ls -la gpurun_out
