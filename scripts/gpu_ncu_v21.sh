set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_v21 -s 200 -c 1 -f -o gpurun_out/prof_v21_final python bench.py --game abalone --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_v21b.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 300 --csv --log-file gpurun_out/launches_abalone.csv python bench.py --game abalone --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_launch_aba.log 2>&1
