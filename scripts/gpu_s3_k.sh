set -x
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_select -s 1500 -c 1 -o gpurun_out/prof_select_final python bench.py --steps 1 --warmup 2 --no-e2e --no-cpu > gpurun_out/ncu_select.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_v80_tc -s 1500 -c 1 -o gpurun_out/prof_net_final python bench.py --steps 1 --warmup 2 --no-e2e --no-cpu > gpurun_out/ncu_net.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_backup -s 1500 -c 1 -o gpurun_out/prof_backup_final python bench.py --steps 1 --warmup 2 --no-e2e --no-cpu > gpurun_out/ncu_backup.log 2>&1
