set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_select -s 1500 -c 1 -o gpurun_out/prof_select_s3c python bench.py --steps 1 --warmup 2 --no-e2e --no-cpu > gpurun_out/ncu_select.log 2>&1
