set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_v21 -s 200 -c 1 -f -o gpurun_out/prof_v21 python bench.py --game abalone --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_v21.log 2>&1
