set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_select -s 2000 -c 1 -o gpurun_out/prof_select_s3 python bench.py --steps 1 --warmup 2 --no-e2e --no-cpu > gpurun_out/ncu_select.log 2>&1
