set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -k "${K:-azul or arena}" > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_quick.log
tail -40 gpurun_out/pytest_quick.log
