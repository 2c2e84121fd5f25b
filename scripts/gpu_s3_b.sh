set -x
mkdir -p gpurun_out
timeout 300 python scripts/dbg_v80tc.py > gpurun_out/dbg_v80tc.log 2>&1; echo "rc=$?" >> gpurun_out/dbg_v80tc.log
