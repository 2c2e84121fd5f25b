# Phase stamps of one tile of the tcgen05 V80 kernel (clock64 deltas of CTA 0; the script sets AZG_V80_PROF=1, no special build needed)
set -x
mkdir -p gpurun_out
timeout 300 python scripts/dbg_v80tc.py > gpurun_out/dbg_v80tc.log 2>&1; echo "rc=$?" >> gpurun_out/dbg_v80tc.log
