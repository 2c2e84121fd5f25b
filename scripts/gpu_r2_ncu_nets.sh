# ncu --set full of the two tensor-core net kernels only (after their last changes); summaries -> gpurun_out/r02_ncu_nets.txt
set -x
mkdir -p gpurun_out
: > gpurun_out/r02_ncu_nets.txt
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-pcr --no-iteration"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_v80_tc -s 600 -c 1 -f -o /tmp/r02_k_v80_tc $B > gpurun_out/ncu_k_v80_tc.log 2>&1
echo "==== k_v80_tc (Splendor bench, launch ~600 of a search) ====" >> gpurun_out/r02_ncu_nets.txt
python scripts/ncu_summary.py /tmp/r02_k_v80_tc.ncu-rep >> gpurun_out/r02_ncu_nets.txt 2>&1
python scripts/ncu_source_hot.py /tmp/r02_k_v80_tc.ncu-rep >> gpurun_out/r02_ncu_nets.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_v89_tc -s 200 -c 1 -f -o /tmp/r02_k_v89_tc python bench.py --game santorini --steps 1 --warmup 1 --no-e2e --no-cpu --no-pcr --no-iteration > gpurun_out/ncu_k_v89_tc.log 2>&1
echo "==== k_v89_tc (Santorini bench) ====" >> gpurun_out/r02_ncu_nets.txt
python scripts/ncu_summary.py /tmp/r02_k_v89_tc.ncu-rep >> gpurun_out/r02_ncu_nets.txt 2>&1
python scripts/ncu_source_hot.py /tmp/r02_k_v89_tc.ncu-rep >> gpurun_out/r02_ncu_nets.txt 2>&1
grep -n "^====\|gpu__time_duration\|tensor_cycles_active.avg" gpurun_out/r02_ncu_nets.txt
