# Round-2 ncu evidence: launch list (gpu__time_duration, no clock control) of a short Splendor bench run, and one --set full capture
# of each main kernel (k_select, k_v80_tc, k_backup ~600 launches into a search; k_v89_tc on the Santorini bench). The reports are
# summarised ON THE BOX (scripts/ncu_summary.py, ncu_source_hot.py) and deleted: only text comes back (gpurun_out is capped at 64 MiB).
# Numbers printed by bench.py under ncu are NOT bench values.
set -x
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-pcr --no-iteration"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 300 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/ncu_launch_run.log 2>&1
: > gpurun_out/r02_ncu_summary.txt
for K in k_select k_v80_tc k_backup; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 600 -c 1 -f -o /tmp/r02_$K $B > gpurun_out/ncu_$K.log 2>&1
  echo "==== $K (Splendor bench, launch ~600 of a search) ====" >> gpurun_out/r02_ncu_summary.txt
  python scripts/ncu_summary.py /tmp/r02_$K.ncu-rep >> gpurun_out/r02_ncu_summary.txt 2>&1
  python scripts/ncu_source_hot.py /tmp/r02_$K.ncu-rep >> gpurun_out/r02_ncu_summary.txt 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_v89_tc -s 200 -c 1 -f -o /tmp/r02_k_v89_tc python bench.py --game santorini --steps 1 --warmup 1 --no-e2e --no-cpu --no-pcr --no-iteration > gpurun_out/ncu_k_v89_tc.log 2>&1
echo "==== k_v89_tc (Santorini bench) ====" >> gpurun_out/r02_ncu_summary.txt
python scripts/ncu_summary.py /tmp/r02_k_v89_tc.ncu-rep >> gpurun_out/r02_ncu_summary.txt 2>&1
python scripts/ncu_source_hot.py /tmp/r02_k_v89_tc.ncu-rep >> gpurun_out/r02_ncu_summary.txt 2>&1
tail -5 gpurun_out/r02_ncu_summary.txt
for GK in "abalone k_v21_forward" "azul k_tokmix_forward"; do
  set -- $GK
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s 200 -c 1 -f -o /tmp/r02_$2 python bench.py --game $1 --steps 1 --warmup 1 --no-e2e --no-cpu --no-pcr --no-iteration > gpurun_out/ncu_$2.log 2>&1
  echo "==== $2 ($1 bench) ====" >> gpurun_out/r02_ncu_summary.txt
  python scripts/ncu_summary.py /tmp/r02_$2.ncu-rep >> gpurun_out/r02_ncu_summary.txt 2>&1
  python scripts/ncu_source_hot.py /tmp/r02_$2.ncu-rep >> gpurun_out/r02_ncu_summary.txt 2>&1
done
tail -5 gpurun_out/r02_ncu_summary.txt
