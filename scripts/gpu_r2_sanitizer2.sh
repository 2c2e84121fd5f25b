# compute-sanitizer memcheck over the net kernels as they stand at the end of round 2 (V80 / V89 tcgen05, V21, token mixer) and a short
# self-play + arena run; appended to gpurun_out/r02_compute_sanitizer.txt
set -x
mkdir -p gpurun_out
run() { echo "==== compute-sanitizer --tool $1 : pytest -k \"$2\" ====" >> gpurun_out/r02_compute_sanitizer.txt; timeout 1500 compute-sanitizer --tool $1 --error-exitcode 9 python -m pytest tests -m gpu -q -x -k "$2" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Hazard|hazard|Invalid|Error" | head -20 >> gpurun_out/r02_compute_sanitizer.txt; }
: > gpurun_out/r02_compute_sanitizer.txt
run memcheck "v80_forward or v80_tensor_core or v89_forward_golden or v21_forward or v84_forward or splendor_np and forward"
run memcheck "test_main_runs_one_iteration or selfplay_examples_match_reference and splendor"
cat gpurun_out/r02_compute_sanitizer.txt
