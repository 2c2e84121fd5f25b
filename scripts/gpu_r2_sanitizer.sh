# compute-sanitizer over the kernels that are new in round 2 (V89 tcgen05 trunk, token mixer, ragged self-play, node query, Azul plugin)
set -x
mkdir -p gpurun_out
: > gpurun_out/r02_compute_sanitizer.txt
run() { echo "==== compute-sanitizer --tool $1 : pytest -k \"$2\" ====" >> gpurun_out/r02_compute_sanitizer.txt; timeout 1200 compute-sanitizer --tool $1 --error-exitcode 9 python -m pytest tests -m gpu -q -x -k "$2" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Hazard|hazard|Invalid|Error" | head -20 >> gpurun_out/r02_compute_sanitizer.txt; }
run memcheck "v89_forward_golden or v84_forward or splendor_np and forward"
run memcheck "ragged or nodes_data or reference_examples and azul"
run racecheck "v80_forward_golden"
run racecheck "v89_forward_golden"
run racecheck "v84_forward"
cat gpurun_out/r02_compute_sanitizer.txt
