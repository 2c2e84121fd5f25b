set -x
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_mcts.py -m gpu -q -x -k "batched or deep or device_buffers" > gpurun_out/san_mcts.log 2>&1; echo "rc=$?" >> gpurun_out/san_mcts.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_net.py -m gpu -q -x -k "golden or ragged" > gpurun_out/san_net.log 2>&1; echo "rc=$?" >> gpurun_out/san_net.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_net.py -m gpu -q -x -k "golden" > gpurun_out/san_net_race.log 2>&1; echo "rc=$?" >> gpurun_out/san_net_race.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_abalone.py -m gpu -q -x -k "v21_forward" > gpurun_out/san_v21.log 2>&1; echo "rc=$?" >> gpurun_out/san_v21.log
