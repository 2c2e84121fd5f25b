set -x
mkdir -p gpurun_out
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r02_bench_splendor_n2.json 2> gpurun_out/bench_splendor_n2.err; echo "rc=$?" >> gpurun_out/bench_splendor_n2.err
tail -3 gpurun_out/bench_splendor_n2.err
