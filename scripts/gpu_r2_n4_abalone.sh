set -x
mkdir -p gpurun_out
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --game abalone > gpurun_out/r02_bench_abalone_n4.json 2> gpurun_out/bench_abalone_n4.err; echo "rc=$?" >> gpurun_out/bench_abalone_n4.err
tail -3 gpurun_out/bench_abalone_n4.err
