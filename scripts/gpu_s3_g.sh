set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/n2_gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?" >> gpurun_out/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_n2_ref.json 2> gpurun_out/bench_n2_ref.err; echo "rc=$?" >> gpurun_out/bench_n2_ref.err
