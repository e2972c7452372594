set -u
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/dist_check.py > gpurun_out/r01r_dist.log 2>&1; tail -6 gpurun_out/r01r_dist.log | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r01r_bench2.json 2> gpurun_out/r01r_bench2.err; cut -c1-900 gpurun_out/r01r_bench2.json; tail -3 gpurun_out/r01r_bench2.err | cut -c1-300
