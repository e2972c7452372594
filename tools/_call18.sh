set -u
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 tests/dist_check.py > gpurun_out/r01s_dist4.log 2>&1; tail -5 gpurun_out/r01s_dist4.log | cut -c1-250
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r01s_bench4.json 2> gpurun_out/r01s_bench4.err; cut -c1-260 gpurun_out/r01s_bench4.json
