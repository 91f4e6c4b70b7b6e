cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2p
nvidia-smi topo -m > gpurun_out/r2p/topo8.txt 2>&1
lscpu | head -20 > gpurun_out/r2p/lscpu8.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2p/bench_n8.json 2> gpurun_out/r2p/bench_n8.err; tail -3 gpurun_out/r2p/bench_n8.err
ls -la gpurun_out/r2p
