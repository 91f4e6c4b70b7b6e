cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2h
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2h/bench_n2.json 2> gpurun_out/r2h/bench_n2.err; tail -5 gpurun_out/r2h/bench_n2.err
nvidia-smi topo -m > gpurun_out/r2h/topo.txt 2>&1
ls -la gpurun_out/r2h
