cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2j
timeout 300 python -m pytest tests/test_transforms.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2j/bench_n2.json 2> gpurun_out/r2j/bench_n2.err; tail -3 gpurun_out/r2j/bench_n2.err
ls -la gpurun_out/r2j
