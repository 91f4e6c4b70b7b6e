set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2a
nvidia-smi topo -m > gpurun_out/r2a/topo.txt 2>&1
lscpu > gpurun_out/r2a/lscpu.txt 2>&1
numactl -H >> gpurun_out/r2a/lscpu.txt 2>&1
python -m pytest tests -m gpu -q > gpurun_out/r2a/gpu_tests.log 2>&1; tail -15 gpurun_out/r2a/gpu_tests.log
python bench.py --gpus 1 --steps 50 --warmup 5 > gpurun_out/r2a/bench_n1.json 2> gpurun_out/r2a/bench_n1.err; tail -2 gpurun_out/r2a/bench_n1.err
ls -la gpurun_out/r2a
