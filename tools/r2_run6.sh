cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2f
python -m pytest tests -m gpu -q -x > gpurun_out/r2f/gpu_tests.log 2>&1; tail -5 gpurun_out/r2f/gpu_tests.log
python bench.py --gpus 1 --steps 50 --warmup 5 > gpurun_out/r2f/bench_n1.json 2> gpurun_out/r2f/bench_n1.err; tail -5 gpurun_out/r2f/bench_n1.err
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r2f/bench_reference.json 2> gpurun_out/r2f/bench_reference.err
ls -la gpurun_out/r2f
