cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2i
timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2i/gpu_tests.log 2>&1; tail -5 gpurun_out/r2i/gpu_tests.log
