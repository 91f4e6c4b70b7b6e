cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2g
python tools/profile_c4_e2e.py > gpurun_out/r2g/profile_c4_e2e.log 2>&1; tail -45 gpurun_out/r2g/profile_c4_e2e.log
python -m pytest tests/test_transforms.py tests/test_gpu_api.py -m gpu -q -x 2>&1 | tail -5
