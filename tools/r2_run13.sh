cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2m
timeout 300 python tools/probe_e2e.py 2>&1 | tee gpurun_out/r2m/probe_e2e.log
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider -k "dense_cubic" 2>&1 | tail -6
