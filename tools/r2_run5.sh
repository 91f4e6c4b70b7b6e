cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2e
python tools/tune_img.py variants/libpm_f4.so > gpurun_out/r2e/tune_img.log 2>&1
cat gpurun_out/r2e/tune_img.log
python -m pytest tests -m gpu -q -x 2>&1 | tail -5
