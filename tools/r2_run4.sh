cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2d
{
for c in 4 5; do
  for o in 1 0; do
  for p in 0 4 8; do
    python tools/tune_img.py variants/libpm_q$c.so:PM_IMG_PER_THREAD=$p:PM_IMG_ORDER=$o
  done
  done
done
} > gpurun_out/r2d/tune_img.log 2>&1
cat gpurun_out/r2d/tune_img.log
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "host_frame or full_size_2048 or image_backplanes" 2>&1 | tail -5
