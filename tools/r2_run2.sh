cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2b
{
for c in 4 5 6 3; do
  for p in 0 4 8 16; do
    python tools/tune_img.py variants/libpm_p$c.so:PM_IMG_PER_THREAD=$p
  done
done
} > gpurun_out/r2b/tune_img.log 2>&1
cat gpurun_out/r2b/tune_img.log
