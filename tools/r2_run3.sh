cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2c
for v in p4 p6; do
PM_B200_LIBRARY=$PWD/variants/libpm_$v.so ncu --set full --clock-control none --import-source on -k regex:backplanes_img_param -s 4 -c 1 -f -o gpurun_out/r2c/prof_img_$v python tools/profile_img.py > gpurun_out/r2c/prof_$v.log 2>&1
ncu -i gpurun_out/r2c/prof_img_$v.ncu-rep --page raw --csv > gpurun_out/r2c/raw_$v.csv 2>/dev/null
done
ls -la gpurun_out/r2c
