cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2l
for v in ca s4ca s4cg; do
echo $v; PM_B200_LIBRARY=$PWD/variants/libpm_$v.so PM_TUNE_NL=512 timeout 300 python tools/tune_cubic.py 1
done 2>&1 | tee gpurun_out/r2l/tune_cubic2.log
PM_B200_LIBRARY=$PWD/variants/libpm_ca.so PM_CUBIC_VARIANT=1 ncu --set full --clock-control none --import-source on -k regex:gather_cubic -s 1 -c 1 -f -o gpurun_out/r2l/prof_cubic python tools/profile_gather.py cubic > gpurun_out/r2l/prof_cubic.log 2>&1
ncu -i gpurun_out/r2l/prof_cubic.ncu-rep --page source --csv > gpurun_out/r2l/source_cubic.csv 2>/dev/null
ncu -i gpurun_out/r2l/prof_cubic.ncu-rep --page raw --csv > gpurun_out/r2l/raw_cubic.csv 2>/dev/null
