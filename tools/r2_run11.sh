cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2k
ncu --set full --clock-control none --import-source on -k regex:gather_cubic -s 1 -c 1 -f -o gpurun_out/r2k/prof_cubic python tools/profile_gather.py cubic > gpurun_out/r2k/prof_cubic.log 2>&1
ncu -i gpurun_out/r2k/prof_cubic.ncu-rep --page raw --csv > gpurun_out/r2k/raw_cubic.csv 2>/dev/null
ncu -i gpurun_out/r2k/prof_cubic.ncu-rep --page source --csv > gpurun_out/r2k/source_cubic.csv 2>/dev/null
ls -la gpurun_out/r2k
