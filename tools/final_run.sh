set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/final
python -m pytest tests -m gpu -x -q > gpurun_out/final/gpu_tests.log 2>&1; tail -3 gpurun_out/final/gpu_tests.log
python tools/count_flops_ncu.py > gpurun_out/final/count_flops.log 2>&1; tail -2 gpurun_out/final/count_flops.log
cp profiles/flops_per_pixel.json gpurun_out/final/
python bench.py --gpus 1 --steps 50 --warmup 5 > gpurun_out/final/bench_n1.json 2> gpurun_out/final/bench_n1.err; tail -2 gpurun_out/final/bench_n1.err
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/final/bench_reference.json 2> gpurun_out/final/bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/final/launches.csv python bench.py --steps 2 --warmup 3 --skip-cpu > gpurun_out/final/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:backplanes_img -s 3 -c 2 -o gpurun_out/final/prof_img python tools/profile_run.py img > gpurun_out/final/prof_img.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gather_ -c 6 -o gpurun_out/final/prof_gather python tools/profile_run.py gather > gpurun_out/final/prof_gather.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:backplanes_map -c 3 -o gpurun_out/final/prof_map python tools/profile_run.py map > gpurun_out/final/prof_map.log 2>&1
ncu --set full --clock-control none -k regex:fits_stage -c 1 -o gpurun_out/final/prof_stage python bench.py --steps 2 --warmup 3 --skip-cpu --skip-cube > gpurun_out/final/prof_stage.log 2>&1
compute-sanitizer --tool memcheck python -m pytest tests/test_fits_stage.py tests/test_gpu_parity.py -m gpu -x -q -k "staged or save_observation or gather_vs_scipy or point_transforms" > gpurun_out/final/sanitizer_memcheck.log 2>&1; tail -4 gpurun_out/final/sanitizer_memcheck.log
ls -la gpurun_out/final
