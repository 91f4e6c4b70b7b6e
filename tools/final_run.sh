# Round-2 measurement run (one B200): tests, flop counts, bench lines, ncu launch list + full captures,
# sanitizer.  Outputs land in gpurun_out/final/ and are summarised into profiles/ by tools/ncu_summary.py.
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/final
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/gpu_tests.log 2>&1; tail -3 $O/gpu_tests.log
timeout 600 python tools/count_flops_ncu.py > $O/count_flops.log 2>&1; tail -2 $O/count_flops.log
cp profiles/flops_per_pixel.json $O/
variants/fp64_operands > $O/fp64_operands.log 2>&1
timeout 900 python bench.py --gpus 1 --steps 50 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; tail -2 $O/bench_n1.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --skip-cpu > $O/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:backplanes_img -s 3 -c 4 -f -o $O/prof_img python tools/profile_run.py img > $O/prof_img.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gather_ -c 6 -f -o $O/prof_gather python tools/profile_run.py gather > $O/prof_gather.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:backplanes_map -c 4 -f -o $O/prof_map python tools/profile_run.py map > $O/prof_map.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:transform_kernel -c 2 -f -o $O/prof_transform python tools/profile_run.py transform > $O/prof_transform.log 2>&1
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py tests/test_transforms.py tests/test_triaxial.py -m gpu -x -q -p no:cacheprovider -k "dense_cubic or gather_vs_scipy or transform_pairs or host_frame or triaxial_image" > $O/sanitizer_memcheck.log 2>&1; tail -4 $O/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider -k "dense_cubic or gather_vs_scipy" > $O/sanitizer_racecheck.log 2>&1; tail -4 $O/sanitizer_racecheck.log
ls -la $O
