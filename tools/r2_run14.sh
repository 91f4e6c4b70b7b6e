cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2n
timeout 600 python -m pytest tests/test_gpu_api.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
timeout 300 python tools/probe_e2e.py 2>&1 | tee gpurun_out/r2n/probe_e2e.log | head -3
timeout 600 python bench.py --steps 20 --warmup 3 --skip-cube --skip-extra --skip-cpu > gpurun_out/r2n/bench_short.json 2>gpurun_out/r2n/bench_short.err; python -c "
import json; d=json.load(open('gpurun_out/r2n/bench_short.json')); e=d['e2e']; print(e['value'], e['ms_per_step'], e['batched']['ms_per_step'], e['d2h_ceiling']['ms_per_step'], e['frac_of_d2h_ceiling'])"
