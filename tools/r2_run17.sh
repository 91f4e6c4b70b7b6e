cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2q
timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "nan_repair or gather or series or smoke or mapped or map_img or golden" 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --skip-cube --skip-cpu > gpurun_out/r2q/bench_ts.json 2> gpurun_out/r2q/bench_ts.err; python -c "
import json; d=json.load(open('gpurun_out/r2q/bench_ts.json')); t=d['time_series']; print(t['backplanes_ms'], t['reprojection_ms'], t['host_constants_s'], t['host_workers'])"
