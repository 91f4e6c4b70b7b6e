#!/usr/bin/env python
"""Times tuning variants of the single-frame image kernel on C2 (debug builds only, see
tools/build_variants.sh):   python tools/tune_img.py variants/libpm_a.so[:PM_IMG_PER_THREAD=8] ..."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, torch
sys.path.insert(0, %r)
import bench
from planetmapper_b200 import _lib as L
bc = bench.load_bc()
fr = bench.c2_frame(bc)
fd = L.to_device(fr[None])
mask = L.mask_from_names(bench.C2_NAMES)
out = torch.empty((1, 12, bench.SZ, bench.SZ), dtype=torch.float64, device='cuda')
def run_dev():
    L.backplanes_img(fd, bench.SZ, bench.SZ, mask, out=out)
def run_host():
    L.backplanes_img_host(fr, bench.SZ, bench.SZ, mask, out=out[0])
res = {}
for name, fn in (('device-frame', run_dev), ('host-frame', run_host)):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for rep in range(5):
        e0.record()
        for _ in range(20):
            fn()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 20)
    chk = torch.nan_to_num(out, nan=0.0).sum().item()
    res[name] = (best, chk)
print('variant', %r, ' '.join('%%s ms %%.4f Mpix/s %%.0f chk %%.17g' %% (k, v[0], bench.SZ * bench.SZ / v[0] / 1e3, v[1]) for k, v in res.items()))
'''
for spec in sys.argv[1:]:
    lib, *envs = spec.split(':')
    env = dict(os.environ, PM_B200_LIBRARY=os.path.abspath(lib))
    for e in envs:
        k, v = e.split('=')
        env[k] = v
    subprocess.run([sys.executable, '-c', CHILD % (ROOT, spec)], env=env, check=False)
