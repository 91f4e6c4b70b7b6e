#!/usr/bin/env python
"""Times the image kernel's tuning variants (PM_IMG_PER_THREAD, debug only) on C2.
   python tools/tune_img.py [variants...]"""
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
for _ in range(5):
    L.backplanes_img(fd, bench.SZ, bench.SZ, mask, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for rep in range(5):
    e0.record()
    for _ in range(20):
        L.backplanes_img(fd, bench.SZ, bench.SZ, mask, out=out)
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 20)
chk = torch.nan_to_num(out, nan=0.0).sum().item()
print('variant', %r, 'ms %%.4f' %% best, 'Mpix/s %%.0f' %% (bench.SZ * bench.SZ / best / 1e3), 'checksum %%.17g' %% chk)
'''
for v in (sys.argv[1:] or ['1', '2', '4', '8', '16', '0']):
    env = dict(os.environ, PM_IMG_PER_THREAD=v)
    subprocess.run([sys.executable, '-c', CHILD % (ROOT, v)], env=env, check=False)
