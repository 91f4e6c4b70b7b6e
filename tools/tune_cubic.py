#!/usr/bin/env python
"""Times the cubic gather tuning variants (PM_CUBIC_VARIANT, debug only) on one C4 chunk."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, numpy as np, torch
sys.path.insert(0, %r)
import bench
from planetmapper_b200 import _lib as L
from planetmapper_b200 import frame as F
bc = bench.load_bc()
sz = 64
fr = F.pack_frame(bc, nx=sz, ny=sz, x0=31.5, y0=31.5, r0=28.0, rotation_radians=0.0)
lons = np.arange(0.05, 360, 0.1)[::-1]; lats = np.arange(-90 + 0.05, 90, 0.1)
lo, la = np.meshgrid(lons, lats)
xy = L.backplanes_map(L.to_device(fr), L.to_device(lo), L.to_device(la), L.mask_from_names(['PIXEL-X', 'PIXEL-Y']))
nl = int(__import__('os').environ.get('PM_TUNE_NL', '256'))
rng = np.random.default_rng(0)
cube_h = rng.normal(1.0, 0.1, (nl, sz, sz)); cube_h[rng.random(cube_h.shape) < 0.01] = np.nan
sp = L.spline_prepare(L.to_device(cube_h), 3)
out = torch.empty((nl,) + lo.shape, dtype=torch.float64, device='cuda')
for _ in range(2): L.gather(sp, xy[0], xy[1], 3, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): L.gather(sp, xy[0], xy[1], 3, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print('variant', %r, 'ms %%.3f' %% ms, 'GB/s written %%.0f' %% (nl * lo.size * 8 / ms / 1e6), 'checksum %%.15g' %% torch.nan_to_num(out).sum().item())
'''
for v in (sys.argv[1:] or ['0', '1', '2', '3', '4', '5', '6']):
    subprocess.run([sys.executable, '-c', CHILD % (ROOT, v)], env=dict(os.environ, PM_CUBIC_VARIANT=v), check=False)
