#!/usr/bin/env python
"""Short driver for ncu captures: launches each hot kernel a few times at bench shapes.
   python tools/profile_run.py img|gather|map"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from planetmapper_b200 import _lib as L  # noqa: E402
from planetmapper_b200 import frame as F  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else 'img'
bc = bench.load_bc()
if what == 'img':
    fr = bench.c2_frame(bc)
    fd = L.to_device(fr[None])
    mask = L.mask_from_names(bench.C2_NAMES)
    out = torch.empty((1, 12, bench.SZ, bench.SZ), dtype=torch.float64, device='cuda')
    for _ in range(4):
        L.backplanes_img(fd, bench.SZ, bench.SZ, mask, out=out)
    out26 = torch.empty((1, 26, bench.SZ, bench.SZ), dtype=torch.float64, device='cuda')
    for _ in range(2):
        L.backplanes_img(fd, bench.SZ, bench.SZ, L.ALL_PLANES, out=out26)
else:
    sz = 64
    fr = F.pack_frame(bc, nx=sz, ny=sz, x0=31.5, y0=31.5, r0=28.0, rotation_radians=0.0)
    lons = np.arange(0.05, 360, 0.1)[::-1]
    lats = np.arange(-90 + 0.05, 90, 0.1)
    lo, la = np.meshgrid(lons, lats)
    fd = L.to_device(fr)
    lod, lad = L.to_device(lo), L.to_device(la)
    for _ in range(2):
        xy = L.backplanes_map(fd, lod, lad, L.mask_from_names(['PIXEL-X', 'PIXEL-Y']))
    if what == 'map':
        for _ in range(2):
            L.backplanes_map(fd, lod, lad, L.ALL_PLANES)
    else:
        nl = 256
        rng = np.random.default_rng(0)
        cube_h = rng.normal(1.0, 0.1, (nl, sz, sz))
        cube_h[rng.random(cube_h.shape) < 0.01] = np.nan
        cube = L.to_device(cube_h)
        out = torch.empty((nl,) + lo.shape, dtype=torch.float64, device='cuda')
        for mode in (0, 1, 3):
            if mode:
                coef = L.spline_prepare(cube, mode)
            else:
                coef = cube
            for _ in range(2):
                L.gather(coef, xy[0], xy[1], mode, out=out)
torch.cuda.synchronize()
print('done', what)
