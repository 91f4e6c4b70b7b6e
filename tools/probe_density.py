#!/usr/bin/env python
"""Gather throughput against map density: the C4 cube (64 x 64 image, 512 planes) mapped to grids from 1 deg to
0.1 deg (6 ... 400 cells per image pixel) and, for a finer image, 256 x 256 at 0.1 deg (25 cells per pixel)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from planetmapper_b200 import _lib as L  # noqa: E402
from planetmapper_b200 import frame as F  # noqa: E402

bc = bench.load_bc()
peak = bench.measured_peaks()[0]['hbm_gbs']
rows = []
for sz, step, nl in ((64, 1.0, 2048), (64, 0.5, 2048), (64, 0.25, 1024), (64, 0.1, 512), (256, 0.1, 512), (1024, 0.1, 256)):
    fr = F.pack_frame(bc, nx=sz, ny=sz, x0=(sz - 1) / 2, y0=(sz - 1) / 2, r0=0.44 * sz, rotation_radians=0.0)
    lons = np.arange(step / 2, 360, step)[::-1]
    lats = np.arange(-90 + step / 2, 90, step)
    lo, la = np.meshgrid(lons, lats)
    xy = L.backplanes_map_host(fr, L.to_device(lo), L.to_device(la), L.mask_from_names(['PIXEL-X', 'PIXEL-Y']))
    rng = np.random.default_rng(0)
    cube_h = rng.normal(1.0, 0.1, (nl, sz, sz))
    cube_h[rng.random(cube_h.shape) < 0.01] = np.nan
    cube = L.to_device(cube_h)
    out = torch.empty((nl,) + lo.shape, dtype=torch.float64, device='cuda')
    vis = float(torch.isfinite(xy[0]).float().mean())
    row = {'image': sz, 'deg': step, 'planes': nl, 'cells': lo.size, 'cells_per_pixel_on_disc': vis * lo.size / (np.pi * (0.44 * sz) ** 2),
           'out_gb': nl * lo.size * 8 / 1e9}
    modes = (('nearest', L.INTERP_NEAREST), ('linear', L.INTERP_LINEAR), ('cubic', L.INTERP_CUBIC))
    if os.environ.get('PM_PROBE_CUBIC_ONLY'):
        modes = modes[2:]
    for name, mode in modes:
        src = cube if mode == L.INTERP_NEAREST else L.spline_prepare(cube, mode)
        for _ in range(2):
            L.gather(src, xy[0], xy[1], mode, out=out)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            L.gather(src, xy[0], xy[1], mode, out=out)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        bytes_ = nl * lo.size * 8 + nl * sz * sz * 8 + 16 * lo.size
        row[name] = {'ms': ms, 'frac_of_hbm_peak': bytes_ / ms / 1e6 / peak}
        del src
    rows.append(row)
    del cube, out
    torch.cuda.empty_cache()
print(json.dumps(rows))
