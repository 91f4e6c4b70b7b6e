#!/usr/bin/env python
"""Device timings of the map-direction kernel variants at C4's grid (6.48 M cells) and C3's orthographic grid."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from planetmapper_b200 import _lib as L  # noqa: E402


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


bc = bench.load_bc()
_, lo, la = bench.c4_inputs(8)
fr = bench.c4_frame(bc)
fd, lod, lad = L.to_device(fr), L.to_device(lo), L.to_device(la)
xy = L.mask_from_names(['PIXEL-X', 'PIXEL-Y'])
wide = xy | L.mask_from_names(['RA'])
res = {'cells': int(lo.size)}
o2 = torch.empty((2,) + lo.shape, dtype=torch.float64, device='cuda')
o3 = torch.empty((3,) + lo.shape, dtype=torch.float64, device='cuda')
o26 = torch.empty((26,) + lo.shape, dtype=torch.float64, device='cuda')
res['xy_host_ms'] = timed(lambda: L.backplanes_map_host(fr, lod, lad, xy, out=o2))
res['xy_dev_ms'] = timed(lambda: L.backplanes_map(fd, lod, lad, xy, out=o2))
res['xy_ra_general_host_ms'] = timed(lambda: L.backplanes_map_host(fr, lod, lad, wide, out=o3))
res['all_host_ms'] = timed(lambda: L.backplanes_map_host(fr, lod, lad, L.ALL_PLANES, out=o26))
res['all_dev_ms'] = timed(lambda: L.backplanes_map(fd, lod, lad, L.ALL_PLANES, out=o26))
print(json.dumps(res))
