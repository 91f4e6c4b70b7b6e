#!/usr/bin/env python
"""ncu driver for the gather kernels at the bench's own launch shape: one 512-plane chunk of the C4 cube
(3000 x 64 x 64 -> 0.1 deg grid) per interpolation.   python tools/profile_gather.py [nearest linear cubic]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from planetmapper_b200 import _lib as L  # noqa: E402

modes = sys.argv[1:] or ['nearest', 'linear', 'cubic']
bc = bench.load_bc()
cube_h, lo, la = bench.c4_inputs(1024)
fd = L.to_device(bench.c4_frame(bc))
xy = L.backplanes_map(fd, L.to_device(lo), L.to_device(la), L.mask_from_names(['PIXEL-X', 'PIXEL-Y']))
cube = L.to_device(cube_h)
out = torch.empty((bench.C4_CHUNK,) + lo.shape, dtype=torch.float64, device='cuda')
for name in modes:
    mode = {'nearest': L.INTERP_NEAREST, 'linear': L.INTERP_LINEAR, 'cubic': L.INTERP_CUBIC}[name]
    src = cube if mode == L.INTERP_NEAREST else L.spline_prepare(cube, mode)
    for begin in (0, 512):
        L.gather(src, xy[0], xy[1], mode, plane_begin=begin, plane_count=bench.C4_CHUNK, out=out)
torch.cuda.synchronize()
print('done')
