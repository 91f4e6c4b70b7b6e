#!/usr/bin/env python
"""Short driver for ncu captures: launches each hot kernel a few times at the BENCH's own launch shapes.
   python tools/profile_run.py img|map|gather|transform"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from planetmapper_b200 import _lib as L  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else 'img'
bc = bench.load_bc()
if what == 'img':
    fr = bench.c2_frame(bc)
    mask = L.mask_from_names(bench.C2_NAMES)
    out = torch.empty((12, bench.SZ, bench.SZ), dtype=torch.float64, device='cuda')
    for _ in range(4):
        L.backplanes_img_host(fr, bench.SZ, bench.SZ, mask, out=out)          # the headline launch
    out26 = torch.empty((26, bench.SZ, bench.SZ), dtype=torch.float64, device='cuda')
    for _ in range(2):
        L.backplanes_img_host(fr, bench.SZ, bench.SZ, L.ALL_PLANES, out=out26)
    fd = L.to_device(np.stack([fr] * 8))
    outb = torch.empty((8, 12, bench.SZ, bench.SZ), dtype=torch.float64, device='cuda')
    for _ in range(2):
        L.backplanes_img(fd, bench.SZ, bench.SZ, mask, out=outb)              # batched (frames in device memory)
elif what == 'transform':
    fr = bench.c2_frame(bc)
    fd = L.to_device(fr)
    g = torch.Generator(device='cuda').manual_seed(1)
    x = torch.rand(10_000_000, dtype=torch.float64, device='cuda', generator=g) * bench.SZ
    y = torch.rand(10_000_000, dtype=torch.float64, device='cuda', generator=g) * bench.SZ
    for _ in range(2):
        L.transform(fd, 'xy', 'radec', x, y)
        L.transform(fd, 'xy', 'lonlat', x, y)
else:
    cube_h, lo, la = bench.c4_inputs(1024)
    fr4 = bench.c4_frame(bc)
    fd = L.to_device(fr4)
    lod, lad = L.to_device(lo), L.to_device(la)
    xy_mask = L.mask_from_names(['PIXEL-X', 'PIXEL-Y'])
    for _ in range(2):
        xy = L.backplanes_map_host(fr4, lod, lad, xy_mask)                    # what map_img launches
    if what == 'map':
        for _ in range(2):
            L.backplanes_map_host(fr4, lod, lad, L.ALL_PLANES)
            L.backplanes_map(fd, lod, lad, xy_mask)                           # frame staged in shared memory
            L.backplanes_map(fd, lod, lad, L.ALL_PLANES)
    else:
        cube = L.to_device(cube_h)
        out = torch.empty((bench.C4_CHUNK,) + lo.shape, dtype=torch.float64, device='cuda')
        for mode in (L.INTERP_NEAREST, L.INTERP_LINEAR, L.INTERP_CUBIC):
            src = cube if mode == L.INTERP_NEAREST else L.spline_prepare(cube, mode)
            for begin in (0, 512):     # one 512-plane chunk per launch, as in bench.py
                L.gather(src, xy[0], xy[1], mode, plane_begin=begin, plane_count=bench.C4_CHUNK, out=out)
torch.cuda.synchronize()
print('done', what)
