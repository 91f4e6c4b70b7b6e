#!/usr/bin/env python
"""Developer smoke check on a GPU box: CUDA kernels vs the CPU oracle + rough timings.
Not a test and not the bench; prints max differences per backplane."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import map_img_oracle as MO  # noqa: E402
from oracle import oracle as O  # noqa: E402
from planetmapper_b200 import _lib as L  # noqa: E402
from planetmapper_b200 import frame as F  # noqa: E402


def load_bc():
    with open(os.path.join(ROOT, 'tests', 'golden', 'jupiter_hst_2005.json')) as f:
        return F.BodyConstants.from_json_dict(json.load(f))


def cmp(name, a, b):
    same = np.array_equal(np.isnan(a), np.isnan(b))
    ok = np.isfinite(a) & np.isfinite(b)
    d = float(np.max(np.abs(a[ok] - b[ok]))) if ok.any() else 0.0
    scale = float(np.max(np.abs(b[ok]))) if ok.any() else 0.0
    nm = int(np.sum(np.isnan(a) != np.isnan(b)))
    print(f'  {name:18s} mask_same={same} (mismatch {nm}) maxabs={d:.3e} scale={scale:.3e}')


def main():
    bc = load_bc()
    print('device', torch.cuda.get_device_name(0), 'fp64 probe TFLOP/s', L.fp64_peak_probe())
    for (nx, ny, x0, y0, r0, rot) in [(7, 10, 2.5, 3.1, 3.9, 123.456), (200, 160, 99.5, 79.5, 70.0, 30.0)]:
        fr = F.pack_frame(bc, nx=nx, ny=ny, x0=x0, y0=y0, r0=r0, rotation_radians=np.deg2rad(rot))
        ref = O.backplanes_img(fr, nx, ny)
        got = L.backplanes_img(L.to_device(fr[None]), nx, ny).cpu().numpy()[0]
        print(f'img {nx}x{ny}')
        for k, n in enumerate(L.PLANE_NAMES):
            cmp(n, got[k], ref[k])
        # map direction
        lons = np.arange(2.5, 360, 5.0)[::-1]
        lats = np.arange(-87.5, 90, 5.0)
        lo, la = np.meshgrid(lons, lats)
        refm = O.backplanes_map(fr, lo, la)
        fd = L.to_device(fr)
        gotm = L.backplanes_map(fd, L.to_device(lo), L.to_device(la)).cpu().numpy()
        print(f'map 5deg for img {nx}x{ny}')
        for k, n in enumerate(L.PLANE_NAMES):
            cmp(n, gotm[k], refm[k])
        # point transforms
        rng = np.random.default_rng(0)
        xs = rng.uniform(-1, nx, 1000)
        ys = rng.uniform(-1, ny, 1000)
        rl, rb, rmiss = O.xy2lonlat(fr, xs, ys)
        gl, gb, gmiss = L.xy2lonlat(fd, L.to_device(xs), L.to_device(ys))
        cmp('xy2lonlat lon', gl.cpu().numpy(), rl)
        cmp('xy2lonlat lat', gb.cpu().numpy(), rb)
        print('   missed', int(gmiss.item()), rmiss)
        ll = rng.uniform(0, 360, 1000)
        bb = rng.uniform(-90, 90, 1000)
        rx, ry = O.lonlat2xy(fr, ll, bb)
        gx, gy = L.lonlat2xy(fd, L.to_device(ll), L.to_device(bb))
        cmp('lonlat2xy x', gx.cpu().numpy(), rx)
        cmp('lonlat2xy y', gy.cpu().numpy(), ry)
        # gather
        xm, ym = gotm[L.PLANE_ID['PIXEL-X']], gotm[L.PLANE_ID['PIXEL-Y']]
        cube = rng.normal(1, 0.1, (5, ny, nx))
        cube[1, ny // 2, nx // 2] = np.nan
        cube[2, ::3, ::2] = np.nan
        cube[3] = np.nan
        cd = L.to_device(cube)
        xd, yd = L.to_device(xm), L.to_device(ym)
        g = L.gather(cd, xd, yd, L.INTERP_NEAREST).cpu().numpy()
        cmp('gather nearest', g, MO.map_img(cube, xm, ym, 'nearest'))
        for mode, name in ((L.INTERP_LINEAR, 'linear'), (L.INTERP_CUBIC, 'cubic')):
            for prop in (True, False):
                coef = L.spline_prepare(cd, mode)
                g = L.gather(coef, xd, yd, mode,
                             propagate_nan=prop).cpu().numpy()
                r = MO.map_img(cube, xm, ym, name, propagate_nan=prop)
                cmp(f'gather {name} p={int(prop)}', g, r)
    # projections
    a, b = bc.r_eq, bc.r_polar
    for kind, lim, lon0, lat0 in [(1, 1.01, 0, 0), (1, 1.01, 123.456, -2), (1, 1.01, 10, 90), (2, 1.01, 0, 0),
                                  (2, 1.01, 123.456, 90), (2, 1.01, 12.345, 42), (3, 1.01, 0, 0), (3, 1.01, 34, -12)]:
        c = np.linspace(-lim, lim, 101)
        xx, yy = np.meshgrid(c, c)
        rlon, rlat = O.proj_inverse(kind, a, b, lon0, lat0, -1.0, xx, yy)
        glon, glat = L.proj_inverse(kind, a, b, lon0, lat0, -1.0, L.to_device(xx), L.to_device(yy))
        cmp(f'proj{kind} lon0={lon0} lat0={lat0} lon', glon.cpu().numpy(), rlon)
        cmp(f'proj{kind} lat', glat.cpu().numpy(), rlat)
    # timing 2048^2, 12-plane stack
    names = ['LON-GRAPHIC', 'LAT-GRAPHIC', 'LON-CENTRIC', 'LAT-CENTRIC', 'INCIDENCE', 'EMISSION', 'PHASE',
             'AZIMUTH', 'LOCAL-SOLAR-TIME', 'DISTANCE', 'RADIAL-VELOCITY', 'DOPPLER']
    mask = L.mask_from_names(names)
    for sz in (1024, 2048):
        fr = F.pack_frame(bc, nx=sz, ny=sz, x0=(sz - 1) / 2, y0=(sz - 1) / 2, r0=0.9 * (sz - 1) / 2,
                          rotation_radians=0.0)
        fd = L.to_device(fr[None])
        out = L.backplanes_img(fd, sz, sz, mask)
        torch.cuda.synchronize()
        for m, label in ((mask, '12-plane'), (L.ALL_PLANES, '26-plane')):
            out = L.backplanes_img(fd, sz, sz, m)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                L.backplanes_img(fd, sz, sz, m, out=out)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print(f'img {sz}^2 {label}: {ms:.3f} ms  {sz * sz / ms / 1e3:.1f} Mpix/s')
    # gather timing: 64x64 cube of 256 planes -> 0.1 deg grid
    sz = 64
    fr = F.pack_frame(bc, nx=sz, ny=sz, x0=31.5, y0=31.5, r0=28.0, rotation_radians=0.0)
    fd = L.to_device(fr)
    lons = np.arange(0.05, 360, 0.1)[::-1]
    lats = np.arange(-90 + 0.05, 90, 0.1)
    lo, la = np.meshgrid(lons, lats)
    t0 = time.time()
    xy = L.backplanes_map(fd, L.to_device(lo), L.to_device(la), L.mask_from_names(['PIXEL-X', 'PIXEL-Y']))
    torch.cuda.synchronize()
    print('xy_map 0.1deg', time.time() - t0, 's; visible cells', int(torch.isfinite(xy[0]).sum()))
    nl = 256
    cube = torch.randn((nl, sz, sz), dtype=torch.float64, device='cuda')
    out = torch.empty((nl,) + tuple(lo.shape), dtype=torch.float64, device='cuda')
    for mode, name in ((0, 'nearest'), (1, 'linear'), (3, 'cubic')):
        if mode:
            coef = L.spline_prepare(cube, mode)
        else:
            coef = cube
        L.gather(coef, xy[0], xy[1], mode, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            L.gather(coef, xy[0], xy[1], mode, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        vox = nl * lo.size
        print(f'gather {name}: {ms:.3f} ms, {vox / ms / 1e6:.2f} Gvox/s, {vox * 8 / ms / 1e6:.1f} GB/s written')


if __name__ == '__main__':
    main()
