"""
The kernels' own per-pixel / per-cell code (planetmapper_b200/csrc/pm_device.cuh is
__host__ __device__) instantiated on the CPU by tests/host_check/host_check.cu and
compared with the oracle with the SAME bars as the GPU parity tests.  This is the
no-GPU safety net for algorithmic changes to the device code; the GPU tests remain
the parity tests proper (the MUFU seeds are emulated here).  Test infrastructure only:
nothing in planetmapper_b200/ loads this library.
"""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

from helpers import (IMG_CASES, PID, angle_diff, check_img_planes, check_map_planes, img_case)
from planetmapper_b200 import frame as F

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'host_check', 'host_check.cu')
SO = os.path.join(HERE, 'host_check', '_build', 'libpm_hostcheck.so')
CSRC = os.path.join(os.path.dirname(HERE), 'planetmapper_b200', 'csrc')


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.fixture(scope='module')
def HC():
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        pytest.skip('nvcc not available')
    deps = [SRC] + [os.path.join(CSRC, f) for f in ('pm_device.cuh', 'pm_math.cuh')]
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(d) for d in deps):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.run([nvcc, '-O2', '-std=c++17', '-shared', '-Xcompiler', '-fPIC', '-Wno-deprecated-gpu-targets',
                        '-o', SO, SRC], check=True, capture_output=True)
    return ctypes.CDLL(SO)


def hc_img(HC, fr, nx, ny, mask=(1 << 26) - 1):
    out = np.empty((26, ny, nx))
    f = np.ascontiguousarray(fr, dtype=np.float64)
    assert HC.hc_backplanes_img(_p(f), ctypes.c_int(nx), ctypes.c_int(ny), ctypes.c_uint64(mask), _p(out)) == 0
    return out


def hc_map(HC, fr, lon, lat, mask=(1 << 26) - 1):
    lon = np.ascontiguousarray(lon, dtype=np.float64)
    lat = np.ascontiguousarray(lat, dtype=np.float64)
    out = np.empty((26,) + lon.shape)
    f = np.ascontiguousarray(fr, dtype=np.float64)
    assert HC.hc_backplanes_map(_p(f), _p(lon), _p(lat), ctypes.c_int64(lon.size), ctypes.c_uint64(mask),
                                _p(out)) == 0
    return out


def _ulps(got, want):
    want = np.asarray(want, dtype=np.float64)
    return np.abs(got - want) / np.spacing(np.abs(want))


def test_math_primitives_ulp(HC):
    """pm_math.cuh on the host build (pessimistic 20-bit seeds) against numpy longdouble."""
    rng = np.random.default_rng(0)
    n = 400000
    a = (rng.uniform(-0.5, 0.5, n)) * 10.0 ** rng.uniform(-5, 15, n)
    b = (rng.uniform(-0.5, 0.5, n)) * 10.0 ** rng.uniform(-5, 15, n)
    al, bl = a.astype(np.longdouble), b.astype(np.longdouble)

    def run(kind, x, y=None):
        out = np.empty_like(x)
        assert HC.hc_math(ctypes.c_int(kind), _p(x), _p(y) if y is not None else None, ctypes.c_int64(x.size),
                          _p(out)) == 0
        return out

    def err(got, want):
        want = np.asarray(want)
        return float(np.max(np.abs(got.astype(np.longdouble) - want) / np.spacing(np.abs(want.astype(np.float64)))))

    assert err(run(0, a), 1 / al) <= 1.0
    assert err(run(7, a, b), al / bl) <= 1.0
    x = np.abs(a)
    assert err(run(1, x), 1 / np.sqrt(x.astype(np.longdouble))) <= 1.5
    assert err(run(2, x), np.sqrt(x.astype(np.longdouble))) <= 1.0
    th = rng.uniform(-np.pi / 4, np.pi / 4, n)
    assert err(run(3, th), np.sin(th.astype(np.longdouble))) <= 1.5
    assert err(run(4, th), np.cos(th.astype(np.longdouble))) <= 2.0
    assert err(run(5, a, b), np.arctan2(al, bl)) <= 3.0
    assert err(run(10, a, b), np.arctan2(np.abs(al), bl)) <= 3.0
    assert err(run(11, a, b), np.arctan2(al, np.abs(bl))) <= 3.0
    u = rng.uniform(-1, 1, n)
    u[::7] = 1 - 1e-9 * rng.uniform(0, 1, u[::7].size)
    assert err(run(6, u), np.arccos(u.astype(np.longdouble))) <= 4.0
    big = rng.uniform(-1000, 1000, n)
    assert np.max(np.abs(run(8, big) - np.sin(big.astype(np.longdouble)).astype(np.float64))) <= 3e-16
    assert np.max(np.abs(run(9, big) - np.cos(big.astype(np.longdouble)).astype(np.float64))) <= 3e-16
    # conventions the geometry relies on
    assert run(2, np.array([0.0]))[0] == 0.0
    assert run(5, np.array([0.0]), np.array([0.0]))[0] == 0.0
    assert run(5, np.array([0.0]), np.array([-1.0]))[0] == np.pi
    assert np.isnan(run(6, np.array([1.0000001]))[0]) and run(6, np.array([1.0]))[0] == 0.0


@pytest.mark.parametrize('case', sorted(IMG_CASES))
def test_device_code_image_planes_vs_oracle(HC, oracle, bc_hst, case):
    nx, ny, x0, y0, r0, rot, alt = IMG_CASES[case]
    fr = img_case(bc_hst, nx, ny, x0, y0, r0, rot, alt)
    ref, margin = oracle.backplanes_img(fr, nx, ny, with_margin=True)
    got = hc_img(HC, fr, nx, ny)
    check_img_planes(got, ref, margin, fr, case)


def test_device_code_saturn_rings_vs_oracle(HC, oracle):
    import planetmapper_b200 as pm

    bc = F.build_body_constants(pm.get_default_provider(), 'Saturn', '2004-12-30T12:00:00', 'EARTH')
    nx = ny = 96
    fr = img_case(bc, nx, ny, 47.5, 47.5, 18.0, 10.0)
    ref, margin = oracle.backplanes_img(fr, nx, ny, with_margin=True)
    got = hc_img(HC, fr, nx, ny)
    check_img_planes(got, ref, margin, fr, 'saturn')


def test_device_code_plane_subsets(HC, bc_hst):
    fr = img_case(bc_hst, 60, 50, 29.5, 24.5, 22.0, 12.0)
    full = hc_img(HC, fr, 60, 50)
    for names in (['EMISSION'], ['LON-GRAPHIC', 'LAT-GRAPHIC'], ['RA', 'DEC', 'KM-X'], ['DOPPLER'],
                  ['RING-RADIUS', 'RING-LON-GRAPHIC', 'RING-DISTANCE', 'DISTANCE'],
                  ['LIMB-DISTANCE'], ['LOCAL-SOLAR-TIME'], ['AZIMUTH', 'PIXEL-X']):
        mask = sum(1 << PID[n] for n in names)
        sub = hc_img(HC, fr, 60, 50, mask)
        for n in names:
            assert np.array_equal(sub[PID[n]], full[PID[n]], equal_nan=True), n


@pytest.mark.parametrize('case', ['golden-7x10', 'rot-200x160', 'golden-7x10-alt'])
def test_device_code_map_planes_vs_oracle(HC, oracle, bc_hst, case):
    nx, ny, x0, y0, r0, rot, alt = IMG_CASES[case]
    fr = img_case(bc_hst, nx, ny, x0, y0, r0, rot, alt)
    lons = np.arange(1.5, 360, 3.0)[::-1]
    lats = np.arange(-88.5, 90, 3.0)
    lo, la = np.meshgrid(lons, lats)
    lo = lo.copy()
    lo[0, 0] = np.nan
    la[1, 1] = np.inf
    lo[2, 2] = -725.0
    ref, margin = oracle.backplanes_map(fr, lo, la, with_margin=True)
    got = hc_map(HC, fr, lo, la)
    check_map_planes(got, ref, margin, fr, nx, ny, case)


def test_device_code_point_transforms_vs_oracle(HC, oracle, bc_hst):
    fr = img_case(bc_hst, 15, 10, 5, 8, 3, 45)
    f = np.ascontiguousarray(fr, dtype=np.float64)
    rng = np.random.default_rng(1)
    xs = np.concatenate([rng.uniform(-2, 12, 4000), [np.nan, 5.0, np.inf, 5.0]])
    ys = np.concatenate([rng.uniform(2, 14, 4000), [8.0, np.nan, 8.0, 8.0]])
    rl, rb, rmiss = oracle.xy2lonlat(fr, xs, ys)
    gl, gb = np.empty_like(xs), np.empty_like(xs)
    gmiss = ctypes.c_int64(0)
    assert HC.hc_xy2lonlat(_p(f), _p(xs), _p(ys), ctypes.c_int64(xs.size), _p(gl), _p(gb), ctypes.byref(gmiss)) == 0
    assert (np.isnan(gl) != np.isnan(rl)).sum() <= 2 and abs(gmiss.value - rmiss) <= 2
    core = np.isfinite(gl) & np.isfinite(rl) & (np.hypot(xs - 5, ys - 8) < 2.5)
    assert np.max(np.abs(gb[core] - rb[core])) < 1e-9
    assert np.max(angle_diff(gl[core], rl[core]) * np.cos(np.deg2rad(rb[core]))) < 1e-9
    lon = np.concatenate([rng.uniform(-360, 720, 4000), [np.nan, 0.0, np.inf]])
    lat = np.concatenate([rng.uniform(-90, 90, 4000), [0.0, np.nan, 0.0]])
    for nvn in (True, False):
        rx, ry = oracle.lonlat2xy(fr, lon, lat, not_visible_nan=nvn)
        gx, gy = np.empty_like(lon), np.empty_like(lon)
        assert HC.hc_lonlat2xy(_p(f), _p(lon), _p(lat), ctypes.c_int64(lon.size), ctypes.c_uint32(1 if nvn else 0),
                               _p(gx), _p(gy)) == 0
        assert np.array_equal(np.isnan(gx), np.isnan(rx))
        ok = np.isfinite(rx)
        assert np.max(np.abs(gx[ok] - rx[ok])) < 1e-9 and np.max(np.abs(gy[ok] - ry[ok])) < 1e-9
