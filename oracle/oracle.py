"""
ctypes loader for the CPU oracle (oracle/libpm_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Nothing under planetmapper_b200/
imports this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
N_PLANES = 26
ALL_PLANES = (1 << N_PLANES) - 1


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, 'libpm_oracle.so')
    src = os.path.join(_HERE, 'pm_oracle.c')
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(['make', '-C', _HERE], check=True, capture_output=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _frame(frame):
    f = np.ascontiguousarray(frame, dtype=np.float64)
    return f


def backplanes_img(frame, nx, ny, mask=ALL_PLANES, with_margin=False):
    f = _frame(frame)
    k = bin(mask).count('1')
    out = np.empty((k, ny, nx))
    margin = np.empty((ny, nx))
    rc = lib().pmo_backplanes_img(_p(f), ctypes.c_int(nx), ctypes.c_int(ny),
                                  ctypes.c_uint64(mask), _p(out), _p(margin))
    assert rc == 0, rc
    return (out, margin) if with_margin else out


def backplanes_map(frame, lon, lat, mask=ALL_PLANES, with_margin=False):
    f = _frame(frame)
    lon = np.ascontiguousarray(lon, dtype=np.float64)
    lat = np.ascontiguousarray(lat, dtype=np.float64)
    k = bin(mask).count('1')
    out = np.empty((k,) + lon.shape)
    margin = np.empty(lon.shape)
    rc = lib().pmo_backplanes_map(_p(f), _p(lon), _p(lat), ctypes.c_int64(lon.size),
                                  ctypes.c_uint64(mask), _p(out), _p(margin))
    assert rc == 0, rc
    return (out, margin) if with_margin else out


def xy2lonlat(frame, x, y):
    f = _frame(frame)
    x, y = np.broadcast_arrays(np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64))
    x = np.ascontiguousarray(x)
    y = np.ascontiguousarray(y)
    lon = np.empty(x.shape)
    lat = np.empty(x.shape)
    missed = ctypes.c_int64(0)
    rc = lib().pmo_xy2lonlat(_p(f), _p(x), _p(y), ctypes.c_int64(x.size), _p(lon), _p(lat),
                             ctypes.byref(missed))
    assert rc == 0, rc
    return lon, lat, missed.value


def lonlat2xy(frame, lon, lat, not_visible_nan=True, alt=0.0, planetocentric=False):
    f = _frame(frame)
    lon, lat = np.broadcast_arrays(np.asarray(lon, dtype=np.float64),
                                   np.asarray(lat, dtype=np.float64))
    lon = np.ascontiguousarray(lon)
    lat = np.ascontiguousarray(lat)
    x = np.empty(lon.shape)
    y = np.empty(lon.shape)
    flags = (1 if not_visible_nan else 0) | (4 if planetocentric else 0)
    rc = lib().pmo_lonlat2xy_alt(_p(f), _p(lon), _p(lat), ctypes.c_int64(lon.size), ctypes.c_double(alt),
                                 ctypes.c_uint32(flags), _p(x), _p(y))
    assert rc == 0, rc
    return x, y


COORD = {'xy': 0, 'angular': 1, 'km': 2, 'radec': 3, 'lonlat': 4, 'centric': 5}


def transform(frame, src, dst, a, b, alt=0.0, not_visible_nan=False, planetocentric=False, aux13=None):
    """One of the Body / BodyXY point transforms, e.g. transform(fr, 'radec', 'lonlat', ra, dec).
    Returns (out_a, out_b, n_missed)."""
    f = _frame(frame)
    a, b = np.broadcast_arrays(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64))
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    oa, ob = np.empty(a.shape), np.empty(a.shape)
    flags = (1 if not_visible_nan else 0) | (4 if planetocentric else 0)
    missed = ctypes.c_int64(0)
    aux = None if aux13 is None else np.ascontiguousarray(aux13, dtype=np.float64)
    rc = lib().pmo_transform(_p(f), ctypes.c_int(COORD[src]), ctypes.c_int(COORD[dst]), _p(a), _p(b),
                             ctypes.c_int64(a.size), ctypes.c_double(alt), ctypes.c_uint32(flags),
                             _p(aux) if aux is not None else None, _p(oa), _p(ob), ctypes.byref(missed))
    assert rc == 0, rc
    return oa, ob, missed.value


def proj_inverse(kind, a, b, lon0, lat0, lon_sign, xx, yy):
    params = np.array([a, b, lon0, lat0, lon_sign], dtype=np.float64)
    xx = np.ascontiguousarray(xx, dtype=np.float64)
    yy = np.ascontiguousarray(yy, dtype=np.float64)
    lon = np.empty(xx.shape)
    lat = np.empty(xx.shape)
    rc = lib().pmo_proj_inverse(ctypes.c_int(kind), _p(params), _p(xx), _p(yy),
                                ctypes.c_int64(xx.size), _p(lon), _p(lat))
    assert rc == 0, rc
    return lon, lat


def proj_forward(kind, a, b, lon0, lat0, lon_sign, lon, lat):
    params = np.array([a, b, lon0, lat0, lon_sign], dtype=np.float64)
    lon = np.ascontiguousarray(lon, dtype=np.float64)
    lat = np.ascontiguousarray(lat, dtype=np.float64)
    xx = np.empty(lon.shape)
    yy = np.empty(lon.shape)
    rc = lib().pmo_proj_forward(ctypes.c_int(kind), _p(params), _p(lon), _p(lat), ctypes.c_int64(lon.size), _p(xx),
                                _p(yy))
    assert rc == 0, rc
    return xx, yy


def gather_nearest(cube, xmap, ymap):
    cube = np.ascontiguousarray(cube, dtype=np.float64)
    if cube.ndim == 2:
        cube = cube[None]
    xmap = np.ascontiguousarray(xmap, dtype=np.float64)
    ymap = np.ascontiguousarray(ymap, dtype=np.float64)
    nl, ny, nx = cube.shape
    out = np.empty((nl,) + xmap.shape)
    rc = lib().pmo_gather_nearest(_p(cube), ctypes.c_int(nl), ctypes.c_int(ny),
                                  ctypes.c_int(nx), _p(xmap), _p(ymap),
                                  ctypes.c_int64(xmap.size), _p(out))
    assert rc == 0, rc
    return out
