"""
A stand-in ``spiceypy`` for CI (TEST INFRASTRUCTURE ONLY): the handful of CSPICE calls
``planetmapper_b200/spice_host.py`` makes - furnsh, clight, bods2c, bodc2n, bodvrd, str2et, spkssb,
sxform - with CSPICE's calling conventions, answered from the repository's MiniSpice reader.

spiceypy / CSPICE are installed neither in the authoring container nor on the GPU box, so without
this module the first branch of ``planetmapper_b200.get_default_provider()`` and ``SpiceProvider``
(kernel load order, argument conventions, the sign convention with which the angular velocity is
read out of ``sxform``) would never execute anywhere.  It is put on ``sys.path`` by
tests/test_spice_host.py only.

``sxform`` is deliberately NOT built from MiniSpice's analytic angular velocity: the derivative block
is a central difference of the rotation, so the test checks ``SpiceProvider.orientation``'s reading of
the 6 x 6 matrix against an independent derivative.
"""
import os

import numpy as np

from planetmapper_b200.minispice import core as _core
from planetmapper_b200.minispice import daf as _daf
from planetmapper_b200.minispice.textkernel import load_text_kernel as _load_text_kernel

_LOADED: list[str] = []
_PROVIDER = None


class NotFoundError(Exception):
    pass


def kclear():
    global _PROVIDER
    _LOADED.clear()
    _PROVIDER = None


def furnsh(path):
    global _PROVIDER
    if not os.path.isfile(path):
        raise OSError(f'SPICE(NOSUCHFILE): {path}')
    _LOADED.append(str(path))
    _PROVIDER = None


def ktotal(kind='ALL'):
    return len(_LOADED)


def _provider():
    """MiniSpice over the furnished files, later loads taking precedence (CSPICE's rule)."""
    global _PROVIDER
    if _PROVIDER is None:
        import json

        segments, pool = [], {}
        for p in _LOADED:
            low = p.lower()
            if low.endswith('.bsp'):
                segments.extend(_daf.read_spk(p))
            elif low.endswith(('.tpc', '.tls', '.tf', '.ti')):
                pool.update(_load_text_kernel(p))
            elif low.endswith('.npz'):      # the repository's ephemeris extract (SPK records of the test kernels)
                segments.extend(_daf.load_extract(p))
            elif low.endswith('.json'):     # ... and its constants pool
                with open(p, encoding='utf-8') as f:
                    pool.update(json.load(f))
        if not segments and not pool:
            raise RuntimeError('SPICE(NOLOADEDFILES)')
        _PROVIDER = _core.MiniSpice(segments, pool)
    return _PROVIDER


def clight():
    return _core.CLIGHT


def bods2c(name):
    try:
        return _provider().bods2c(name)
    except KeyError as exc:
        raise NotFoundError(str(exc)) from exc


def bodc2n(code):
    return _provider().bodc2n(int(code))


def bodvrd(bodynm, item, maxn):
    body = bods2c(bodynm)
    values = _provider().bodvar(body, item)
    return len(values), np.array(values[:maxn], dtype=float)


def str2et(time):
    return _core.utc2et(str(time))


def spkssb(targ, et, ref):
    if str(ref).upper() != 'J2000':
        raise NotImplementedError('fake spiceypy: J2000 only')
    return _provider().ssb_state(int(targ), float(et))


def _rotation(frame, et):
    frame = str(frame).upper()
    if not frame.startswith('IAU_'):
        raise NotImplementedError(f'fake spiceypy: frame {frame}')
    return _provider().orientation(bods2c(frame[4:]), float(et))[0]


def pxform(fromstr, tostr, et):
    a, b = str(fromstr).upper(), str(tostr).upper()
    if a == 'J2000':
        return _rotation(b, et)
    if b == 'J2000':
        return _rotation(a, et).T
    return _rotation(b, et) @ _rotation(a, et).T


def sxform(instring, tostring, et):
    """6 x 6 state transformation [[R, 0], [dR/dt, R]]; dR/dt by central difference (step 1 s: the rotation
    is smooth over hours, the truncation error ~ omega^3 h^2 / 6 ~ 1e-12 relative)."""
    h = 1.0
    r = pxform(instring, tostring, et)
    dr = (pxform(instring, tostring, et + h) - pxform(instring, tostring, et - h)) / (2.0 * h)
    out = np.zeros((6, 6))
    out[:3, :3] = r
    out[3:, 3:] = r
    out[3:, :3] = dr
    return out
