"""
MiniSpice: a small, self-contained stand-in for the handful of *scalar, once per
frame* SPICE services the hot path's host side needs when ``spiceypy`` is not
installed (it is installed neither in the authoring container nor on the GPU
box): UTC->ET, barycentric states from type 2/3 SPK segments, IAU body
orientation from PCK constants, and body constants.

It is NOT on the per-pixel path: the per-pixel work is done by the CUDA kernels
(``planetmapper_b200/csrc``). The reference gets these numbers from CSPICE via
spiceypy (planetmapper/base.py:815-836, planetmapper/body.py:522-588).

The IAU orientation model is the published one used by the PCK
(``pck00010.tpc`` comments): RA/DEC/W polynomials plus NUT_PREC trigonometric
series with the system barycentre's NUT_PREC_ANGLES; the J2000->body matrix is
``Rz(W) Rx(pi/2 - DEC) Rz(pi/2 + RA)`` with coordinate-rotation matrices.
"""

from __future__ import annotations

import glob
import json
import math
import os
from pathlib import Path

import numpy as np

from . import daf
from .textkernel import load_text_kernel

CLIGHT = 299792.458  # km/s, IAU value returned by spice.clight()
SPD = 86400.0
J2000_JD = 2451545.0

BODY_IDS = {
    'SSB': 0, 'SOLAR SYSTEM BARYCENTER': 0,
    'MERCURY BARYCENTER': 1, 'VENUS BARYCENTER': 2, 'EARTH BARYCENTER': 3, 'EMB': 3,
    'MARS BARYCENTER': 4, 'JUPITER BARYCENTER': 5, 'SATURN BARYCENTER': 6,
    'URANUS BARYCENTER': 7, 'NEPTUNE BARYCENTER': 8, 'PLUTO BARYCENTER': 9,
    'SUN': 10, 'MERCURY': 199, 'VENUS': 299, 'EARTH': 399, 'MOON': 301,
    'MARS': 499, 'JUPITER': 599, 'IO': 501, 'EUROPA': 502, 'GANYMEDE': 503,
    'CALLISTO': 504, 'AMALTHEA': 505, 'THEBE': 514, 'ADRASTEA': 515, 'METIS': 516,
    'SATURN': 699, 'URANUS': 799, 'NEPTUNE': 899, 'PLUTO': 999,
    'HST': -48, 'HUBBLE SPACE TELESCOPE': -48, 'JWST': -170,
}
BODY_NAMES = {}
for _k, _v in BODY_IDS.items():
    BODY_NAMES.setdefault(_v, _k)
BODY_NAMES[0] = 'SOLAR SYSTEM BARYCENTER'
BODY_NAMES[-48] = 'HST'

# (TAI-UTC, first formal-UTC second past J2000 at which it applies); naif0012.tls
_LEAP_DATES = [
    (10, (1972, 1, 1)), (11, (1972, 7, 1)), (12, (1973, 1, 1)), (13, (1974, 1, 1)),
    (14, (1975, 1, 1)), (15, (1976, 1, 1)), (16, (1977, 1, 1)), (17, (1978, 1, 1)),
    (18, (1979, 1, 1)), (19, (1980, 1, 1)), (20, (1981, 7, 1)), (21, (1982, 7, 1)),
    (22, (1983, 7, 1)), (23, (1985, 7, 1)), (24, (1988, 1, 1)), (25, (1990, 1, 1)),
    (26, (1991, 1, 1)), (27, (1992, 7, 1)), (28, (1993, 7, 1)), (29, (1994, 7, 1)),
    (30, (1996, 1, 1)), (31, (1997, 7, 1)), (32, (1999, 1, 1)), (33, (2006, 1, 1)),
    (34, (2009, 1, 1)), (35, (2012, 7, 1)), (36, (2015, 7, 1)), (37, (2017, 1, 1)),
]
_DELTA_T_A = 32.184
_K = 1.657e-3
_EB = 1.671e-2
_M0, _M1 = 6.239996, 1.99096871e-7


def _days_from_civil(y: int, m: int, d: int) -> int:
    """Days from 2000-01-01 (proleptic Gregorian) to y-m-d."""
    y -= m <= 2
    era = (y if y >= 0 else y - 399) // 400
    yoe = y - era * 400
    doy = (153 * (m + (-3 if m > 2 else 9)) + 2) // 5 + d - 1
    doe = yoe * 365 + yoe // 4 - yoe // 100 + doy
    return era * 146097 + doe - 730425  # 730425 = days 0000-03-01 -> 2000-01-01


def formal_utc_seconds(y, mo, d, h=0, mi=0, s=0.0) -> float:
    """Seconds past J2000 (2000-01-01T12:00:00) counted as 86400 s/day."""
    return (_days_from_civil(y, mo, d) - 0.5) * SPD + h * 3600.0 + mi * 60.0 + s


def utc2et(utc: str) -> float:
    """ISO-like UTC string -> TDB seconds past J2000 (``spice.str2et`` for ISO)."""
    s = utc.strip().upper().rstrip('Z')
    date, _, time = s.replace(' ', 'T').partition('T')
    y, mo, d = (int(v) for v in date.split('-'))
    h = mi = 0
    sec = 0.0
    if time:
        parts = time.split(':')
        h = int(parts[0])
        if len(parts) > 1:
            mi = int(parts[1])
        if len(parts) > 2:
            sec = float(parts[2])
    formal = formal_utc_seconds(y, mo, d, h, mi, sec)
    day_start = formal_utc_seconds(y, mo, d)
    dat = 9
    for leap, (ly, lm, ld) in _LEAP_DATES:
        if day_start >= formal_utc_seconds(ly, lm, ld):
            dat = leap
    tai = formal + dat
    tdt = tai + _DELTA_T_A
    m = _M0 + _M1 * tdt
    e = m + _EB * math.sin(m)
    return tdt + _K * math.sin(e)


def rot_axis(theta: float, axis: int) -> np.ndarray:
    """``spice.rotate``: coordinate-system rotation by theta about axis 1/2/3."""
    c, s = math.cos(theta), math.sin(theta)
    if axis == 1:
        return np.array([[1, 0, 0], [0, c, s], [0, -s, c]], dtype=float)
    if axis == 2:
        return np.array([[c, 0, -s], [0, 1, 0], [s, 0, c]], dtype=float)
    return np.array([[c, s, 0], [-s, c, 0], [0, 0, 1]], dtype=float)


def _drot_axis(theta: float, axis: int) -> np.ndarray:
    """d/dtheta of rot_axis."""
    c, s = math.cos(theta), math.sin(theta)
    if axis == 1:
        return np.array([[0, 0, 0], [0, -s, c], [0, -c, -s]], dtype=float)
    if axis == 2:
        return np.array([[-s, 0, -c], [0, 0, 0], [c, 0, -s]], dtype=float)
    return np.array([[-s, c, 0], [-c, -s, 0], [0, 0, 0]], dtype=float)


class MiniSpice:
    """Ephemeris + constants provider backed by SPK/PCK files or an extract."""

    name = 'minispice'

    def __init__(self, segments: list[daf.Segment], pool: dict[str, list]):
        self.segments = segments  # in load order; later entries take precedence
        self.pool = pool
        # per body, highest precedence first (same order as scanning `segments` backwards)
        self._by_target: dict[int, list[daf.Segment]] = {}
        for seg in reversed(segments):
            self._by_target.setdefault(seg.target, []).append(seg)

    # ---- construction -----------------------------------------------------------
    @classmethod
    def from_kernel_dir(cls, kernel_dir: str) -> 'MiniSpice':
        """Load every .bsp/.tpc under kernel_dir in the reference's load order
        (deepest directory first, then alphabetical; later loads win:
        planetmapper/base.py:939-977)."""
        paths = set(glob.glob(os.path.join(kernel_dir, '**', '*.*'), recursive=True))
        paths = sorted(
            paths,
            key=lambda p: (-len(Path(p).resolve().parts), os.path.dirname(p),
                           os.path.basename(p), os.path.normpath(p), p),
        )
        segments: list[daf.Segment] = []
        pool: dict[str, list] = {}
        for p in paths:
            low = p.lower()
            if low.endswith('.bsp'):
                segments.extend(daf.read_spk(p))
            elif low.endswith(('.tpc', '.tls', '.tf', '.ti')):
                pool.update(load_text_kernel(p))
        return cls(segments, pool)

    @classmethod
    def from_extract(cls, npz_path: str, pool_json_path: str) -> 'MiniSpice':
        with open(pool_json_path, 'r', encoding='utf-8') as f:
            pool = json.load(f)
        return cls(daf.load_extract(npz_path), pool)

    # ---- constants -----------------------------------------------------------------
    def clight(self) -> float:
        return CLIGHT

    def bods2c(self, name) -> int:
        if isinstance(name, (int, np.integer)):
            return int(name)
        s = str(name).strip().upper()
        try:
            return int(s)
        except ValueError:
            pass
        if s not in BODY_IDS:
            raise KeyError(f'unknown body name {name!r}')
        return BODY_IDS[s]

    def bodc2n(self, code: int) -> str:
        return BODY_NAMES.get(int(code), str(code))

    def bodvar(self, body: int, item: str) -> np.ndarray:
        key = f'BODY{int(body)}_{item}'
        if key not in self.pool:
            raise KeyError(f'{key} not found in the constants pool')
        return np.array(self.pool[key], dtype=float)

    def utc2et(self, utc: str) -> float:
        return utc2et(utc)

    # ---- ephemeris -----------------------------------------------------------------
    def _find_segment(self, target: int, et: float) -> daf.Segment:
        for seg in self._by_target.get(target, ()):
            if seg.covers(et):
                return seg
        raise LookupError(
            f'no SPK data for body {target} at et={et!r} (types 2/3 only)'
        )

    def ssb_state(self, body: int, et: float) -> np.ndarray:
        """State of body relative to the solar system barycentre, J2000."""
        state = np.zeros(6)
        b = int(body)
        guard = 0
        while b != 0:
            seg = self._find_segment(b, et)
            if seg.frame != 1:
                raise NotImplementedError('only J2000 (frame 1) segments are supported')
            state += seg.state(et)
            b = seg.center
            guard += 1
            if guard > 8:
                raise RuntimeError('SPK centre chain does not reach the barycentre')
        return state

    # ---- orientation ---------------------------------------------------------------
    def _euler_angles(self, body: int, et: float):
        """RA, DEC, W (radians) and their time derivatives (rad/s)."""
        body = int(body)
        t_cy = et / (SPD * 36525.0)
        d = et / SPD
        ra_c = self.bodvar(body, 'POLE_RA')
        dec_c = self.bodvar(body, 'POLE_DEC')
        pm_c = self.bodvar(body, 'PM')
        ra = ra_c[0] + ra_c[1] * t_cy + (ra_c[2] * t_cy * t_cy if len(ra_c) > 2 else 0.0)
        dra = (ra_c[1] + (2 * ra_c[2] * t_cy if len(ra_c) > 2 else 0.0)) / (SPD * 36525.0)
        dec = dec_c[0] + dec_c[1] * t_cy + (dec_c[2] * t_cy * t_cy if len(dec_c) > 2 else 0.0)
        ddec = (dec_c[1] + (2 * dec_c[2] * t_cy if len(dec_c) > 2 else 0.0)) / (SPD * 36525.0)
        w = pm_c[0] + pm_c[1] * d + (pm_c[2] * d * d if len(pm_c) > 2 else 0.0)
        dw = (pm_c[1] + (2 * pm_c[2] * d if len(pm_c) > 2 else 0.0)) / SPD

        def series(name):
            key = f'BODY{body}_{name}'
            return np.array(self.pool[key], dtype=float) if key in self.pool else None

        nra, ndec, npm = series('NUT_PREC_RA'), series('NUT_PREC_DEC'), series('NUT_PREC_PM')
        if nra is not None or ndec is not None or npm is not None:
            if body >= 100:
                bary = body // 100
            else:
                bary = body
            ang = np.array(self.pool[f'BODY{bary}_NUT_PREC_ANGLES'], dtype=float)
            ang = ang.reshape(-1, 2)
            theta = np.deg2rad(ang[:, 0] + ang[:, 1] * t_cy)
            dtheta = np.deg2rad(ang[:, 1]) / (SPD * 36525.0)
            if nra is not None:
                k = len(nra)
                ra += float(nra @ np.sin(theta[:k]))
                dra += float(nra @ (np.cos(theta[:k]) * dtheta[:k]))
            if ndec is not None:
                k = len(ndec)
                dec += float(ndec @ np.cos(theta[:k]))
                ddec += float(-ndec @ (np.sin(theta[:k]) * dtheta[:k]))
            if npm is not None:
                k = len(npm)
                w += float(npm @ np.sin(theta[:k]))
                dw += float(npm @ (np.cos(theta[:k]) * dtheta[:k]))
        rad = math.pi / 180.0
        w = math.fmod(w, 360.0)
        return ra * rad, dec * rad, w * rad, dra * rad, ddec * rad, dw * rad

    def orientation(self, body: int, et: float) -> tuple[np.ndarray, np.ndarray]:
        """J2000 -> IAU body-fixed rotation R (v_body = R v_j2000) at et, and the
        angular velocity of the body frame expressed in body-fixed axes, defined by
        dR/dt = -[omega]x R."""
        ra, dec, w, dra, ddec, dw = self._euler_angles(body, et)
        a3, a1, b3 = w, math.pi / 2 - dec, math.pi / 2 + ra
        r3, r1, q3 = rot_axis(a3, 3), rot_axis(a1, 1), rot_axis(b3, 3)
        rmat = r3 @ r1 @ q3
        drmat = (
            (_drot_axis(a3, 3) * dw) @ r1 @ q3
            + r3 @ (_drot_axis(a1, 1) * (-ddec)) @ q3
            + r3 @ r1 @ (_drot_axis(b3, 3) * dra)
        )
        om = -drmat @ rmat.T
        omega = np.array([om[2, 1], om[0, 2], om[1, 0]])
        return rmat, omega

    def body_frame_name(self, body: int) -> str:
        return 'IAU_' + self.bodc2n(body)
