"""
Analytic two-body ephemerides for bodies the loaded SPK kernels do not cover.

BASELINE config C5 names Europa, but the reference's bundled test kernels hold no SPK
segment for body 502 ("Europa has no SPK in the bundled kernels ... option
(ii): synthesise Europa PMFrame constants from a Keplerian orbit about 599 with PCK
radii / orientation").  :class:`KeplerOrbitProvider` wraps any provider and answers
``ssb_state`` for such a body with an elliptic orbit about its planet-system barycentre,
laid in the planet's IAU equatorial plane; every other service (radii, IAU orientation
model, time conversion, the planet and Sun ephemerides) is the wrapped provider's, so the
frames carry the body's real PCK constants - for Europa the triaxial radii
1562.6 / 1560.3 / 1559.5 km - and a physically plausible state.  The positions are
SYNTHETIC: good for parity tests (the kernels and the oracle only see constants) and for
throughput, not for science.
"""
from __future__ import annotations

import math

import numpy as np

# body id -> (centre id, semi-major axis km, eccentricity, sidereal period days, mean anomaly at J2000 deg,
#             argument of pericentre deg); orbit plane = IAU equator of planet `centre // 100 * 100 + 99`
ORBITS = {
    501: (5, 421_800.0, 0.0041, 1.769138, 171.0, 84.1),     # Io
    502: (5, 671_100.0, 0.0094, 3.551181, 345.4, 88.97),    # Europa
    503: (5, 1_070_400.0, 0.0013, 7.154553, 324.8, 192.4),  # Ganymede
    504: (5, 1_882_700.0, 0.0074, 16.689017, 87.4, 52.6),   # Callisto
    # Amalthea: an OBSERVER 2.5 Jupiter radii from the centre (the reference's close-range mapping test,
    # tests/test_body_xy.py:2592-2608; its jup120 kernel holds 505 as an SPK type 17 segment, which the
    # bundled reader does not evaluate)
    505: (5, 181_365.8, 0.0032, 0.498179, 185.2, 155.9),
}


class KeplerOrbitProvider:
    """Provider decorator: ``ssb_state`` of the bodies in :data:`ORBITS` from an analytic orbit when
    the wrapped provider has no ephemeris for them."""

    name = 'kepler'

    def __init__(self, base, orbits: dict | None = None) -> None:
        self.base = base
        self.orbits = dict(ORBITS if orbits is None else orbits)
        self._planes: dict[int, tuple[np.ndarray, np.ndarray]] = {}

    def __getattr__(self, item):   # clight, bods2c, bodc2n, bodvar, utc2et, orientation ...
        return getattr(self.base, item)

    def _plane(self, centre: int) -> tuple[np.ndarray, np.ndarray]:
        """Orthonormal (p, q) spanning the planet's equator of J2000 (pole from the PCK constants)."""
        if centre not in self._planes:
            planet = centre * 100 + 99
            ra = math.radians(float(self.base.bodvar(planet, 'POLE_RA')[0]))
            dec = math.radians(float(self.base.bodvar(planet, 'POLE_DEC')[0]))
            pole = np.array([math.cos(dec) * math.cos(ra), math.cos(dec) * math.sin(ra), math.sin(dec)])
            p = np.cross([0.0, 0.0, 1.0], pole)
            p /= np.linalg.norm(p)
            self._planes[centre] = (p, np.cross(pole, p))
        return self._planes[centre]

    def kepler_state(self, body: int, et: float) -> np.ndarray:
        """State (km, km/s) of ``body`` relative to its orbit centre, J2000."""
        centre, a, e, period_d, m0_deg, argp_deg = self.orbits[int(body)]
        n = 2.0 * math.pi / (period_d * 86400.0)
        m = math.radians(m0_deg) + n * et
        ecc = m
        for _ in range(12):   # Newton on Kepler's equation (e << 1: converges in 3-4 steps)
            ecc -= (ecc - e * math.sin(ecc) - m) / (1.0 - e * math.cos(ecc))
        ce, se = math.cos(ecc), math.sin(ecc)
        b = a * math.sqrt(1.0 - e * e)
        x, y = a * (ce - e), b * se
        edot = n / (1.0 - e * ce)
        vx, vy = -a * se * edot, b * ce * edot
        w = math.radians(argp_deg)
        cw, sw = math.cos(w), math.sin(w)
        p, q = self._plane(centre)
        pos = (x * cw - y * sw) * p + (x * sw + y * cw) * q
        vel = (vx * cw - vy * sw) * p + (vx * sw + vy * cw) * q
        return np.concatenate([pos, vel])

    def ssb_state(self, body: int, et: float) -> np.ndarray:
        body = int(body)
        if body in self.orbits:
            try:
                return self.base.ssb_state(body, et)   # a real ephemeris wins when there is one
            except (LookupError, KeyError, NotImplementedError):
                return self.base.ssb_state(self.orbits[body][0], et) + self.kepler_state(body, et)
        return self.base.ssb_state(body, et)
