"""
Per-frame constants ("PMFrame") for the B200 kernels, and the host code that
derives them ONCE per frame from an ephemeris provider.

This is the scalar host half of the hot path (north_star: "computes each frame's
scalar geometry once on the host"). It mirrors what the reference's constructors
obtain from SPICE:

- ``BodyBase.__init__`` planetmapper/base.py:795-837 (``str2et``, ``spkezr`` 'CN')
- ``Body.__init__`` planetmapper/body.py:522-588 (``bodvar`` RADII/PM, ``subpnt``,
  target diameter, ring plane ``nvp2pl``)
- ``Body._get_obsvec2angular_matrix`` planetmapper/body.py:1318-1343
- ``Body._get_km2angular_matrix`` planetmapper/body.py:1625-1639
- ``BodyXY._get_xy2angular_matrix`` planetmapper/body_xy.py:355-373
- ``spice.et2lst``'s frame-wide Sun longitude

A provider offers five primitives - ``utc2et``, ``ssb_state(id, et)``,
``orientation(id, et) -> (R, omega)``, ``bodvar(id, item)``, ``clight()`` - and is
either :class:`planetmapper_b200.minispice.MiniSpice` or the spiceypy adaptor in
``planetmapper_b200/spice_host.py``. Everything derived (light time, sub-point,
ring plane ...) is computed here so both providers share one tested code path.

The struct is all-doubles so that it maps 1:1 onto ``struct PMFrame`` in
``include/pm_b200.h`` (the layout is asserted in tests/test_frame_layout.py).
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

# name -> number of doubles; ORDER MATTERS and must match include/pm_b200.h
PMFRAME_FIELDS: list[tuple[str, int]] = [
    ('et', 1),
    ('clight', 1),
    ('lt0', 1),
    ('t_ref', 1),
    ('P0', 3),
    ('VT', 3),
    ('AT', 3),
    ('VO', 3),
    ('S0', 3),
    ('VS', 3),
    ('lts0', 1),
    ('R0', 9),
    ('omega', 3),
    ('radii', 3),
    ('re', 1),
    ('f', 1),
    ('lon_sign', 1),
    ('prograde', 1),
    ('sub_t', 3),
    ('sub_ray', 3),
    ('sub_obs', 3),
    ('sub_dt', 1),
    ('sub_dist', 1),
    ('ring_n', 3),
    ('ring_c', 1),
    ('sun_lon_lst', 1),
    ('M', 9),
    ('ang2km', 4),
    ('km_per_arcsec', 1),
    ('A', 6),
    ('Ainv', 6),
    ('nx', 1),
    ('ny', 1),
    ('x0', 1),
    ('y0', 1),
    ('r_cut2', 1),
    ('optimize_speed', 1),
    ('r_eq', 1),
    ('reserved', 1),
]
PMFRAME_OFFSETS: dict[str, tuple[int, int]] = {}
_off = 0
for _name, _n in PMFRAME_FIELDS:
    PMFRAME_OFFSETS[_name] = (_off, _n)
    _off += _n
PMFRAME_NDOUBLES = _off


def _rot_axis(theta: float, axis: int) -> np.ndarray:
    c, s = math.cos(theta), math.sin(theta)
    if axis == 1:
        return np.array([[1, 0, 0], [0, c, s], [0, -s, c]], dtype=float)
    if axis == 2:
        return np.array([[c, 0, -s], [0, 1, 0], [s, 0, c]], dtype=float)
    return np.array([[c, s, 0], [-s, c, 0], [0, 0, 1]], dtype=float)


def _recrad(v) -> tuple[float, float, float]:
    """spice.recrad: range, RA in [0, 2pi), Dec."""
    x, y, z = float(v[0]), float(v[1]), float(v[2])
    big = max(abs(x), abs(y), abs(z))
    if big == 0.0:
        return 0.0, 0.0, 0.0
    xs, ys, zs = x / big, y / big, z / big
    r = big * math.sqrt(xs * xs + ys * ys + zs * zs)
    dec = math.atan2(zs, math.sqrt(xs * xs + ys * ys))
    ra = 0.0 if (xs == 0.0 and ys == 0.0) else math.atan2(ys, xs)
    if ra < 0.0:
        ra += 2.0 * math.pi
    return r, ra, dec


def _radrec(r: float, ra: float, dec: float) -> np.ndarray:
    return np.array([r * math.cos(ra) * math.cos(dec), r * math.sin(ra) * math.cos(dec),
                     r * math.sin(dec)])


def _surfpt(o: np.ndarray, u: np.ndarray, a: float, b: float, c: float):
    """Near intersection of the ray o + s u (s >= 0) with the ellipsoid a,b,c, or
    None (spice.surfpt).  Worked in the space where the ellipsoid is the unit sphere through the
    foot of the perpendicular from the centre to the ray: the textbook quadratic in s subtracts
    two numbers of size (distance / radius)^2 and loses ~8 digits for an observer 10^4 radii away
    (metres on the sub-observer point)."""
    radii = np.array([a, b, c], dtype=float)
    os_ = o / radii
    us = u / radii
    us = us / math.sqrt(float(us @ us))
    along = float(os_ @ us)
    perp = os_ - along * us
    pm2 = float(perp @ perp)
    if pm2 > 1.0:
        return None
    half_chord = math.sqrt(1.0 - pm2)
    if -along - half_chord < 0.0:
        return None
    return (perp - half_chord * us) * radii


def pgrrec(lon: float, lat: float, alt: float, re: float, f: float,
           lon_sign: float) -> np.ndarray:
    """spice.pgrrec / georec: planetographic (radians) -> body-fixed rectangular."""
    lam = lon_sign * lon
    rp = re - f * re
    clat, slat = math.cos(lat), math.sin(lat)
    clon, slon = math.cos(lam), math.sin(lam)
    big = max(abs(re * clat), abs(rp * slat))
    x = re * clat / big
    y = rp * slat / big
    scale = 1.0 / math.sqrt(x * x + y * y)
    height = alt
    base = np.array([scale * clon * x * re, scale * slon * x * re, scale * y * rp])
    normal = np.array([clat * clon, clat * slon, slat])
    return base + height * normal


@dataclass
class BodyConstants:
    """Everything about a (target, epoch, observer) triple the kernels need that
    does not depend on the image disc parameters."""

    target: str
    target_id: int
    observer: str
    utc: str
    et: float
    clight: float
    lt0: float
    P0: np.ndarray
    VT: np.ndarray
    AT: np.ndarray
    VO: np.ndarray
    S0: np.ndarray
    VS: np.ndarray
    lts0: float
    R0: np.ndarray
    omega: np.ndarray
    radii: np.ndarray
    prograde: bool
    positive_longitude_direction: str
    sub_t: np.ndarray
    sub_ray: np.ndarray
    sub_obs: np.ndarray
    sub_et: float
    sub_dist: float
    ring_n: np.ndarray
    ring_c: float
    sun_lon_lst: float
    M: np.ndarray
    target_ra: float
    target_dec: float
    target_distance: float
    target_diameter_arcsec: float
    km_per_arcsec: float
    north_pole_angle: float
    subpoint_lon: float = math.nan
    subpoint_lat: float = math.nan
    subsol_lon: float = math.nan
    subsol_lat: float = math.nan
    extra: dict = field(default_factory=dict)

    @property
    def t_ref(self) -> float:
        return self.et - self.lt0

    @property
    def r_eq(self) -> float:
        return float(self.radii[0])

    @property
    def r_polar(self) -> float:
        return float(self.radii[2])

    @property
    def lon_sign(self) -> float:
        return -1.0 if self.positive_longitude_direction == 'W' else 1.0

    def to_json_dict(self) -> dict:
        out = {}
        for k, v in self.__dict__.items():
            if k.startswith('_'):
                continue   # per-instance caches (_packed_constants)
            if isinstance(v, np.ndarray):
                out[k] = [float.hex(float(x)) for x in v.ravel()]
            elif isinstance(v, float):
                out[k] = float.hex(v)
            elif isinstance(v, (bool, int, str, dict)):
                out[k] = v
            else:
                out[k] = v
        return out

    @classmethod
    def from_json_dict(cls, d: dict) -> 'BodyConstants':
        shapes = {'R0': (3, 3), 'M': (3, 3)}
        kw = {}
        for k, v in d.items():
            if isinstance(v, list):
                arr = np.array([float.fromhex(x) for x in v])
                kw[k] = arr.reshape(shapes.get(k, arr.shape))
            elif isinstance(v, str) and (v.startswith('0x') or v.startswith('-0x')
                                         or v in ('nan', 'inf', '-inf')):
                kw[k] = float.fromhex(v)
            else:
                kw[k] = v
        return cls(**kw)


def _light_time_state(provider, target_id: int, et: float, obs_pos: np.ndarray):
    """Converged-Newtonian ('CN') light time from target to a fixed observer
    position; returns (target_state_at_et_minus_lt, lt)."""
    c = provider.clight()
    lt = 0.0
    state = provider.ssb_state(target_id, et)
    for _ in range(12):
        new_lt = float(np.linalg.norm(state[:3] - obs_pos)) / c
        state = provider.ssb_state(target_id, et - new_lt)
        if new_lt == lt:
            break
        lt = new_lt
    return state, lt


def targvec2obsvec(bc_like, provider, targvec: np.ndarray) -> np.ndarray:
    """Host scalar copy of Body._targvec2obsvec (planetmapper/body.py:917-948) using
    the provider's exact orientation; only used while building frame constants."""
    off = targvec - bc_like['sub_t']
    dist_offset = float(np.linalg.norm(bc_like['sub_ray'] + off)) - bc_like['sub_dist']
    tt = bc_like['sub_et'] - dist_offset / provider.clight()
    rmat, _ = provider.orientation(bc_like['target_id'], tt)
    return bc_like['sub_obs'] + rmat.T @ off


def obsvec2angular(M: np.ndarray, obsvec: np.ndarray) -> tuple[float, float]:
    """Body._obsvec2angular (planetmapper/body.py:1345-1361), arcseconds."""
    v = M @ obsvec
    _, x, y = _recrad(v)
    x = (-math.degrees(x)) % 360.0
    if x > 180.0:
        x -= 360.0
    return x * 3600.0, math.degrees(y) * 3600.0


def obsvec2angular_matrix(origin_ra: float, origin_dec: float, coordinate_rotation: float = 0.0) -> np.ndarray:
    """Body._get_obsvec2angular_matrix (planetmapper/body.py:1318-1343): rotation taking observer-frame
    vectors to the angular system centred on (origin_ra, origin_dec) [degrees], rotated by
    ``coordinate_rotation`` degrees."""
    origin = _radrec(1.0, math.radians(origin_ra), math.radians(origin_dec))
    _, ra_angle, _ = _recrad(origin)
    ra_matrix = _rot_axis(ra_angle, 3)
    _, _, dec_angle = _recrad(ra_matrix @ origin)
    dec_matrix = _rot_axis(-dec_angle, 2)
    return _rot_axis(math.radians(coordinate_rotation), 1) @ dec_matrix @ ra_matrix


def km2angular_matrix(bc: 'BodyConstants') -> np.ndarray:
    """Body._get_km2angular_matrix (planetmapper/body.py:1625-1634)."""
    return (1.0 / bc.km_per_arcsec) * rotation_matrix_radians(np.deg2rad(bc.north_pole_angle))


def build_body_constants(provider, target, utc: str | None, observer='EARTH', *,
                         et: float | None = None,
                         observer_state: np.ndarray | None = None,
                         with_subsol: bool = True) -> BodyConstants:
    """Derive BodyConstants. ``observer_state`` (6-vector, SSB J2000 at et) overrides
    the provider's observer ephemeris (used for observers whose SPK type the
    provider cannot read, e.g. HST's type-10 TLE segment)."""
    target_id = provider.bods2c(target)
    c = provider.clight()
    if et is None:
        et = provider.utc2et(utc)
    et = float(et)
    if observer_state is None:
        obs_state = provider.ssb_state(provider.bods2c(observer), et)
    else:
        obs_state = np.asarray(observer_state, dtype=float)
    obs_pos, obs_vel = obs_state[:3], obs_state[3:]

    # spkezr(target, et, 'J2000', 'CN', observer)  (base.py:828-837)
    tstate, lt0 = _light_time_state(provider, target_id, et, obs_pos)
    t_ref = et - lt0
    tstate = provider.ssb_state(target_id, t_ref)
    P0 = tstate[:3] - obs_pos
    VT = tstate[3:].copy()
    h_acc = 16.0  # s; central difference of the ephemeris velocity
    AT = (provider.ssb_state(target_id, t_ref + h_acc)[3:]
          - provider.ssb_state(target_id, t_ref - h_acc)[3:]) / (2.0 * h_acc)
    R0, omega = provider.orientation(target_id, t_ref)

    radii = provider.bodvar(target_id, 'RADII')[:3].astype(float)
    pm = provider.bodvar(target_id, 'PM')
    prograde = bool(pm[1] >= 0)
    lon_dir = 'W' if (prograde and target_id not in (10, 301, 399)) else 'E'
    lon_sign = -1.0 if lon_dir == 'W' else 1.0
    re, rp = float(radii[0]), float(radii[2])
    flat = (re - rp) / re

    # subpnt 'INTERCEPT/ELLIPSOID', 'CN' (body.py:538-546): intercept of the ray from
    # the observer towards the target centre, with the light time iterated on the
    # intercept point.
    lt = lt0
    sub_t = None
    for _ in range(12):
        t = et - lt
        ts = provider.ssb_state(target_id, t)
        rmat, _ = provider.orientation(target_id, t)
        o = rmat @ (obs_pos - ts[:3])  # observer wrt target centre, body-fixed
        u = -o / np.linalg.norm(o)
        p = _surfpt(o, u, *radii)
        if p is None:
            raise RuntimeError('observer is inside the target ellipsoid')
        new_lt = float(np.linalg.norm(p - o)) / c
        sub_t, sub_ray, sub_et, sub_rmat = p, p - o, t, rmat
        if abs(new_lt - lt) <= 1e-17 * abs(t):
            lt = new_lt
            break
        lt = new_lt
    sub_et = et - lt
    ts = provider.ssb_state(target_id, sub_et)
    sub_rmat, _ = provider.orientation(target_id, sub_et)
    o = sub_rmat @ (obs_pos - ts[:3])
    u = -o / np.linalg.norm(o)
    sub_t = _surfpt(o, u, *radii)
    sub_ray = sub_t - o
    sub_dist = float(np.linalg.norm(sub_ray))
    sub_obs = sub_rmat.T @ sub_ray  # _rayvec2obsvec (body.py:950-961)

    # Sun as seen from the target centre at t_ref, converged light time
    sun_state, lts0 = _light_time_state(provider, 10, t_ref, tstate[:3])
    sun_state = provider.ssb_state(10, t_ref - lts0)
    S0 = sun_state[:3] - tstate[:3]
    VS = sun_state[3:].copy()

    # et2lst Sun longitude: Sun wrt body at et_l = et - lt0 with 'LT+S', expressed in
    # the body frame evaluated at et_l (body.py:2364-2374)
    et_l = t_ref
    body_l = tstate
    sun_geo = provider.ssb_state(10, et_l)
    lt_s = float(np.linalg.norm(sun_geo[:3] - body_l[:3])) / c
    sun_lt = provider.ssb_state(10, et_l - lt_s)
    pobj = sun_lt[:3] - body_l[:3]
    vobs = body_l[3:]
    uvec = pobj / np.linalg.norm(pobj)
    vbyc = vobs / c
    h = np.cross(uvec, vbyc)
    sinphi = float(np.linalg.norm(h))
    if sinphi != 0.0:
        phi = math.asin(sinphi)
        k = h / sinphi
        app = (pobj * math.cos(phi) + np.cross(k, pobj) * math.sin(phi)
               + k * float(k @ pobj) * (1.0 - math.cos(phi)))
    else:
        app = pobj
    q = R0 @ app
    sun_lon_lst = math.atan2(q[1], q[0])

    # target RA/Dec & obsvec -> angular matrix (body.py:1318-1343, defaults)
    _, ra, dec = _recrad(P0)
    target_ra, target_dec = math.degrees(ra), math.degrees(dec)
    M = obsvec2angular_matrix(target_ra, target_dec, 0.0)

    target_distance = lt0 * c
    target_diameter_arcsec = float(
        2.0 * 60.0 * 60.0 * np.rad2deg(np.arcsin(re / target_distance)))
    km_per_arcsec = (2.0 * re) / target_diameter_arcsec

    partial = dict(target_id=target_id, sub_t=sub_t, sub_ray=sub_ray, sub_obs=sub_obs,
                   sub_et=sub_et, sub_dist=sub_dist)

    # ring plane (body.py:583-588): nvp2pl(normal, point)
    np_targvec = pgrrec(0.0, math.radians(90.0), 0.0, re, flat, lon_sign)
    np_obsvec = targvec2obsvec(partial, provider, np_targvec)
    normal = np_obsvec - P0
    normal = normal / np.linalg.norm(normal)
    const = float(normal @ P0)
    if const < 0.0:
        const, normal = -const, -normal

    # north pole angle (body.py:2998-3007)
    np_x, np_y = obsvec2angular(M, np_obsvec)
    tx, ty = obsvec2angular(M, _radrec(1.0, math.radians(target_ra),
                                       math.radians(target_dec)))
    theta = -math.atan2(tx - np_x, np_y - ty)
    theta = math.degrees(theta) % 360.0
    if theta > 180.0:
        theta -= 360.0

    # sub-observer planetographic lon/lat for metadata (body.py:547-549)
    sp_lon = math.degrees(math.atan2(sub_t[1], sub_t[0]) * lon_sign) % 360.0
    sp_lat = math.degrees(math.atan2(sub_t[2] / ((1 - flat) ** 2),
                                     math.hypot(sub_t[0], sub_t[1])))

    # sub-solar point for metadata: spice.subslr('INTERCEPT/ELLIPSOID', ..., 'CN') at
    # body.py:559-567 - the point where the Sun -> target-centre line meets the surface, at
    # the epoch et - lt with lt the light time from that point to the observer
    ss_lon = ss_lat = math.nan
    if target_id != 10 and with_subsol:   # header metadata only: a time series skips it
        lt = lt0
        for _ in range(12):
            t = et - lt
            ts = provider.ssb_state(target_id, t)
            rmat, _ = provider.orientation(target_id, t)
            sun_t, _ = _light_time_state(provider, 10, t, ts[:3])
            s_b = rmat @ (sun_t[:3] - ts[:3])
            p = _surfpt(s_b, -s_b / np.linalg.norm(s_b), *radii)
            o = rmat @ (obs_pos - ts[:3])
            new_lt = float(np.linalg.norm(p - o)) / c
            done = abs(new_lt - lt) <= 1e-17 * abs(t)
            lt = new_lt
            if done:
                break
        ss_lon = math.degrees(math.atan2(p[1], p[0]) * lon_sign) % 360.0
        ss_lat = math.degrees(math.atan2(p[2] / ((1 - flat) ** 2), math.hypot(p[0], p[1])))

    return BodyConstants(
        target=provider.bodc2n(target_id), target_id=target_id,
        observer=str(observer).upper(), utc=str(utc), et=et, clight=c, lt0=lt0,
        P0=P0, VT=VT, AT=AT, VO=obs_vel.copy(), S0=S0, VS=VS, lts0=lts0, R0=R0, omega=omega,
        radii=radii, prograde=prograde, positive_longitude_direction=lon_dir,
        sub_t=sub_t, sub_ray=sub_ray, sub_obs=sub_obs, sub_et=sub_et,
        sub_dist=sub_dist, ring_n=normal, ring_c=const, sun_lon_lst=sun_lon_lst, M=M,
        target_ra=target_ra, target_dec=target_dec, target_distance=target_distance,
        target_diameter_arcsec=target_diameter_arcsec, km_per_arcsec=km_per_arcsec,
        north_pole_angle=theta, subpoint_lon=sp_lon, subpoint_lat=sp_lat,
        subsol_lon=ss_lon, subsol_lat=ss_lat,
    )


def rotation_matrix_radians(theta: float) -> np.ndarray:
    """SpiceBase._rotation_matrix_radians (planetmapper/base.py:684-687)."""
    return np.array([[np.cos(theta), np.sin(theta)], [-np.sin(theta), np.cos(theta)]])


def xy2angular_matrix(bc: BodyConstants, x0: float, y0: float, r0: float,
                      rotation_radians: float) -> np.ndarray:
    """BodyXY._get_xy2angular_matrix (planetmapper/body_xy.py:355-369)."""
    s = bc.target_diameter_arcsec / (2 * r0)
    m2 = s * rotation_matrix_radians(-rotation_radians)
    offset = -m2.dot(np.array([x0, y0]))
    m3 = np.identity(3)
    m3[:2, :2] = m2
    m3[:2, 2] = offset
    return m3


def _pack_constants(bc: BodyConstants, alt: float) -> np.ndarray:
    """The disc-independent part of the PMFrame array (everything but A, Ainv, nx, ny, x0, y0, r_cut2,
    optimize_speed).  Cached on the BodyConstants instance per altitude: a BodyXY whose disc parameters
    change (or many BodyXY sharing one set of constants) re-packs only the disc fields."""
    cache = bc.__dict__.setdefault('_packed_constants', {})
    buf = cache.get(alt)
    if buf is not None:
        return buf
    buf = np.zeros(PMFRAME_NDOUBLES, dtype=np.float64)

    def put(name, value):
        o, n = PMFRAME_OFFSETS[name]
        buf[o : o + n] = np.asarray(value, dtype=np.float64).ravel()

    radii = bc.radii + alt
    re, rp = float(radii[0]), float(radii[2])
    put('et', bc.et)
    put('clight', bc.clight)
    put('lt0', bc.lt0)
    put('t_ref', bc.t_ref)
    put('P0', bc.P0)
    put('VT', bc.VT)
    put('AT', bc.AT)
    put('VO', bc.VO)
    put('S0', bc.S0)
    put('VS', bc.VS)
    put('lts0', bc.lts0)
    put('R0', bc.R0)
    put('omega', bc.omega)
    put('radii', radii)
    put('re', re)
    put('f', (re - rp) / re)
    put('lon_sign', bc.lon_sign)
    put('prograde', 1.0 if bc.prograde else 0.0)
    put('sub_t', bc.sub_t)
    put('sub_ray', bc.sub_ray)
    put('sub_obs', bc.sub_obs)
    put('sub_dt', bc.sub_et - bc.t_ref)
    put('sub_dist', bc.sub_dist)
    put('ring_n', bc.ring_n)
    put('ring_c', bc.ring_c)
    put('sun_lon_lst', bc.sun_lon_lst)
    put('M', bc.M)
    put('ang2km', np.linalg.inv(km2angular_matrix(bc)))
    put('km_per_arcsec', bc.km_per_arcsec)
    put('r_eq', re)
    buf.setflags(write=False)
    if len(cache) < 64:   # altitudes are user input: do not grow without bound
        cache[alt] = buf
    return buf


def pack_frame(bc: BodyConstants, *, nx: int, ny: int, x0: float, y0: float,
               r0: float, rotation_radians: float, alt: float = 0.0,
               optimize_speed: bool = True) -> np.ndarray:
    """Pack BodyConstants + disc parameters into the flat PMFrame double array.

    ``alt`` reproduces _AdjustedSurfaceAltitude (planetmapper/body.py:172-229): the
    radii handed to the per-pixel code grow by alt; nothing computed at construction
    (sub-point, ring plane, plate scale) changes.
    """
    buf = _pack_constants(bc, float(alt)).copy()
    off = PMFRAME_OFFSETS
    a3 = xy2angular_matrix(bc, x0, y0, r0, rotation_radians)
    a3inv = np.linalg.inv(a3)
    o = off['A'][0]
    buf[o : o + 6] = a3[:2, :].ravel()
    o = off['Ainv'][0]
    buf[o : o + 6] = a3inv[:2, :].ravel()
    buf[off['nx'][0]] = nx
    buf[off['ny'][0]] = ny
    buf[off['x0'][0]] = x0
    buf[off['y0'][0]] = y0
    radii = buf[off['radii'][0] : off['radii'][0] + 3]
    r_cut = (r0 * float(max(radii)) / float(radii[0])) * 1.05 + 1  # body_xy.py:3190-3202
    buf[off['r_cut2'][0]] = r_cut**2
    buf[off['optimize_speed'][0]] = 1.0 if optimize_speed else 0.0
    return buf


def frame_field(buf: np.ndarray, name: str) -> np.ndarray:
    o, n = PMFRAME_OFFSETS[name]
    return buf[..., o : o + n]
