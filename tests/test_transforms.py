"""
The coordinate-pair transforms of Body / BodyXY beyond xy <-> lon / lat (SURVEY.md section 2 "xy <->
other coordinate transforms"; planetmapper/body.py:1083-1900, planetmapper/body_xy.py:385-676):

  * the ORACLE's restatement (oracle/pm_oracle.c::transform_one) is pinned on the 16-digit literals of the
    reference's own tests (tests/test_body.py:675-720, :864-983, :1078-1140, :1142-1170, :1357-1420,
    :1536-1575) - CPU;
  * the kernels' per-point code against the oracle for every pair: host instantiation (CPU) and pm_transform
    through the C ABI (GPU), plus the public methods on the same literals (GPU).
"""
import ctypes
import itertools

import numpy as np
import pytest

from helpers import angle_diff, img_case
from planetmapper_b200 import frame as F

nan = np.nan
SYSTEMS = ['xy', 'angular', 'km', 'radec', 'lonlat']


def _aux(bc, **kw):
    ra = bc.target_ra if kw.get('origin_ra') is None else kw['origin_ra']
    dec = bc.target_dec if kw.get('origin_dec') is None else kw['origin_dec']
    m = F.obsvec2angular_matrix(ra, dec, kw.get('coordinate_rotation', 0.0))
    return np.concatenate([m.ravel(), F.km2angular_matrix(bc).ravel()])


def _close(got, want, atol=1e-8, rtol=1e-9):
    got, want = np.asarray(got, dtype=float), np.asarray(want, dtype=float)
    return np.allclose(got, want, atol=atol, rtol=rtol, equal_nan=True)


# ---- literals of the reference's tests (Body('Jupiter', '2005-01-01T00:00:00', observer='HST')) ----
LONLAT2RADEC = [((0, 90), (196.37390490466322, -5.561534444253404)),
                ((0, 0), (196.36982789576643, -5.565060944053696)),
                ((123.456, -56.789), (196.3691609381441, -5.5685956879058764)),
                ((nan, nan), (nan, nan)), ((nan, 0), (nan, nan)), ((0, nan), (nan, nan)), ((np.inf, np.inf), (nan, nan))]
LONLAT_ALT2RADEC = [((42, 23.4, 0), (196.36871162182828, -5.5624995718895915)),
                    ((42, 23.4, -123.456), (196.36871704240835, -5.562505596011716)),
                    ((42, 23.4, 1234.567), (196.3686574157507, -5.562439330354751))]
RADEC2LONLAT = [((196.37198562427025, -5.565793847134351), (153.1235185909613, -3.0887371238645795)),
                ((196.372, -5.566), (154.24480750302573, -5.475831082435726)),
                ((196.3742715121965, -5.561743939677709), (180.00086055026196, 80.00042229835671)),
                ((0, 0), (nan, nan)), ((nan, nan), (nan, nan)), ((nan, 0), (nan, nan)), ((np.inf, np.inf), (nan, nan))]
RADEC_ALT2LONLAT = [((196.37198562427025, -5.565793847134351, 123456.789), (153.12766781084477, -2.834663828028037)),
                    ((196.37198562427025, -5.565793847134351, -1000), (153.12348498172653, -3.0948138787454225))]
RADEC_CENTRIC_ALT = (([196.36982417, 196.37008339, 196.37670856, 196.37344939, 196.37390845],
                      [-5.56505968, -5.5646989, -5.565273, -5.57029209, -5.56152658]), 123.44,
                     ([-126.12884032, -127.36719725, 137.03419591, 130.7518092, -153.12075396],
                      [-4.81959108, 0.13509603, 28.23425081, -47.39154262, 83.81857224]))
ANGULAR2RADEC = [((0, 0), {}, (196.37198562131056, -5.565793839734843)),
                 ((0, 0), {'coordinate_rotation': 123}, (196.37198562131056, -5.565793839734843)),
                 ((1.234, 5.678), {}, (196.37164122076928, -5.564216617412704)),
                 ((-3600.1234, 45678), {}, (197.35518558863563, 7.1233716685998285)),
                 ((1.234, 5.678), {'coordinate_rotation': 123}, (196.3708441579451, -5.566940333059796)),
                 ((1.234, 5.678), {'origin_ra': 123}, (122.99965559945868, -5.564216624812211)),
                 ((1.234, 5.678), {'origin_dec': 12.3}, (196.37163479126497, 12.301577221998656)),
                 ((1.234, 5.678), {'origin_ra': -123, 'origin_dec': -12.3}, (236.99964917120613, -12.298422777554215)),
                 ((1.234, 5.678), {'origin_ra': -123, 'origin_dec': 12.3, 'coordinate_rotation': -123},
                  (237.001544919471, 12.299428456509167))]
ANGULAR2LONLAT = [((0, 0), {}, (153.12351859061235, -3.0887371240013572)),
                  ((0, 0), {'coordinate_rotation': 123}, (153.12351859061235, -3.0887371240013572)),
                  ((1.234, 5.678), {}, (141.76181779277195, 14.187903497915688)),
                  ((-3600.1234, 45678), {}, (nan, nan)),
                  ((1.234, 5.678), {'coordinate_rotation': 123}, (146.10317442767905, -23.08048248991215)),
                  ((1.234, 5.678), {'origin_ra': 196.372, 'origin_dec': -5.566}, (143.01960641488623, 11.717675615612585)),
                  ((1.234, 0.678), {'origin_ra': 196.372, 'origin_dec': -5.566, 'coordinate_rotation': -123},
                   (156.98171972231182, -1.4107148298315533))]
KM2RADEC = [((0, 0), (196.3719856242702, -5.56579384713435)), ((99999, 99999), (196.36845127590436, -5.556555100442686)),
            ((1234, -5678), (196.37174335301282, -5.566120708196197)),
            ((-0.1234, 9999.5678), (196.37227302705824, -5.565156047930656))]
KM2LONLAT = [((0, 0), (153.12351859061235, -3.0887371240013572)), ((123, 456.789), (153.02485721448028, -2.6703253305682195)),
             ((-500, -200), (153.52477375354786, -3.2718421646109985)), ((5000, 50001), (147.39408652731262, 47.4410279733397))]
ANGULAR2KM = [((0, 0), {}, (0.0, 0.0)), ((0, 0), {'coordinate_rotation': 123}, (0.0, 0.0)),
              ((1.234, 5.678), {}, (13707.106875939699, 18580.59989529313)),
              ((-3600.1234, 45678), {}, (61222909.71285939, 171472523.56580824)),
              ((1.234, 5.678), {'coordinate_rotation': 123}, (8117.576807789242, -21615.467104869596)),
              ((1.234, 5.678), {'origin_ra': 123}, (928803175.7862874, -478472263.2296324)),
              ((1.234, 5.678), {'origin_dec': 12.3}, (104598412.22915992, 233217325.082532)),
              ((1.234, 5.678), {'origin_ra': -123, 'origin_dec': -12.3}, (-569001780.3607075, 128938234.54185842)),
              ((1.234, 5.678), {'origin_ra': -123, 'origin_dec': 12.3, 'coordinate_rotation': -123},
               (-446038232.73474604, 458652497.8006319))]


def _literal_cases(bc):
    """(src, dst, a, b, kwargs for oracle.transform incl. alt / aux, frame altitude, expected, atol)."""
    cases = []
    for (lon, lat), want in LONLAT2RADEC:
        cases.append(('lonlat', 'radec', lon, lat, dict(), 0.0, want, 1e-10))
    for (lon, lat, alt), want in LONLAT_ALT2RADEC:
        cases.append(('lonlat', 'radec', lon, lat, dict(alt=alt), 0.0, want, 1e-10))
    for (ra, dec), want in RADEC2LONLAT:
        cases.append(('radec', 'lonlat', ra, dec, dict(), 0.0, want, 1e-8))
    for (ra, dec, alt), want in RADEC_ALT2LONLAT:
        cases.append(('radec', 'lonlat', ra, dec, dict(alt=alt), alt, want, 1e-8))
    (ras, decs), alt, (lons, lats) = RADEC_CENTRIC_ALT
    for ra, dec, lon, lat in zip(ras, decs, lons, lats):
        cases.append(('radec', 'lonlat', ra, dec, dict(alt=alt, planetocentric=True), alt, (lon, lat), 6e-9))
    for (ax, ay), kw, want in ANGULAR2RADEC:
        # these literals were written by an older reference version: its target RA / Dec sit 3e-9 / 7e-9 deg from
        # the ones test_radec2lonlat (and the golden FITS header) carry, so they pin the formula at 1e-8 deg
        cases.append(('angular', 'radec', ax, ay, dict(aux13=_aux(bc, **kw)), 0.0, want, 1e-8))
        # the inverse at the reference's own atol = 1e-4 arcsec (3e-9 deg of origin shift = 1e-5 arcsec)
        cases.append(('radec', 'angular', want[0], want[1], dict(aux13=_aux(bc, **kw)), 0.0, (ax, ay), 1e-4))
    for (ax, ay), kw, want in ANGULAR2LONLAT:
        # asserted at atol = 1e-3 by the reference (test_body.py:1174-1179): older numbers, like the ones above
        cases.append(('angular', 'lonlat', ax, ay, dict(aux13=_aux(bc, **kw)), 0.0, want, 1e-3))
        if np.isfinite(want[0]):
            cases.append(('lonlat', 'angular', want[0], want[1], dict(aux13=_aux(bc, **kw), not_visible_nan=True),
                          0.0, (ax, ay), 1e-4))
    for (kx, ky), want in KM2RADEC:
        cases.append(('km', 'radec', kx, ky, dict(), 0.0, want, 1e-10))
        cases.append(('radec', 'km', want[0], want[1], dict(), 0.0, (kx, ky), 1e-3))
    for (kx, ky), want in KM2LONLAT:
        cases.append(('km', 'lonlat', kx, ky, dict(), 0.0, want, 1e-3))   # np.allclose defaults in the reference
        # the reference's lon / lat -> observer direction is not the exact inverse of its ray cast (body.py:917-948
        # dates the point with the sub-observer light time): np.allclose(..., atol=1e-3) with rtol = 1e-5 there
        cases.append(('lonlat', 'km', want[0], want[1], dict(not_visible_nan=True), 0.0, (kx, ky),
                      1e-3 + 1e-5 * max(abs(kx), abs(ky))))
    for (ax, ay), kw, want in ANGULAR2KM:
        cases.append(('angular', 'km', ax, ay, dict(aux13=_aux(bc, **kw)), 0.0, want, 1e-3 + 1e-5 * abs(want[0])))
    return cases


def test_oracle_transforms_match_reference_literals(oracle, bc_hst):
    for src, dst, a, b, kw, frame_alt, want, atol in _literal_cases(bc_hst):
        fr = img_case(bc_hst, 15, 10, 5, 8, 3, 45, alt=frame_alt)
        ga, gb, _ = oracle.transform(fr, src, dst, a, b, **kw)
        got = (float(ga.reshape(())), float(gb.reshape(())))
        for g, w in zip(got, want):
            if np.isnan(w):
                assert np.isnan(g), (src, dst, a, b, got, want)
            else:
                d = angle_diff(g, w) if dst in ('radec', 'lonlat') else abs(g - w)
                assert d <= atol, (src, dst, a, b, kw.keys(), got, want)


def test_oracle_transform_conventions(oracle, bc_hst):
    """graphic <-> centric round trips and the xy pairs agree with the dedicated oracle entry points."""
    fr = img_case(bc_hst, 15, 10, 5, 8, 3, 45)
    rng = np.random.default_rng(0)
    lon, lat = rng.uniform(0, 360, 200), rng.uniform(-89, 89, 200)
    for alt in (0.0, 1234.5):
        lc, bc_, _ = oracle.transform(fr, 'lonlat', 'centric', lon, lat, alt=alt)
        lg, bg, _ = oracle.transform(fr, 'lonlat', 'lonlat', lc, bc_, alt=0.0, planetocentric=True)
        if alt == 0.0:   # centric2graphic is the inverse of graphic2centric on the surface
            assert np.max(angle_diff(lg, lon)) < 1e-9 and np.max(np.abs(bg - lat)) < 1e-9
        assert np.all(np.abs(bc_) <= np.abs(lat) + 1e-12)    # oblate body: centric latitude is the smaller one
    x, y = rng.uniform(0, 14, 300), rng.uniform(0, 9, 300)
    l1, b1, m1 = oracle.xy2lonlat(fr, x, y)
    l2, b2, m2 = oracle.transform(fr, 'xy', 'lonlat', x, y)
    assert m1 == m2 and np.array_equal(l1, l2, equal_nan=True) and np.array_equal(b1, b2, equal_nan=True)
    x1, y1 = oracle.lonlat2xy(fr, lon, lat, not_visible_nan=True)
    x2, y2, _ = oracle.transform(fr, 'lonlat', 'xy', lon, lat, not_visible_nan=True)
    assert np.array_equal(x1, x2, equal_nan=True) and np.array_equal(y1, y2, equal_nan=True)


def _random_inputs(rng, bc, src, n):
    if src == 'xy':
        a, b = rng.uniform(0, 10, n), rng.uniform(3, 13, n)
    elif src == 'angular':
        a, b = rng.uniform(-30, 30, n), rng.uniform(-30, 30, n)
    elif src == 'km':
        a, b = rng.uniform(-1.2e5, 1.2e5, n), rng.uniform(-1.2e5, 1.2e5, n)
    elif src == 'radec':
        a = bc.target_ra + rng.uniform(-30, 30, n) / 3600 / np.cos(np.deg2rad(bc.target_dec))
        b = bc.target_dec + rng.uniform(-30, 30, n) / 3600
    else:
        a, b = rng.uniform(-360, 720, n), rng.uniform(-90, 90, n)
    a[:3], b[:3] = [np.nan, 1.0, np.inf], [1.0, np.nan, 1.0]
    return a, b


def _tolerance(dst, emission_ok=True):
    return {'xy': 1e-9 * 15, 'angular': 2e-8, 'km': 1e-4, 'radec': 1e-12, 'lonlat': 1e-9, 'centric': 1e-9}[dst]


def _compare_pair(run, oracle, bc, fr_of_alt, src, dst, rng):
    a, b = _random_inputs(rng, bc, src, 3000)
    variants = [dict()]
    if src == 'lonlat':
        variants = [dict(not_visible_nan=True), dict(not_visible_nan=False), dict(not_visible_nan=True, alt=777.7),
                    dict(not_visible_nan=True, planetocentric=True), dict(planetocentric=True, alt=-55.5)]
    elif dst == 'lonlat':
        variants = [dict(), dict(alt=4321.0), dict(planetocentric=True), dict(planetocentric=True, alt=250.0)]
    if 'angular' in (src, dst):
        variants = [dict(v, aux13=aux) for v in variants
                    for aux in (None, _aux(bc, origin_ra=bc.target_ra + 0.002, origin_dec=bc.target_dec - 0.001,
                                           coordinate_rotation=33.0))]
    for kw in variants:
        alt = kw.get('alt', 0.0)
        fr = fr_of_alt(alt if dst == 'lonlat' else 0.0)
        kw_full = dict(kw)
        if kw_full.get('aux13') is None:
            kw_full['aux13'] = _aux(bc)
        wa, wb, wm = oracle.transform(fr, src, dst, a, b, **kw)
        ga, gb, gm = run(fr, src, dst, a, b, kw_full)
        mism = np.isnan(ga) != np.isnan(wa)
        assert mism.sum() <= 2 and abs(gm - wm) <= 2, (src, dst, kw.keys(), int(mism.sum()))   # limb grazers only
        ok = np.isfinite(ga) & np.isfinite(wa)
        assert ok.sum() > 100, (src, dst, kw.keys())
        tol = _tolerance(dst)
        if dst == 'lonlat':
            # conditioning near the limb and the poles (see helpers.surface_tolerances): compare through cos(lat)
            # and only well inside the disc; the rest is covered by the mask check above
            # within 30 deg of the sub-observer point (emission < ~30 deg): 2 ulp(|P0|) / (r cos e) < 1e-9 deg there
            lo0, la0 = np.deg2rad(bc.subpoint_lon), np.deg2rad(bc.subpoint_lat)
            lo1, la1 = np.deg2rad(wa), np.deg2rad(wb)
            with np.errstate(invalid='ignore'):
                core = (np.sin(la0) * np.sin(la1) + np.cos(la0) * np.cos(la1) * np.cos(lo1 - lo0)) > np.cos(np.deg2rad(30))
            if kw.get('planetocentric'):
                core &= np.abs(wb) < 30
            # bar: the conditioning floor of a 7e4 km body 8e8 km away, 4 * 2 ulp(|P0|) / (r cos e) = 9.5e-10 deg at
            # 30 deg (tests/helpers.py::surface_tolerances), once for each of the two evaluations compared
            tol = 2e-9
            sel = ok & core
            assert sel.sum() > 30, (src, dst)
            assert np.max(angle_diff(ga[sel], wa[sel]) * np.cos(np.deg2rad(wb[sel]))) <= tol, (src, dst, kw.keys())
            assert np.max(np.abs(gb[sel] - wb[sel])) <= tol, (src, dst, kw.keys())
        else:
            da = angle_diff(ga[ok], wa[ok]) if dst == 'radec' else np.abs(ga[ok] - wa[ok])
            assert np.max(da) <= tol and np.max(np.abs(gb[ok] - wb[ok])) <= tol, (src, dst, kw.keys(), np.max(da))


PAIRS = [(s, d) for s, d in itertools.product(SYSTEMS, SYSTEMS) if s != d]


@pytest.mark.parametrize('src,dst', PAIRS)
def test_device_code_transform_pairs_vs_oracle(HC, oracle, bc_hst, src, dst):
    """Host instantiation of pm_transform's per-point code against the oracle, every pair."""
    import test_host_check as T

    def run(fr, src, dst, a, b, kw):
        f = np.ascontiguousarray(fr, dtype=np.float64)
        oa, ob = np.empty_like(a), np.empty_like(a)
        missed = ctypes.c_int64(0)
        flags = (1 if kw.get('not_visible_nan') else 0) | (4 if kw.get('planetocentric') else 0)
        aux = np.ascontiguousarray(kw['aux13'], dtype=np.float64)
        assert HC.hc_transform(T._p(f), ctypes.c_int(oracle.COORD[src]), ctypes.c_int(oracle.COORD[dst]), T._p(a),
                               T._p(b), ctypes.c_int64(a.size), ctypes.c_double(kw.get('alt', 0.0)),
                               ctypes.c_uint32(flags), T._p(aux), T._p(oa), T._p(ob), ctypes.byref(missed)) == 0
        return oa, ob, missed.value

    rng = np.random.default_rng(1000 + PAIRS.index((src, dst)))
    _compare_pair(run, oracle, bc_hst, lambda alt: img_case(bc_hst, 15, 10, 5, 8, 3, 45, alt=alt), src, dst, rng)


# ---------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope='module')
def L():
    import torch

    from planetmapper_b200 import _lib

    assert torch.cuda.is_available()
    _lib.load_library()
    return _lib


@pytest.mark.gpu
@pytest.mark.parametrize('src,dst', PAIRS)
def test_gpu_transform_pairs_vs_oracle(L, oracle, bc_hst, src, dst):
    def run(fr, src, dst, a, b, kw):
        oa, ob, missed = L.transform(L.to_device(fr), src, dst, L.to_device(a), L.to_device(b),
                                     alt=kw.get('alt', 0.0), not_visible_nan=bool(kw.get('not_visible_nan')),
                                     planetocentric=bool(kw.get('planetocentric')), aux13=kw['aux13'])
        return oa.cpu().numpy(), ob.cpu().numpy(), int(missed.item())

    rng = np.random.default_rng(1000 + PAIRS.index((src, dst)))
    _compare_pair(run, oracle, bc_hst, lambda alt: img_case(bc_hst, 15, 10, 5, 8, 3, 45, alt=alt), src, dst, rng)
    # default matrices (aux13 = None) are the frame's own
    a, b = _random_inputs(rng, bc_hst, src, 64)
    fr = img_case(bc_hst, 15, 10, 5, 8, 3, 45)
    d1 = L.transform(L.to_device(fr), src, dst, L.to_device(a), L.to_device(b), aux13=None)
    d2 = L.transform(L.to_device(fr), src, dst, L.to_device(a), L.to_device(b), aux13=_aux(bc_hst))
    for u, v in zip(d1[:2], d2[:2]):
        u, v = u.cpu().numpy(), v.cpu().numpy()
        assert np.array_equal(np.isnan(u), np.isnan(v))
        ok = np.isfinite(u)
        assert np.allclose(u[ok], v[ok], rtol=1e-12, atol=1e-9)


@pytest.mark.gpu
def test_api_transforms_on_reference_literals(bc_hst):
    """The public methods (names, keyword arguments, scalar / array behaviour, NotFoundError) on the
    literals of the reference's own tests."""
    import planetmapper_b200 as pm

    body = pm.BodyXY(constants=bc_hst, nx=15, ny=10)
    body.set_disc_params(5, 8, 3, 45)
    for (lon, lat), want in LONLAT2RADEC:
        assert _close(body.lonlat2radec(lon, lat, not_visible_nan=False), want, atol=1e-9)
    for (lon, lat, alt), want in LONLAT_ALT2RADEC:
        assert _close(body.lonlat2radec(lon, lat, alt=alt, not_visible_nan=False), want, atol=1e-9)
    for (ra, dec), want in RADEC2LONLAT:
        got = body.radec2lonlat(ra, dec)
        assert isinstance(got[0], float) and _close(got, want, atol=1e-7)
    with pytest.raises(pm.NotFoundError):
        body.radec2lonlat(0, 0, not_found_nan=False)
    for (ra, dec, alt), want in RADEC_ALT2LONLAT:
        assert _close(body.radec2lonlat(ra, dec, alt=alt), want, atol=1e-7)
    (ras, decs), alt, want = RADEC_CENTRIC_ALT
    got = body.radec2lonlat([ras], [decs], alt=alt, planetocentric=True)
    assert got[0].shape == (1, 5) and _close(got[0][0], want[0], atol=1e-7) and _close(got[1][0], want[1], atol=1e-7)
    for (ax, ay), kw, want in ANGULAR2RADEC:
        assert _close(body.angular2radec(ax, ay, **kw), want, atol=1e-8)    # older literals, see _literal_cases
        assert _close(body.radec2angular(*want, **kw), (ax, ay), atol=1e-4)
    for (ax, ay), kw, want in ANGULAR2LONLAT:
        assert _close(body.angular2lonlat(ax, ay, **kw), want, atol=1e-3)
        if np.isfinite(want[0]):
            assert _close(body.lonlat2angular(*want, **kw), (ax, ay), atol=1e-4)
        else:
            with pytest.raises(pm.NotFoundError):
                body.angular2lonlat(ax, ay, **kw, not_found_nan=False)
    for (kx, ky), want in KM2RADEC:
        assert _close(body.km2radec(kx, ky), want, atol=1e-9)
        assert _close(body.radec2km(*want), (kx, ky), atol=1e-3)
    for (kx, ky), want in KM2LONLAT:
        assert _close(body.km2lonlat(kx, ky), want, atol=1e-3)
        assert _close(body.lonlat2km(*want), (kx, ky), atol=1e-3, rtol=1e-5)
        centric = body.graphic2centric_lonlat(*want)
        assert _close(body.km2lonlat(kx, ky, planetocentric=True), centric, atol=1e-3)
        exact = body.graphic2centric_lonlat(*body.km2lonlat(kx, ky))
        assert _close(body.km2lonlat(kx, ky, planetocentric=True), exact, atol=1e-9)
        assert _close(body.lonlat2km(*centric, planetocentric=True), (kx, ky), atol=1e-3, rtol=1e-5)
        assert _close(body.centric2graphic_lonlat(*centric), want, atol=1e-9)
    for (ax, ay), kw, want in ANGULAR2KM:
        assert _close(body.angular2km(ax, ay, **kw), want, atol=1e-3, rtol=1e-5)
        assert _close(body.km2angular(*want, **kw), (ax, ay), atol=1e-3, rtol=1e-5)   # the reference's own bar
    # BodyXY pairs: consistency with the backplane images of the same frame and round trips
    x, y = np.meshgrid(np.arange(15.0), np.arange(10.0))
    ra, dec = body.xy2radec(x, y)
    assert _close(ra, body.get_backplane_img('RA'), atol=1e-12) and _close(dec, body.get_backplane_img('DEC'), atol=1e-12)
    kx, ky = body.xy2km(x, y)
    assert _close(kx, body.get_backplane_img('KM-X'), atol=1e-5) and _close(ky, body.get_backplane_img('KM-Y'), atol=1e-5)
    # the ANGULAR-X / -Y backplanes are the km planes over km_per_arcsec (north up, body_xy.py:3611-3656), not
    # the RA / Dec aligned system of xy2angular
    assert _close(kx / body.km_per_arcsec, body.get_backplane_img('ANGULAR-X'), atol=1e-8)
    ax, ay = body.xy2angular(x, y)
    ra0, dec0 = body.angular2radec(ax, ay)
    assert _close(ra0, ra, atol=1e-11) and _close(dec0, dec, atol=1e-11)
    for fwd, inv, args in ((body.xy2radec, body.radec2xy, {}), (body.xy2km, body.km2xy, {}),
                           (body.xy2angular, body.angular2xy, dict(coordinate_rotation=12.0, origin_ra=196.0))):
        u, v = fwd(x, y, **args)
        x2, y2 = inv(u, v, **args)
        assert _close(x2, x, atol=1e-6) and _close(y2, y, atol=1e-6)
    assert isinstance(body.xy2radec(1.0, 2.0)[0], float)
    assert all(np.isnan(v) for v in body.xy2radec(np.nan, 2.0))
    with pytest.raises(TypeError):
        body.xy2angular(1.0, 2.0, origin=3)
