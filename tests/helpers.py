"""Shared comparison helpers for the parity tests."""
import numpy as np

PLANE_NAMES = [
    'LON-GRAPHIC', 'LAT-GRAPHIC', 'LON-CENTRIC', 'LAT-CENTRIC', 'RA', 'DEC', 'PIXEL-X',
    'PIXEL-Y', 'KM-X', 'KM-Y', 'ANGULAR-X', 'ANGULAR-Y', 'PHASE', 'INCIDENCE',
    'EMISSION', 'AZIMUTH', 'LOCAL-SOLAR-TIME', 'DISTANCE', 'RADIAL-VELOCITY', 'DOPPLER',
    'LIMB-DISTANCE', 'LIMB-LON-GRAPHIC', 'LIMB-LAT-GRAPHIC', 'RING-RADIUS',
    'RING-LON-GRAPHIC', 'RING-DISTANCE',
]
PID = {n: i for i, n in enumerate(PLANE_NAMES)}
ANGLE_PLANES = ['LON-GRAPHIC', 'LAT-GRAPHIC', 'LON-CENTRIC', 'LAT-CENTRIC', 'PHASE', 'INCIDENCE',
                'EMISSION', 'AZIMUTH']


def angle_diff(a, b):
    """|a - b| in degrees, treating 0 / 360 as the same longitude."""
    d = np.abs(a - b)
    return np.minimum(d, np.abs(360.0 - d))


def max_diff(a, b, wrap=False):
    ok = np.isfinite(a) & np.isfinite(b)
    if not ok.any():
        return 0.0
    d = angle_diff(a[ok], b[ok]) if wrap else np.abs(a[ok] - b[ok])
    return float(d.max())


def masks_equal(a, b, exclude=None):
    """NaN masks identical outside `exclude`; returns (ok, n_mismatch, n_excluded)."""
    mism = np.isnan(a) != np.isnan(b)
    n_ex = 0
    if exclude is not None:
        n_ex = int((mism & exclude).sum())
        mism = mism & ~exclude
    return (not mism.any()), int(mism.sum()), n_ex


def lst_boundary(lon_deg, frame_lon_sign, sun_lon, prograde=True, tol_s=1e-6):
    """True where the local-solar-time second count is within tol_s of an integer
    boundary (LOCAL-SOLAR-TIME is quantised to whole seconds: body.py:2397-2398)."""
    ang = frame_lon_sign * np.deg2rad(lon_deg) - sun_lon
    if not prograde:
        ang = -ang
    ang = np.mod(ang, 2 * np.pi)
    sec = ang * (86400.0 / (2 * np.pi)) + 43200.0
    frac = sec - np.floor(sec)
    return (frac < tol_s) | (frac > 1 - tol_s)


# ---------------------------------------------------------------------------------
# Conditioning-aware tolerances for the image-direction surface planes.
#
# north_star's bars are 1e-9 deg (angles) and 1e-12 relative (distances, velocities).
# The reference itself cannot deliver that near the limb: CSPICE forms the observer
# position in the body frame with magnitude |P0| ~ 8e8 km, i.e. with a rounding
# granularity ulp(|P0|) ~ 1.2e-7 km, and a perpendicular error d of the ray moves the
# intercept by d / cos(emission) along the surface.  Two correct FP64 evaluations
# (different operation order, FMA contraction) therefore differ by about
# 2 ulp(|P0|) / (r cos e) radians.  The tolerances below are the stated bars, widened
# ONLY by that amplification factor kappa = 1 / cos(emission) where it exceeds them.
# ---------------------------------------------------------------------------------
def surface_tolerances(ref_planes, p0_norm, r_min, omega_norm):
    emi = ref_planes[PID['EMISSION']]
    latc = ref_planes[PID['LAT-CENTRIC']]
    with np.errstate(invalid='ignore', divide='ignore'):
        kappa = 1.0 / np.maximum(np.cos(np.deg2rad(emi)), 1e-12)
        delta = 2.0 * np.spacing(p0_norm)                 # km, positional rounding noise
        ang = np.rad2deg(delta * kappa / r_min)           # deg, induced angle noise
        base = np.maximum(1e-9, 4.0 * ang)
        coslat = np.maximum(np.cos(np.deg2rad(latc)), 1e-12)
    tol = {
        'LAT-GRAPHIC': base, 'LAT-CENTRIC': base, 'INCIDENCE': base, 'EMISSION': base,
        'LON-GRAPHIC': base / coslat, 'LON-CENTRIC': base / coslat,
        'PHASE': np.full_like(base, 1e-9),
        'DISTANCE': np.full_like(base, 1e-12 * p0_norm),
        'RADIAL-VELOCITY': 1e-12 * np.abs(ref_planes[PID['RADIAL-VELOCITY']]) + 4.0 * omega_norm * delta * kappa + 1e-13,
    }
    tol['DOPPLER'] = tol['RADIAL-VELOCITY'] / 299792.458 + 4e-16
    # AZIMUTH is a float64 formula of (phase, incidence, emission) that is itself
    # ill-conditioned near the sub-observer / sub-solar points: propagate numerically
    g, i, e = (np.deg2rad(ref_planes[PID[n]]) for n in ('PHASE', 'INCIDENCE', 'EMISSION'))
    t = np.deg2rad(base)

    def az(g, i, e):
        with np.errstate(invalid='ignore', divide='ignore'):
            a = np.cos(g) - np.cos(e) * np.cos(i)
            b = np.sqrt(1.0 - np.cos(e) ** 2) * np.sqrt(1.0 - np.cos(i) ** 2)
            return np.rad2deg(np.pi - np.arccos(np.clip(a / b, -1, 1)))

    az0 = az(g, i, e)
    dev = np.zeros_like(base)
    for sg in (-1, 1):
        for si in (-1, 1):
            for se in (-1, 1):
                with np.errstate(invalid='ignore'):
                    dev = np.fmax(dev, np.abs(az(g + sg * np.deg2rad(1e-9), i + si * t, e + se * t) - az0))
    tol['AZIMUTH'] = 2.0 * dev + 1e-9
    return tol, kappa
