#!/usr/bin/env python
"""
Regenerates the committed fixtures under tests/golden/ and planetmapper_b200/data/
from the reference checkout at /root/reference (present only in the authoring
container; nothing at test/bench run time reads /root/reference).

    python tests/golden/make_golden.py

Outputs
-------
tests/golden/ref_outputs.npz
    Every HDU of the reference's own golden FITS files
    (/root/reference/tests/data/outputs/*.fits, compared by the reference at
    tests/test_observation.py:1016-1280) plus the input cube
    tests/data/inputs/test.fits, as float64 arrays keyed "<file>/<EXTNAME>".
tests/golden/ref_headers.json
    The primary-header cards of those files (disc parameters, ET, light time ...).
tests/golden/ref_cards.json
    Raw 80-character header cards (primary HDU + first two extensions) and the EXTNAME order
    of a few of those files: the pin for the FITS staging's card formatting and HDU layout.
tests/golden/jupiter_hst_2005.json
    BodyConstants for the reference's main fixture, Jupiter from HST at
    2005-01-01T00:00:00 (tests/test_body_xy.py:69-76).  HST's ephemeris is an SPK
    type 10 (TLE) segment MiniSpice does not read, so the observer position is
    back-derived from the golden header (TARGET RA / DEC, LIGHT-TIME) and the
    observer velocity is SOLVED by least squares from the golden RADIAL-VELOCITY
    plane (documented as derived, not measured).
planetmapper_b200/data/ephem_extract.npz, pck_pool.json
    The Chebyshev records (SPK types 2/3) and PCK constants MiniSpice needs for the
    bodies and epoch windows used by tests and bench.py, cut from the reference's
    bundled test kernels (tests/data/kernels), so frames can be built on machines
    that have neither spiceypy nor the kernel files.
"""
import json
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from fits_min import read_cards, read_fits  # noqa: E402
from oracle import oracle as O  # noqa: E402
from planetmapper_b200 import frame as F  # noqa: E402
from planetmapper_b200.minispice import MiniSpice, daf  # noqa: E402

REF = '/root/reference/tests/data'
KERNELS = os.path.join(REF, 'kernels')


def export_fits():
    arrays = {}
    headers = {}
    out_dir = os.path.join(REF, 'outputs')
    files = sorted(os.listdir(out_dir))
    for fn in files:
        if not fn.endswith('.fits'):
            continue
        hdus = read_fits(os.path.join(out_dir, fn))
        headers[fn] = {k: v for k, v in hdus[0][0].items()
                       if isinstance(v, (int, float, str, bool))}
        for i, (hdr, arr) in enumerate(hdus):
            name = hdr.get('EXTNAME', 'PRIMARY') if i else 'PRIMARY'
            if arr is None or name == 'WIREFRAME':
                continue
            arrays[f'{fn}/{name}'] = np.asarray(arr, dtype=np.float64)
    hdus = read_fits(os.path.join(REF, 'inputs', 'test.fits'))
    arrays['inputs/test.fits/PRIMARY'] = np.asarray(hdus[0][1], dtype=np.float64)
    headers['inputs/test.fits'] = {k: v for k, v in hdus[0][0].items()
                                   if isinstance(v, (int, float, str, bool))}
    # raw header cards (primary + first two extensions) of the saved files the FITS staging
    # tests pin their card formatting and HDU structure against
    cards = {fn: read_cards(os.path.join(out_dir, fn), 3)
             for fn in ('test_nav.fits', 'map_rectangular-linear.fits', 'map_rectangular-smooth.fits',
                        'map_rectangular-cubic.fits', 'map_orthographic-1.fits', 'map_azimuthal-1.fits')}
    extnames = {fn: [h[0].get('EXTNAME', 'PRIMARY') for h in read_fits(os.path.join(out_dir, fn))]
                for fn in cards}
    with open(os.path.join(HERE, 'ref_cards.json'), 'w') as f:
        about = {h[0]['EXTNAME']: h[0]['ABOUT']
                 for h in read_fits(os.path.join(out_dir, 'test_nav.fits'))[1:]}
        json.dump({'cards': cards, 'extnames': extnames, 'about': about}, f, indent=0)
    np.savez_compressed(os.path.join(HERE, 'ref_outputs.npz'), **arrays)
    with open(os.path.join(HERE, 'ref_headers.json'), 'w') as f:
        json.dump(headers, f, indent=1, sort_keys=True)
    return arrays, headers


def hst_fixture(ms, arrays, headers):
    hdr = headers['test_nav.fits']
    et = hdr['PLANMAP ET-OBS']
    lt0 = hdr['PLANMAP LIGHT-TIME']
    c = ms.clight()
    ra = math.radians(hdr['PLANMAP TARGET RA'])
    dec = math.radians(hdr['PLANMAP TARGET DEC'])
    P0 = F._radrec(lt0 * c, ra, dec)
    T = ms.ssb_state(599, et - lt0)
    obs_pos = T[:3] - P0
    earth = ms.ssb_state(399, et)

    def rv_plane(vo):
        state = np.concatenate([obs_pos, vo])
        bc = F.build_body_constants(ms, 'JUPITER', '2005-01-01T00:00:00', observer='HST',
                                    et=et, observer_state=state)
        fr = F.pack_frame(bc, nx=7, ny=10, x0=2.5, y0=3.1, r0=3.9,
                          rotation_radians=np.deg2rad(123.456))
        return O.backplanes_img(fr, 7, 10, 1 << 18)[0], bc

    gold = arrays['test_nav.fits/RADIAL-VELOCITY']
    ok = np.isfinite(gold)
    base, _ = rv_plane(earth[3:])
    # radial velocity is affine in the observer velocity: fit the correction
    cols = []
    for k in range(3):
        dv = np.zeros(3)
        dv[k] = 1.0
        plane, _ = rv_plane(earth[3:] + dv)
        cols.append((plane - base)[ok])
    A = np.stack(cols, axis=1)
    sol, *_ = np.linalg.lstsq(A, (gold - base)[ok], rcond=None)
    vo = earth[3:] + sol
    plane, bc = rv_plane(vo)
    resid = float(np.nanmax(np.abs(plane - gold)))
    bc.extra = {
        'note': 'observer position back-derived from golden header; observer velocity '
                'solved from golden RADIAL-VELOCITY plane',
        'hst_speed_wrt_earth_km_s': float(np.linalg.norm(sol)),
        'hst_distance_from_earth_km': float(np.linalg.norm(obs_pos - earth[:3])),
        'rv_fit_residual_km_s': resid,
    }
    with open(os.path.join(HERE, 'jupiter_hst_2005.json'), 'w') as f:
        json.dump(bc.to_json_dict(), f, indent=1)
    print('HST fixture:', bc.extra)
    return bc


def ephem_extract(ms):
    # windows: (body ids, et_lo, et_hi)
    et_fix = 157809664.1839331
    day = 86400.0
    windows = [
        # the 2005-01-01 fixture +- a few days (covers bench time series of 4096 frames
        # 60 s apart = 2.85 days, ending at the fixture epoch; Saturn's SPK ends there)
        ((10, 399, 3, 5, 599, 6, 699, 7, 799, 301, 4, 499, 8, 899, 2, 299, 1, 199), et_fix - 6 * day, et_fix + 6 * day),
        # 2000-01-01 (docs example BodyXY('Jupiter','2000-01-01'), Saturn tests)
        ((10, 399, 3, 5, 599, 6, 699), -2 * day, 2 * day),
    ]
    keep = []
    for seg in ms.segments:
        pieces = []
        for bodies, lo, hi in windows:
            if seg.target in bodies:
                w = daf.window_segment(seg, lo, hi)
                if w is not None:
                    pieces.append(w)
        # merge contiguous/overlapping pieces is unnecessary: keep them separately
        keep.extend(pieces)
    out = os.path.join(ROOT, 'planetmapper_b200', 'data', 'ephem_extract.npz')
    daf.save_extract(out, keep)
    wanted = {}
    for key, val in ms.pool.items():
        if not key.startswith('BODY'):
            continue
        body = key[4:].split('_')[0]
        if body in ('10', '399', '301', '3', '5', '599', '6', '699', '502', '501', '7',
                    '799', '4', '499', '8', '899', '2', '299', '1', '199'):
            if all(isinstance(v, float) for v in val):
                wanted[key] = val
    with open(os.path.join(ROOT, 'planetmapper_b200', 'data', 'pck_pool.json'), 'w') as f:
        json.dump(wanted, f, indent=0, sort_keys=True)
    print('extract:', len(keep), 'segments,', os.path.getsize(out), 'bytes;',
          len(wanted), 'pool variables')


def main():
    ms = MiniSpice.from_kernel_dir(KERNELS)
    arrays, headers = export_fits()
    print('exported', len(arrays), 'golden arrays')
    hst_fixture(ms, arrays, headers)
    ephem_extract(ms)


if __name__ == '__main__':
    main()
