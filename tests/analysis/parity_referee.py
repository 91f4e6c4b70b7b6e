#!/usr/bin/env python
"""
Extended-precision referee at BASELINE config C2's own geometry (CPU only; test infrastructure).

north_star's bars are 1e-9 deg (angles) and 1e-12 relative (distances / velocities) against the
reference's FP64 spiceypy path.  This script measures, pixel by pixel on the C2 frame (Jupiter / HST,
r0 = 0.9 (n-1)/2; default 1024 x 1024, same disc geometry as the 2048 x 2048 bench frame), how far

  (a) the kernels' per-pixel code (host instantiation of planetmapper_b200/csrc/pm_device.cuh,
      tests/host_check/host_check.cu - the same source the GPU runs, with emulated 20-bit MUFU seeds) and
  (b) the CSPICE-shaped FP64 oracle (oracle/pm_oracle.c)

each sit from (c) the SAME oracle evaluated in 80-bit extended precision (64-bit mantissa), and how far
(a) sits from (b).  It shows where two correct FP64 evaluations of this path cannot agree to the bare
bars: the oracle itself misses them against extended precision on the same pixels.

    python tests/analysis/parity_referee.py [size] > profiles/parity_referee_c2.json
"""
import ctypes
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
ROOT = os.path.dirname(TESTS)
sys.path[:0] = [ROOT, TESTS]

import test_host_check as T  # noqa: E402
from helpers import (BARE_ANGLE_PLANES, BARE_REL_PLANES, EMISSION_BINS, PID, WRAP, angle_diff, img_case)  # noqa: E402
from oracle import oracle as O  # noqa: E402
from planetmapper_b200 import frame as F  # noqa: E402

NAMES = ['LON-GRAPHIC', 'LAT-GRAPHIC', 'LON-CENTRIC', 'LAT-CENTRIC', 'INCIDENCE', 'EMISSION', 'PHASE', 'AZIMUTH',
         'DISTANCE', 'RADIAL-VELOCITY', 'DOPPLER']


def diff(name, a, b):
    d = angle_diff(a, b) if name in WRAP else np.abs(a - b)
    if name in BARE_REL_PLANES:
        with np.errstate(invalid='ignore', divide='ignore'):
            d = d / np.abs(b)
    return d


def main():
    sz = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    with open(os.path.join(TESTS, 'golden', 'jupiter_hst_2005.json')) as f:
        bc = F.BodyConstants.from_json_dict(json.load(f))
    fr = img_case(bc, sz, sz, (sz - 1) / 2, (sz - 1) / 2, 0.9 * (sz - 1) / 2, 0.0)
    O.build()
    import conftest
    if not os.path.exists(conftest.HOST_CHECK_SO):
        raise SystemExit('run `pytest tests/test_host_check.py` once to build the host instantiation')
    hc = ctypes.CDLL(conftest.HOST_CHECK_SO)
    with tempfile.TemporaryDirectory() as tmp:
        exact = T.build_ld_oracle(tmp)(fr, sz, sz)
    oracle = O.backplanes_img(fr, sz, sz)
    device = T.hc_img(hc, fr, sz, sz)
    emi = exact[PID['EMISSION']]
    on = np.isfinite(emi) & np.isfinite(oracle[PID['EMISSION']]) & np.isfinite(device[PID['EMISSION']])
    out = {'config': f'C2 geometry at {sz} x {sz}: Jupiter / HST 2005-01-01T00:00:00, r0 = 0.9 (n-1)/2',
           'on_disc_px': int(on.sum()), 'emission_bins_deg': EMISSION_BINS,
           'columns': {'device_vs_exact': 'kernel per-pixel code (host instantiation) vs 80-bit oracle',
                       'oracle_vs_exact': 'FP64 oracle vs 80-bit oracle',
                       'device_vs_oracle': 'kernel per-pixel code vs FP64 oracle (what the GPU parity tests compare)'}}
    for name in NAMES:
        bar = 1e-12 if name in BARE_REL_PLANES else 1e-9
        k = PID[name]
        entry = {'bare_bar': bar, 'unit': 'relative' if name in BARE_REL_PLANES else 'deg'}
        for label, a, b in (('device_vs_exact', device[k], exact[k]), ('oracle_vs_exact', oracle[k], exact[k]),
                            ('device_vs_oracle', device[k], oracle[k])):
            d = np.where(on, diff(name, a, b), 0.0)
            over = on & (d > bar)
            hist, _ = np.histogram(emi[over], bins=EMISSION_BINS)
            entry[label] = {'max': float(d.max()), 'rms': float(np.sqrt(np.mean(d[on] ** 2))),
                            'n_over_bare_bar': int(over.sum()), 'over_by_emission_bin': [int(n) for n in hist]}
        out[name] = entry
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == '__main__':
    main()
