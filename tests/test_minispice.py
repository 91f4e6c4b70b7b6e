"""MiniSpice (host-side, once-per-frame ephemeris stand-in) against the reference's
known answers, and the bundled extract against itself."""
import math
import os

import numpy as np
import pytest

import planetmapper_b200 as pm
from planetmapper_b200 import frame as F
from planetmapper_b200.minispice import MiniSpice, utc2et
from planetmapper_b200.minispice.textkernel import parse_text_kernel

REF_KERNELS = '/root/reference/tests/data/kernels'


def test_utc2et_known_answer():
    # tests/test_body.py:110
    assert utc2et('2005-01-01T00:00:00') == pytest.approx(157809664.1839331, abs=1e-7)
    assert utc2et('2000-01-01T12:00:00') == pytest.approx(64.18392728473108, abs=1e-6)
    assert utc2et('2005-01-01') == utc2et('2005-01-01T00:00:00')


def test_subpoint_lon_known_answer_from_extract():
    # tests/test_body.py:39-42: Body('Jupiter', utc='2005-01-01').subpoint_lon
    ms = pm.get_default_provider()
    bc = F.build_body_constants(ms, 'Jupiter', '2005-01-01', 'EARTH')
    assert bc.subpoint_lon == pytest.approx(153.12547767272153, abs=1e-9)
    assert bc.positive_longitude_direction == 'W'
    assert bc.target_id == 599


@pytest.mark.skipif(not os.path.isdir(REF_KERNELS), reason='reference kernels not present')
def test_extract_reproduces_full_kernels():
    full = MiniSpice.from_kernel_dir(REF_KERNELS)
    ext = pm.get_default_provider()
    if not isinstance(ext, MiniSpice):
        pytest.skip('default provider is spiceypy')
    for body in (10, 399, 599, 699):
        for et in (157809664.1839331 - 3 * 86400, 157809000.0, 0.0):
            assert np.array_equal(full.ssb_state(body, et), ext.ssb_state(body, et)), (body, et)
    r1, w1 = full.orientation(599, 1.5e8)
    r2, w2 = ext.orientation(599, 1.5e8)
    assert np.array_equal(r1, r2) and np.array_equal(w1, w2)


def test_orientation_is_a_rotation_and_omega_matches_finite_difference():
    ms = pm.get_default_provider()
    if not isinstance(ms, MiniSpice):
        pytest.skip('default provider is spiceypy')
    et = 157806930.0
    r, w = ms.orientation(599, et)
    assert np.allclose(r @ r.T, np.eye(3), atol=1e-14)
    h = 1.0  # truncation ~ omega^3 h^2 / 6 = 9e-13; W (1.6e6 deg) quantisation ~ 2e-12
    rp, _ = ms.orientation(599, et + h)
    rm, _ = ms.orientation(599, et - h)
    drdt = (rp - rm) / (2 * h)
    om = -drdt @ r.T
    w_fd = np.array([om[2, 1], om[0, 2], om[1, 0]])
    assert np.allclose(w, w_fd, rtol=0, atol=1e-11)
    # Jupiter System III: 870.536 deg/day
    assert math.degrees(w[2]) * 86400 == pytest.approx(870.536, abs=1e-6)


def test_text_kernel_parser():
    pool = parse_text_kernel("""
    junk BODY1_X = 5
    \\begindata
      BODY599_RADII = ( 71492   71492   66854 )
      BODY599_PM    = ( 284.95  870.5360000  0. )
      DELTET/K      = 1.657D-3
      NAME          = 'it''s'
      LIST         += ( 1 2 )
      LIST         += 3
    \\begintext
      BODY599_RADII = ( 1 1 1 )
    """)
    assert pool['BODY599_RADII'] == [71492.0, 71492.0, 66854.0]
    assert pool['DELTET/K'] == [1.657e-3]
    assert pool['NAME'] == ["it's"]
    assert pool['LIST'] == [1.0, 2.0, 3.0]
    assert 'BODY1_X' not in pool


def test_saturn_frame_builds_before_spk_edge():
    ms = pm.get_default_provider()
    bc = F.build_body_constants(ms, 'Saturn', '2004-12-30T00:00:00', 'EARTH')
    assert bc.radii[0] == 60268.0 and bc.prograde
    assert 1.1e9 < bc.target_distance < 1.7e9
