"""
The reference's own host setup - spiceypy - through ``planetmapper_b200/spice_host.py`` (north_star: "The
Python host keeps the existing spiceypy setup"; planetmapper/base.py:794-839 ``str2et`` / ``spkezr``,
:939-977 kernel load order, body.py:522-567).  spiceypy is not installable here, so a stand-in module with
CSPICE's calling conventions (tests/fake_spiceypy/spiceypy.py, answered by MiniSpice, angular velocity by an
independent finite difference) is put on sys.path: the adapter, the provider selection of
``get_default_provider`` and the sign convention of omega all execute in CI.
"""
import importlib
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
FAKE = os.path.join(HERE, 'fake_spiceypy')


@pytest.fixture()
def fake_spiceypy(tmp_path):
    import planetmapper_b200 as pm

    sys.path.insert(0, FAKE)
    sys.modules.pop('spiceypy', None)
    spice = importlib.import_module('spiceypy')
    spice.kclear()
    # a kernel directory laid out like a user's (nested folders): the repository's ephemeris extract + pool
    deep = tmp_path / 'kernels' / 'spk' / 'planets'
    deep.mkdir(parents=True)
    (tmp_path / 'kernels' / 'pck').mkdir()
    import shutil

    shutil.copy(os.path.join(pm._DATA_DIR, 'ephem_extract.npz'), deep / 'ephem_extract.npz')
    shutil.copy(os.path.join(pm._DATA_DIR, 'pck_pool.json'), tmp_path / 'kernels' / 'pck' / 'pck_pool.json')
    old_path, old_provider, old_custom = pm._KERNEL_PATH, pm._PROVIDER, pm._PROVIDER_IS_CUSTOM
    pm.set_kernel_path(str(tmp_path / 'kernels'))
    yield spice, str(tmp_path / 'kernels')
    pm._KERNEL_PATH, pm._PROVIDER, pm._PROVIDER_IS_CUSTOM = old_path, old_provider, old_custom
    sys.path.remove(FAKE)
    sys.modules.pop('spiceypy', None)


def test_default_provider_prefers_spiceypy_and_loads_kernels_in_reference_order(fake_spiceypy):
    import planetmapper_b200 as pm
    from planetmapper_b200.spice_host import SpiceProvider

    spice, kdir = fake_spiceypy
    provider = pm.get_default_provider()
    assert isinstance(provider, SpiceProvider) and provider.name == 'spiceypy'
    # deepest directory first, then alphabetical (base.py:968-977): spk/planets/... before pck/...
    assert [os.path.relpath(p, kdir) for p in spice._LOADED] == [os.path.join('spk', 'planets', 'ephem_extract.npz'),
                                                                os.path.join('pck', 'pck_pool.json')]
    assert provider.bods2c('jupiter') == 599 and provider.bods2c(10) == 10 and provider.bodc2n(599) == 'JUPITER'
    assert provider.clight() == 299792.458
    radii = provider.bodvar(599, 'RADII')
    assert radii.shape == (3,) and radii[0] == 71492.0 and radii[2] == 66854.0


def test_spice_provider_constants_equal_minispice(fake_spiceypy):
    """The five primitives through the spiceypy adapter give the same frame constants as MiniSpice asked
    directly; omega - read out of sxform's derivative block by SpiceProvider.orientation - has the sign and
    size of the analytic angular velocity."""
    import planetmapper_b200 as pm
    from planetmapper_b200 import frame as F
    from planetmapper_b200.minispice import MiniSpice

    provider = pm.get_default_provider()
    direct = MiniSpice.from_extract(os.path.join(pm._DATA_DIR, 'ephem_extract.npz'),
                                    os.path.join(pm._DATA_DIR, 'pck_pool.json'))
    utc = '2004-12-30T12:00:00'
    assert provider.utc2et(utc) == direct.utc2et(utc)
    et = direct.utc2et(utc)
    for body in (599, 699, 399, 10):
        assert np.array_equal(provider.ssb_state(body, et), direct.ssb_state(body, et))
    for body in (599, 699, 299, 399):     # prograde, prograde, retrograde (Venus), Earth
        r_a, w_a = provider.orientation(body, et)
        r_b, w_b = direct.orientation(body, et)
        assert np.array_equal(r_a, r_b)
        assert np.allclose(w_a, w_b, rtol=1e-7, atol=1e-7 * np.linalg.norm(w_b)), (body, w_a, w_b)
        assert np.sign(w_a[2]) == np.sign(w_b[2]) and abs(w_a[2]) > 1e-7
    for target, observer in (('Jupiter', 'EARTH'), ('Saturn', 'EARTH'), ('Venus', 'EARTH')):
        a = F.build_body_constants(provider, target, utc, observer)
        b = F.build_body_constants(direct, target, utc, observer)
        fa = F.pack_frame(a, nx=64, ny=48, x0=30.0, y0=20.0, r0=15.0, rotation_radians=0.2)
        fb = F.pack_frame(b, nx=64, ny=48, x0=30.0, y0=20.0, r0=15.0, rotation_radians=0.2)
        o, n = F.PMFRAME_OFFSETS['omega']
        rest = np.ones(F.PMFRAME_NDOUBLES, dtype=bool)
        rest[o:o + n] = False
        assert np.allclose(fa[rest], fb[rest], rtol=1e-13, atol=1e-13), target
        assert np.allclose(fa[o:o + n], fb[o:o + n], rtol=1e-7, atol=1e-12), target


def test_spiceypy_without_kernels_falls_back_to_the_bundled_extract(fake_spiceypy, tmp_path):
    """spiceypy importable but nothing to furnish (the reference would fail in str2et): the provider falls back
    to MiniSpice over the bundled extract instead of handing out a SpiceProvider with no kernels loaded."""
    import planetmapper_b200 as pm

    empty = tmp_path / 'no_kernels_here'
    empty.mkdir()
    pm.set_kernel_path(str(empty))
    provider = pm.get_default_provider()
    assert provider.name.startswith('minispice')
    assert pm.get_kernel_path(return_source=True) == (str(empty), 'set_kernel_path()')
    pm.set_kernel_path(None)
    assert pm.get_kernel_path() in (os.environ.get('PLANETMAPPER_KERNEL_PATH'), pm.DEFAULT_KERNEL_PATH)
