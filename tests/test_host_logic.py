"""Host-side logic that needs no GPU: grid builders, sharding (incl. a 2-process gloo
run), option validation."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from planetmapper_b200.shard import shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
inf = np.inf


@pytest.fixture()
def body(bc_hst):
    import planetmapper_b200 as pm

    return pm.BodyXY(constants=bc_hst, nx=15, ny=10)


def test_rectangular_and_manual_grids_known_answers(body):
    """tests/test_body_xy.py:1612-1686."""
    cases = [
        (None, None, [[315.0, 225.0, 135.0, 45.0]] * 2, [[-45.0] * 4, [45.0] * 4]),
        ((-inf, inf), (-inf, inf), [[315.0, 225.0, 135.0, 45.0]] * 2, [[-45.0] * 4, [45.0] * 4]),
        ((135, -inf), (45, inf), [[135.0, 45.0]], [[45.0, 45.0]]),
        ((100, 300), (-50, 50), [[225.0, 135.0]] * 2, [[-45.0, -45.0], [45.0, 45.0]]),
        ((300, 100), (50, -50), [[225.0, 135.0]] * 2, [[-45.0, -45.0], [45.0, 45.0]]),
    ]
    for xlim, ylim, elon, elat in cases:
        lons, lats, xx, yy, _t, info = body.generate_map_coordinates(degree_interval=90, xlim=xlim, ylim=ylim)
        assert np.array_equal(lons, np.array(elon)) and np.array_equal(lats, np.array(elat))
        assert np.array_equal(xx, np.array(elon)) and np.array_equal(yy, np.array(elat))
        assert info['xlim'] == xlim and info['ylim'] == ylim
        assert not lons.flags.writeable
    lons, lats, *_ = body.generate_map_coordinates(degree_interval=123)
    assert np.array_equal(lons, [[307.5, 184.5, 61.5]]) and np.array_equal(lats, [[-28.5] * 3])
    lons, lats, *_rest = body.generate_map_coordinates('manual', lon_coords=[1, 2, 3], lat_coords=[4, 5])
    assert lons.shape == (2, 3) and lats[1, 0] == 5
    info = body.generate_map_coordinates(degree_interval=30, alt=12.5)[5]
    assert info['alt'] == 12.5
    for kw in (dict(), dict(lon_coords=np.array([1, 2, 3]), lat_coords=np.array([[1, 2, 3], [4, 5, 6]])),
               dict(lon_coords=np.array([[[1, 2, 3]]]), lat_coords=np.array([[[1, 2, 3]]])),
               dict(lon_coords=np.array([[1, 2, 3]]), lat_coords=np.array([[1, 2, 3], [4, 5, 6]]))):
        with pytest.raises(ValueError):
            body.generate_map_coordinates('manual', **kw)


def test_create_proj_string_literals(body):
    """tests/test_body_xy.py:1990-2029 (the Jupiter literals; Earth's follow the same code)."""
    f = body.create_proj_string
    assert f('ortho') == '+proj=ortho +a=71492.0 +b=66854.0 +axis=wnu +type=crs'
    assert f('ortho', axis=None) == '+proj=ortho +a=71492.0 +b=66854.0 +type=crs'
    assert f('ortho', a=None, axis=None) == '+proj=ortho +b=66854.0 +type=crs'
    assert f('ortho', axis='123') == '+proj=ortho +axis=123 +a=71492.0 +b=66854.0 +type=crs'
    assert (f('eqc', string='a_string', number=123, lat_0=-1.234)
            == '+proj=eqc +string=a_string +number=123 +lat_0=-1.234 +a=71492.0 +b=66854.0 +axis=wnu +type=crs')
    assert f('ortho', lon_0=180, lat_0=45, axis=None, a=None, b=None) == '+proj=ortho +lon_0=180 +lat_0=45 +type=crs'


def test_custom_proj_string_validation_needs_no_device(body):
    """Argument and axis checks of generate_map_coordinates for proj strings
    (tests/test_body_xy.py:1563-1590, :1948-1988) happen before any device work."""
    import planetmapper_b200 as pm

    g = body.generate_map_coordinates
    with pytest.raises(ValueError):
        g('+proj=ortho +R=1 +axis=wnu +type=crs')                       # x coords must be provided
    with pytest.raises(ValueError):
        g('proj=ortho +R=1 +axis=wnu +type=crs', projection_x_coords=np.array([1, 2, 3]),
          projection_y_coords=np.array([[1, 2, 3], [4, 5, 6]]))
    with pytest.raises(ValueError):
        g('proj=ortho +R=1 +axis=wnu +type=crs', projection_x_coords=np.array([[[1, 2, 3]]]))
    with pytest.raises(ValueError):
        g('proj=ortho +R=1 +axis=wnu +type=crs', projection_x_coords=np.array([[1, 2, 3]]),
          projection_y_coords=np.array([[1, 2, 3], [4, 5, 6]]))
    for bad_axis in ('+proj=ortho +R=1 +type=crs', '+proj=ortho +R=1 +axis=enu +type=crs',
                     '+proj=ortho +R=1 +axis=neu +type=crs'):
        with pytest.raises(pm.ProjStringError):
            g(bad_axis, projection_x_coords=np.array([0, 0.25, 0.5]))
    # outside the accelerated subset: a clear error, never a silent approximation
    for unsupported in ('+proj=eqc +axis=wnu +type=crs', '+proj=ortho +k_0=2 +axis=wnu', '+proj=aeqd +axis=wnu',
                        '+proj=laea +a=2 +b=1 +axis=wnu', '+proj=ortho +R=abc +axis=wnu',
                        '+proj=ortho +units=km +axis=wnu'):
        with pytest.raises(pm.ProjStringError):
            g(unsupported, projection_x_coords=np.array([0, 0.25, 0.5]))


def test_disc_parameter_interface(body):
    """tests/test_body_xy.py set/get behaviour incl. centre_disc (body_xy.py:791-803)."""
    assert body.get_disc_params() == (7.0, 4.5, 0.9 * 4.5, 0.0)
    assert body.get_disc_method() == 'centre_disc'
    body.set_disc_params(5, 8, 3, 45)
    assert body.get_disc_params() == pytest.approx((5, 8, 3, 45))
    assert body.get_disc_method() == 'default'          # cleared with the cache
    assert body.get_plate_scale_arcsec() == pytest.approx(35.98242689969618 / 6)
    assert body.get_plate_scale_km() == pytest.approx(35.98242689969618 / 6 * 3973.7175149019004)
    for bad in (np.nan, np.inf):
        with pytest.raises(ValueError):
            body.set_x0(bad)
    with pytest.raises(ValueError):
        body.set_r0(0)
    body.set_img_size(20, 30)
    assert body.get_img_size() == (20, 30)
    body.rotate_north_to_top()
    assert body.get_rotation() == pytest.approx(24.15516987997688, abs=1e-8)


def test_series_frames_over_host_processes_are_bit_identical():
    """planetmapper_b200.series: epochs sharded over worker processes give exactly the serial
    constants, in order; worker errors surface in the parent."""
    import planetmapper_b200 as pm
    from planetmapper_b200 import frame as F
    from planetmapper_b200 import series as S

    prov = pm.get_default_provider()
    ets = prov.utc2et('2005-01-01T00:00:00') - 3600.0 - 60.0 * np.arange(70)[::-1]
    disc = dict(nx=64, ny=48, x0=31.5, y0=23.5, r0=20.0, rotation_radians=0.3)
    serial = S.build_series_frames('Jupiter', ets, 'EARTH', workers=1, **disc)
    assert serial.shape == (70, F.PMFRAME_NDOUBLES)
    bc = F.build_body_constants(prov, 'Jupiter', None, 'EARTH', et=float(ets[5]))
    assert np.array_equal(serial[5], F.pack_frame(bc, **disc))
    try:
        for workers in (2, 3):
            assert np.array_equal(S.build_series_frames('Jupiter', ets, 'EARTH', workers=workers, **disc), serial)
        assert np.array_equal(S.build_series_frames('Jupiter', ets[:5], 'EARTH', workers=4, **disc), serial[:5])
        with pytest.raises(RuntimeError, match='unknown body'):
            S.build_series_frames('Nosuchbody', ets, 'EARTH', workers=2, **disc)
        assert np.array_equal(S.build_series_frames('Jupiter', ets, 'EARTH', workers=2, **disc), serial)  # pool restarts
    finally:
        S.shutdown_pool()
    assert S.default_workers() >= 1


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 3000, 4096, 4097):
        for w in (1, 2, 3, 4, 8):
            blocks = [shard_range(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_two_process_gloo_sharding(tmp_path):
    """world_size 2 over gloo: each rank takes its block of frames, builds the frame
    constants for its block on the host, and the gathered checksums equal the
    single-process result (the N > 1 control path of bench.py, minus the GPU)."""
    script = tmp_path / 'worker.py'
    script.write_text(textwrap.dedent(f'''
        import sys, json
        sys.path.insert(0, {ROOT!r})
        import numpy as np, torch, torch.distributed as dist
        import planetmapper_b200 as pm
        from planetmapper_b200 import frame as F
        from planetmapper_b200.shard import env_rank_world, shard_range, gather_shard_results, max_over_ranks, gather_blocks
        rank, local_rank, world = env_rank_world()
        dist.init_process_group('gloo')
        n_frames = 9
        ets = 157500000.0 + 60.0 * np.arange(n_frames)
        lo, hi = shard_range(n_frames, rank, world)
        prov = pm.get_default_provider()
        sums = []
        for et in ets[lo:hi]:
            bc = F.build_body_constants(prov, 'JUPITER', None, 'EARTH', et=float(et))
            fr = F.pack_frame(bc, nx=64, ny=64, x0=31.5, y0=31.5, r0=28.0, rotation_radians=0.0)
            sums.append(float(np.sum(fr)))
        dist.barrier()
        t = max_over_ranks(float(rank + 1), world)
        allsums = gather_shard_results((lo, hi, sums), world)
        # result assembly: each rank's block of planes lands in one tensor on rank 0
        local = torch.arange(lo, hi, dtype=torch.float64)[:, None] * torch.ones(1, 3, dtype=torch.float64)
        whole = gather_blocks(local, n_frames, world, rank)
        if rank == 0:
            assert whole.shape == (n_frames, 3) and torch.equal(whole[:, 0], torch.arange(n_frames, dtype=torch.float64))
        else:
            assert whole is None
        if rank == 0:
            flat = [s for _, _, ss in sorted(allsums) for s in ss]
            print('RESULT ' + json.dumps(dict(t=t, n=len(flat), checksum=sum(flat), blocks=[(a, b) for a, b, _ in sorted(allsums)])))
        dist.destroy_process_group()
    '''))
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', OMP_NUM_THREADS='1')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', '29731', str(script)]
    proc = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert proc.returncode == 0, proc.stderr[-2000:]
    import json

    line = [l for l in proc.stdout.splitlines() if l.startswith('RESULT ')][0]
    res = json.loads(line[7:])
    assert res['t'] == 2.0 and res['n'] == 9 and res['blocks'] == [[0, 5], [5, 9]]
    # single-process reference
    import planetmapper_b200 as pm
    from planetmapper_b200 import frame as F

    prov = pm.get_default_provider()
    total = 0.0
    for et in 157500000.0 + 60.0 * np.arange(9):
        bc = F.build_body_constants(prov, 'JUPITER', None, 'EARTH', et=float(et))
        total += float(np.sum(F.pack_frame(bc, nx=64, ny=64, x0=31.5, y0=31.5, r0=28.0, rotation_radians=0.0)))
    assert res['checksum'] == pytest.approx(total, rel=1e-15)


def test_progress_hooks_follow_the_reference_protocol():
    """progress.py:16-41 / base.py:760-781: 0 and 1 around every decorated call, the stack of running
    qualified names, fractions in between, exceptions from the hook propagate (cancel) and unwind the stack."""
    from planetmapper_b200.progress import ProgressMixin, progress_decorator

    class Thing(ProgressMixin):
        def __init__(self):
            self._progress_hook = None
            self._progress_call_stack = []

        @progress_decorator
        def outer(self, n):
            for k in range(n):
                self.inner()
                self._update_progress_hook((k + 1) / n)
            return n

        @progress_decorator
        def inner(self):
            return 1

    calls = []
    t = Thing()
    assert t.outer(2) == 2 and calls == []          # no hook: plain call
    t._set_progress_hook(lambda p, stack: calls.append((p, tuple(stack))))
    assert t.outer(2) == 2
    o, i = 'Thing.outer', 'Thing.inner'
    names = lambda c: tuple(n.split('<locals>.')[-1] for n in c)   # noqa: E731
    assert [(p, names(s)) for p, s in calls] == [
        (0, (o,)), (0, (o, i)), (1, (o, i)), (0.5, (o,)), (0, (o, i)), (1, (o, i)), (1.0, (o,)), (1, (o,))]
    assert t._progress_call_stack == []

    def cancel(p, stack):
        if len(stack) == 2:
            raise KeyboardInterrupt('cancelled from the hook')
    t._set_progress_hook(cancel)
    with pytest.raises(KeyboardInterrupt):
        t.outer(3)
    assert t._progress_call_stack == []
    t._remove_progress_hook()
    assert t._get_progress_hook() is None and t.outer(1) == 1


def test_interpolation_modes_and_the_smoothing_refusal():
    """map_img's `interpolation` argument (body_xy.py:1592-1631) as the kernels see it.  FITPACK smoothing
    splines (spline_smoothing > 0, body_xy.py:1673-1680) are NOT accelerated: the reference's own tests call
    their values scipy-version dependent (tests/test_observation.py:1163-1169), so there is nothing to hold
    parity against and the call is refused loudly instead of approximated."""
    from planetmapper_b200 import _lib as L
    from planetmapper_b200.body_xy import INTERP_SMOOTH, _interpolation_mode

    assert _interpolation_mode('nearest', 0) == L.INTERP_NEAREST
    assert _interpolation_mode('linear', 0) == _interpolation_mode(1, 0) == _interpolation_mode((1, 1), 0) == L.INTERP_LINEAR
    assert _interpolation_mode('quadratic', 0) == _interpolation_mode(2, 0) == L.INTERP_QUADRATIC
    assert _interpolation_mode('cubic', 0) == _interpolation_mode(3, 0) == _interpolation_mode((3, 3), 0) == L.INTERP_CUBIC
    assert _interpolation_mode((1, 3), 0) == L.INTERP_MIXED | (1 << 4) | 3
    assert _interpolation_mode('smooth', 0) == INTERP_SMOOTH
    assert _interpolation_mode('nearest', 0.5) == L.INTERP_NEAREST and _interpolation_mode('smooth', 2) == INTERP_SMOOTH
    for interp in ('linear', 'quadratic', 'cubic', 1, 3, (2, 3)):
        with pytest.raises(NotImplementedError, match='smoothing'):
            _interpolation_mode(interp, 0.5)
    for interp in (4, 5, (1, 5), (0, 2)):
        with pytest.raises(NotImplementedError):
            _interpolation_mode(interp, 0)
    for interp in ('<<<test>>>', (1, 2, 3), None, True):
        with pytest.raises(ValueError):
            _interpolation_mode(interp, 0)
