"""
Frame constants for a TIME SERIES (BASELINE config C5: thousands of frames, a fresh
``BodyXY`` per epoch in the reference).

Once the per-pixel work takes ~40 microseconds per 1024 x 1024 frame on the GPU, the serial
host extraction of each frame's 92 constants (about 55 ephemeris / orientation evaluations,
~3 ms in Python, the same order through spiceypy: SURVEY.md 8(e) "limited only by serial
host constant extraction") is the whole cost of a series.  Epochs are independent, so the
extraction shards over host processes exactly like frames shard over GPUs: contiguous
blocks of epochs, no exchange, the SAME scalar code per epoch (``frame.build_body_constants``
+ ``frame.pack_frame``), hence bit-identical constants.  SPICE itself is not re-entrant
(SURVEY.md 8(b) "Threading"), which is why the workers are processes, each with its own
provider, and never threads.
"""
from __future__ import annotations

import atexit
import os
import pickle
import struct
import subprocess
import sys

import numpy as np

from . import frame as F
from .shard import shard_range


def _block(provider, target, observer, ets, disc) -> np.ndarray:
    out = np.empty((len(ets), F.PMFRAME_NDOUBLES))
    for i, et in enumerate(ets):
        bc = F.build_body_constants(provider, target, None, observer, et=float(et), with_subsol=False)
        out[i] = F.pack_frame(bc, **disc)
    return out


# ---- worker processes ---------------------------------------------------------------------
# Plain child interpreters running `python -m planetmapper_b200.series` and talking
# length-prefixed pickles over their pipes.  (multiprocessing's spawn / forkserver start
# methods re-import the caller's __main__, which breaks unguarded user scripts, and fork is not
# safe once the parent holds a CUDA context.)
def _send(stream, obj) -> None:
    data = pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL)
    stream.write(struct.pack('<Q', len(data)))
    stream.write(data)
    stream.flush()


def _recv(stream):
    head = stream.read(8)
    if len(head) < 8:
        raise EOFError('series worker closed its pipe')
    (n,) = struct.unpack('<Q', head)
    return pickle.loads(stream.read(n))


def _worker_main() -> None:
    import planetmapper_b200 as pm

    inp, out = sys.stdin.buffer, sys.stdout.buffer
    sys.stdout = sys.stderr  # stray prints must not corrupt the reply stream
    provider, provider_path = None, None
    while True:
        try:
            msg = _recv(inp)
        except EOFError:
            return
        try:
            kernel_path, target, observer, ets, disc, kepler = msg
            if provider is None or kernel_path != provider_path:   # the cached provider is keyed on its path
                pm.set_kernel_path(kernel_path)
                provider, provider_path = pm.get_default_provider(), kernel_path
            _send(out, ('ok', _block(_with_kepler(provider) if kepler else provider, target, observer, ets, disc)))
        except Exception as exc:  # reported to the parent, which raises
            _send(out, ('error', f'{type(exc).__name__}: {exc}'))


def _with_kepler(provider):
    """Analytic orbits for the bodies the kernels do not cover (minispice/kepler.py: BASELINE config C5 names
    Europa, which has no SPK segment in the bundled kernels)."""
    from .minispice.kepler import KeplerOrbitProvider

    return provider if isinstance(provider, KeplerOrbitProvider) else KeplerOrbitProvider(provider)


_WORKERS: list[subprocess.Popen] = []


def _workers(n: int) -> list[subprocess.Popen]:
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=root + os.pathsep + os.environ.get('PYTHONPATH', ''),
               OMP_NUM_THREADS='1', OPENBLAS_NUM_THREADS='1', MKL_NUM_THREADS='1')
    _WORKERS[:] = [w for w in _WORKERS if w.poll() is None]
    while len(_WORKERS) < n:
        _WORKERS.append(subprocess.Popen([sys.executable, '-m', 'planetmapper_b200.series'], stdin=subprocess.PIPE,
                                         stdout=subprocess.PIPE, env=env))
    return _WORKERS[:n]


def shutdown_pool() -> None:
    """Stop the worker processes (they are otherwise kept for the next series)."""
    for w in _WORKERS:
        try:
            w.stdin.close()
            w.wait(timeout=10)
        except Exception:
            w.kill()
    _WORKERS.clear()


atexit.register(shutdown_pool)


def default_workers() -> int:
    """Host processes to use: the cores this process may run on, shared between the ranks
    of a torchrun job (LOCAL_WORLD_SIZE)."""
    total = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        cores = total
    if cores < total:   # this rank already owns a slice of the host (shard.bind_rank_to_cores, taskset, cgroups)
        return max(1, cores)
    ranks = max(1, int(os.environ.get('LOCAL_WORLD_SIZE', '1')))
    return max(1, cores // ranks)


def build_series_frames(target, ets, observer='EARTH', *, nx: int, ny: int, x0: float, y0: float,
                        r0: float, rotation_radians: float = 0.0, alt: float = 0.0,
                        workers: int | None = None, provider=None, kepler: bool = False) -> np.ndarray:
    """PMFrame constants, shape (len(ets), 92), for ``target`` seen from ``observer`` at the
    ephemeris times ``ets`` with one set of disc parameters.

    ``workers`` host processes (default: :func:`default_workers`) each take one contiguous
    block of epochs; ``workers <= 1``, a short series, or an explicit in-process ``provider``
    runs serially in this process.  The result does not depend on ``workers``.  ``kepler=True`` answers
    the ephemeris of bodies without SPK coverage from the analytic orbits of ``minispice/kepler.py``
    (synthetic positions: Europa of BASELINE config C5).
    """
    import planetmapper_b200 as pm

    ets = np.asarray(ets, dtype=np.float64).reshape(-1)
    disc = dict(nx=nx, ny=ny, x0=x0, y0=y0, r0=r0, rotation_radians=rotation_radians, alt=alt)
    workers = default_workers() if workers is None else int(workers)
    workers = min(workers, max(1, len(ets) // 16))   # a block under ~16 epochs does not pay for the hand-off
    if pm.provider_is_custom():   # an in-process provider cannot be rebuilt inside a worker
        workers = 1
    if provider is not None or workers <= 1:
        prov = provider if provider is not None else pm.get_default_provider()
        return _block(_with_kepler(prov) if kepler else prov, target, observer, ets, disc)
    procs = _workers(workers)
    try:
        for w, proc in enumerate(procs):
            block = ets[slice(*shard_range(len(ets), w, workers))]
            _send(proc.stdin, (pm._KERNEL_PATH, str(target), str(observer), block, disc, bool(kepler)))
        replies = [_recv(proc.stdout) for proc in procs]   # every reply is drained before any error is raised
    except BaseException:
        shutdown_pool()   # a dead or half-read worker must not serve the next call
        raise
    failed = [payload for status, payload in replies if status != 'ok']
    if failed:
        raise RuntimeError(f'series worker failed: {failed[0]}')
    return np.concatenate([payload for _, payload in replies], axis=0)


def iter_backplane_batches(frames: np.ndarray, nx: int, ny: int, names, batch: int = 32):
    """Backplane images of a series, ``batch`` frames per fused launch.

    Yields ``(first, planes)`` with ``planes`` a CUDA tensor of shape
    ``(n, len(set(names)), ny, nx)`` holding frames ``first .. first + n`` (planes in
    backplane-registration order, like ``BodyXY.get_backplane_imgs``).  The tensor is REUSED by
    the next iteration (a 4096-frame series of 12 planes would be 412 GB): consume or copy it
    before advancing.  Equivalent to ``BodyXY(target, utc_i, ...).get_backplane_img(name)`` for
    every epoch and name in the reference (body_xy.py:2586), one launch per batch instead of one
    Python loop per pixel, stage and frame.
    """
    from . import _lib as L

    torch = L._torch()
    mask = L.mask_from_names([str(n).strip().upper() for n in names])
    fd = L.to_device(np.ascontiguousarray(frames, dtype=np.float64))
    out = torch.empty((min(batch, len(frames)), L.popcount(mask), ny, nx), dtype=torch.float64, device=fd.device)
    for first in range(0, len(frames), batch):
        n = min(batch, len(frames) - first)
        L.backplanes_img(fd[first:first + n], nx, ny, mask, out=out[:n])
        yield first, out[:n]


def map_series(frames: np.ndarray, imgs, nx: int, ny: int, lons, lats, *, interpolation='linear',
               propagate_nan: bool = True, batch: int = 32, out=None):
    """``BodyXY(...).map_img(img_f, ...)`` for every frame f of a series (body_xy.py:1414-1631 once per
    epoch in the reference): frame f's image is mapped onto the lon / lat grid with frame f's own
    disc geometry.  Per batch of frames: ONE launch for all x / y maps (``pm_backplanes_map_batch``),
    the NaN repair of the batch's images (linear), ONE paired gather.

    ``imgs``: (n_frames, ny, nx) array or CUDA tensor; ``lons`` / ``lats``: the map grid (degrees,
    e.g. from ``generate_map_coordinates``); ``interpolation``: 'nearest' or 'linear'.  Returns a CUDA
    tensor (n_frames,) + grid shape (``out`` if given).
    """
    from . import _lib as L

    torch = L._torch()
    mode = {'nearest': L.INTERP_NEAREST, 'linear': L.INTERP_LINEAR, 1: L.INTERP_LINEAR}.get(interpolation)
    if mode is None:
        raise NotImplementedError("map_series: interpolation must be 'nearest' or 'linear'")
    fd = L.to_device(np.ascontiguousarray(frames, dtype=np.float64))
    if not isinstance(imgs, torch.Tensor):
        imgs = L.to_device(imgs)
    if tuple(imgs.shape) != (len(frames), ny, nx):
        raise ValueError(f'imgs must have shape ({len(frames)}, {ny}, {nx})')
    lod = L.to_device(np.asarray(lons, dtype=np.float64) % 360)
    lad = L.to_device(lats)
    if out is None:
        out = torch.empty((len(frames),) + tuple(lod.shape), dtype=torch.float64, device=fd.device)
    xy_mask = L.mask_from_names(['PIXEL-X', 'PIXEL-Y'])
    xy = torch.empty((min(batch, len(frames)), 2) + tuple(lod.shape), dtype=torch.float64, device=fd.device)
    for first in range(0, len(frames), batch):
        n = min(batch, len(frames) - first)
        L.backplanes_map_batch(fd[first:first + n], lod, lad, xy_mask, out=xy[:n])
        cube = imgs[first:first + n].contiguous()
        src = cube if mode == L.INTERP_NEAREST else L.spline_prepare(cube, 1)
        L.gather_paired(src, xy[:n, 0], xy[:n, 1], mode, propagate_nan=propagate_nan, out=out[first:first + n])
    return out


if __name__ == '__main__':
    _worker_main()
