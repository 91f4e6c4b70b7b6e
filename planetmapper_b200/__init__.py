"""
planetmapper_b200 - B200-native (sm_100a) implementation of PlanetMapper's per-pixel
geometry and mapping hot path behind the BodyXY / Observation API.

    from planetmapper_b200 import BodyXY, Observation

The per-pixel work runs in hand-written FP64 CUDA kernels (libpm_b200.so, C ABI in
include/pm_b200.h).  There is no CPU fallback.
"""
from __future__ import annotations

import os

from .body_xy import (Backplane, BackplaneNotFoundError, BodyXY, MapTransformer, NotFoundError,  # noqa: F401
                      ProjStringError)
from .frame import BodyConstants, build_body_constants, pack_frame  # noqa: F401
from .observation import Observation  # noqa: F401

__version__ = '0.1.0'

_DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data')
_KERNEL_PATH: str | None = None
_PROVIDER = None
_PROVIDER_IS_CUSTOM = False


DEFAULT_KERNEL_PATH = '~/spice_kernels/'   # base.py:33


def _drop_series_pool() -> None:
    # series workers cache a provider built from the kernel path: a new path / provider must not
    # be served by workers holding the old one
    from . import series

    series.shutdown_pool()


def set_kernel_path(path: str | None) -> None:
    """Directory holding SPICE kernels (planetmapper.set_kernel_path, base.py:1018)."""
    global _KERNEL_PATH, _PROVIDER, _PROVIDER_IS_CUSTOM
    _KERNEL_PATH = path
    _PROVIDER = None
    _PROVIDER_IS_CUSTOM = False
    _drop_series_pool()


def get_kernel_path(return_source: bool = False):
    """planetmapper.get_kernel_path (base.py:1054-1086): set_kernel_path() value, else the
    PLANETMAPPER_KERNEL_PATH environment variable, else DEFAULT_KERNEL_PATH."""
    if _KERNEL_PATH is not None:
        path, source = _KERNEL_PATH, 'set_kernel_path()'
    elif os.environ.get('PLANETMAPPER_KERNEL_PATH'):
        path, source = os.environ['PLANETMAPPER_KERNEL_PATH'], 'PLANETMAPPER_KERNEL_PATH'
    else:
        path, source = DEFAULT_KERNEL_PATH, 'default'
    return (path, source) if return_source else path


def set_default_provider(provider) -> None:
    """Install an in-process ephemeris provider.  It cannot be reproduced in the series worker
    processes, so series built while it is installed run serially in this process."""
    global _PROVIDER, _PROVIDER_IS_CUSTOM
    _PROVIDER = provider
    _PROVIDER_IS_CUSTOM = provider is not None
    _drop_series_pool()


def provider_is_custom() -> bool:
    return _PROVIDER_IS_CUSTOM


def _kernel_dir_has_kernels(path: str | None) -> bool:
    if not path:
        return False
    path = os.path.expanduser(path)
    if not os.path.isdir(path):
        return False
    for _root, _dirs, files in os.walk(path):
        if any(not f.startswith('.') for f in files):
            return True
    return False


def get_default_provider():
    """Ephemeris provider for the once-per-frame host constants: spiceypy when it is
    installed (the reference's setup), else MiniSpice over the kernel directory, else
    MiniSpice over the bundled ephemeris extract (planetmapper_b200/data); MiniSpice's two hot
    primitives run natively (minispice/native.py) when libpm_b200.so is built."""
    global _PROVIDER
    if _PROVIDER is not None:
        return _PROVIDER
    kernel_path = get_kernel_path()
    have_kernels = _kernel_dir_has_kernels(kernel_path)
    if have_kernels:   # spiceypy with nothing furnished cannot answer str2et / spkssb: use the extract then
        try:
            import spiceypy  # noqa: F401

            from .spice_host import SpiceProvider

            _PROVIDER = SpiceProvider(kernel_path)
            return _PROVIDER
        except ImportError:
            pass
    from .minispice import MiniSpice

    if have_kernels:
        _PROVIDER = MiniSpice.from_kernel_dir(os.path.expanduser(kernel_path))
    else:
        _PROVIDER = MiniSpice.from_extract(os.path.join(_DATA_DIR, 'ephem_extract.npz'),
                                           os.path.join(_DATA_DIR, 'pck_pool.json'))
    # same tables, the two hot primitives evaluated by the library's host functions (~3x faster frames)
    try:
        from .minispice.native import NativeSpice

        _PROVIDER = NativeSpice.from_minispice(_PROVIDER)
    except Exception:   # library not built yet: the pure-Python reader gives the same numbers
        pass
    return _PROVIDER
