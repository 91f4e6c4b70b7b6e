"""
``Observation``: data cube + BodyXY, mirroring the mapping half of the reference's
``planetmapper.Observation`` (planetmapper/observation.py:826-905:
``get_mapped_data`` / ``_get_mapped_data``).  FITS / PNG I/O, WCS and disc fitting
are out of scope (north_star: "FITS I/O ... untouched"), so the cube is passed in
as an array.

Where the reference loops ``map_img`` over wavelength planes in Python
(observation.py:892-905), this class hands the whole cube to ONE gather launch
(plus the NaN-repair / B-spline prefilter launches for spline modes).
"""
from __future__ import annotations

import numpy as np

from .body_xy import BodyXY


class Observation(BodyXY):
    def __init__(self, *, data, **kwargs) -> None:
        data = np.asarray(data)
        if data.ndim == 2:
            data = data[None]
        if data.ndim != 3:
            raise ValueError('data must be a 2D image or 3D cube')
        self.data = data
        kwargs.pop('nx', None)
        kwargs.pop('ny', None)
        kwargs.pop('sz', None)
        super().__init__(nx=data.shape[2], ny=data.shape[1], **kwargs)
        self._data_dev = None

    def _get_data_device(self):
        from . import _lib as L

        if self._data_dev is None:
            self._data_dev = L.to_device(np.asarray(self.data, dtype=np.float64))
        return self._data_dev

    def get_mapped_data(self, interpolation='linear', *, spline_smoothing: float = 0,
                        propagate_nan: bool = True, warn_nan: bool = False,
                        smooth_oversample_by: int = 5,
                        smooth_max_oversampled_img_size: int = 10_000,
                        **map_kwargs) -> np.ndarray:
        """Map every plane of ``data`` (copy returned; observation.py:826-874)."""
        return np.array(self._get_mapped_data(
            interpolation, spline_smoothing=spline_smoothing, propagate_nan=propagate_nan,
            warn_nan=warn_nan, smooth_oversample_by=smooth_oversample_by,
            smooth_max_oversampled_img_size=smooth_max_oversampled_img_size, **map_kwargs), copy=True)

    def _get_mapped_data(self, interpolation, *, spline_smoothing, propagate_nan, warn_nan,
                         smooth_oversample_by=5, smooth_max_oversampled_img_size=10_000,
                         **map_kwargs) -> np.ndarray:
        # alt-keyed clearable cache (observation.py:876-890)
        key = ('mapped_data', repr(interpolation), spline_smoothing, propagate_nan, smooth_oversample_by,
               smooth_max_oversampled_img_size, self._map_key(map_kwargs), self._alt_adjustment)
        if key not in self._cache:
            out = self.map_img_device(self._get_data_device(), interpolation=interpolation,
                                      spline_smoothing=spline_smoothing,
                                      propagate_nan=propagate_nan, warn_nan=warn_nan,
                                      smooth_oversample_by=smooth_oversample_by,
                                      smooth_max_oversampled_img_size=smooth_max_oversampled_img_size,
                                      **map_kwargs)
            self._cache[key] = out.cpu().numpy()
        return self._cache[key]

    def get_mapped_data_device(self, interpolation='linear', *, propagate_nan: bool = True,
                               planes: slice | None = None, out=None, **map_kwargs):
        """Device-resident variant for cubes whose mapped output does not fit on the
        host (e.g. 3000 x 1800 x 3600 float64 = 155 GB): maps ``data[planes]`` into
        ``out`` (a CUDA tensor) without any device->host copy."""
        cube = self._get_data_device()
        if planes is not None:
            cube = cube[planes]
        return self.map_img_device(cube, interpolation=interpolation,
                                   propagate_nan=propagate_nan, out=out, **map_kwargs)
