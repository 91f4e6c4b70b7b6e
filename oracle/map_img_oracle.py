"""
CPU oracle for the image -> map resampling half of the path (TEST INFRASTRUCTURE
ONLY - see oracle/pm_oracle.c).

Unlike CSPICE / PROJ, the third-party library behind this half of the reference IS
installed here: scipy (1.18.x; the reference pins scipy<=1.18.0, requirements.txt:6).
So this oracle calls the very same scipy routines the reference calls, arranged
exactly as the reference arranges them:

- ``map_img`` follows BodyXY.map_img (planetmapper/body_xy.py:1571-1631) given
  x_map / y_map arrays,
- ``_do_nearest_interpolation``   body_xy.py:1633-1649
- ``_do_spline_interpolation``    body_xy.py:1651-1702 (RectBivariateSpline + .ev)
- ``_should_propagate_nan_to_map`` body_xy.py:1855-1866
- ``_replace_nans_with_interpolated_values`` body_xy.py:1871-1904
- ``get_mapped_data`` follows Observation._get_mapped_data
  (planetmapper/observation.py:876-905): plane by plane.
"""

from __future__ import annotations

import math

import numpy as np
import scipy.interpolate
import scipy.ndimage


def replace_nans_with_interpolated_values(img: np.ndarray) -> np.ndarray:
    bad = ~np.isfinite(img)
    cleaned = img.astype(float, copy=True)
    if np.any(np.isinf(img)):
        img = np.nan_to_num(img, nan=np.nan, posinf=np.nan, neginf=np.nan, copy=True)
    if np.all(bad):
        median = 0.0
    else:
        median = np.nanmedian(img)
    cleaned[bad] = median
    to_fix = bad & ~scipy.ndimage.uniform_filter(bad, size=3)
    for i, j in np.argwhere(to_fix):
        cleaned[i, j] = np.nanmean(img[max(i - 1, 0):i + 2, max(j - 1, 0):j + 2])
    return cleaned


def should_propagate_nan_to_map(x, y, nans, nx, ny) -> bool:
    if x < 0.0 or y < 0.0 or x > nx - 1 or y > ny - 1:
        return True
    x0 = max(math.floor(x), 0)
    x1 = min(math.ceil(x), nx - 1)
    y0 = max(math.floor(y), 0)
    y1 = min(math.ceil(y), ny - 1)
    return bool(nans[y0, x0] or nans[y0, x1] or nans[y1, x0] or nans[y1, x1])


def map_img(img, x_map, y_map, interpolation='nearest', propagate_nan=True,
            spline_smoothing=0.0, smooth_oversample_by=5,
            smooth_max_oversampled_img_size=10_000) -> np.ndarray:
    img = np.asarray(img)
    if img.ndim == 3:
        return np.array([map_img(s, x_map, y_map, interpolation, propagate_nan,
                                 spline_smoothing, smooth_oversample_by,
                                 smooth_max_oversampled_img_size) for s in img])
    ny, nx = img.shape
    projected = np.full(x_map.shape, np.nan)
    spline_k = {'linear': 1, 'quadratic': 2, 'cubic': 3}
    if interpolation in spline_k:
        interpolation = spline_k[interpolation]
    if interpolation == 'nearest':
        nan_sentinel = -999
        xm = np.asarray(np.nan_to_num(np.round(x_map), nan=nan_sentinel), dtype=int)
        ym = np.asarray(np.nan_to_num(np.round(y_map), nan=nan_sentinel), dtype=int)
        for a in range(projected.shape[0]):
            for b in range(projected.shape[1]):
                x = xm[a, b]
                if x == nan_sentinel:
                    continue
                projected[a, b] = img[ym[a, b], x]
        return projected
    if interpolation == 'smooth':
        return _smooth(img, x_map, y_map, projected, propagate_nan, smooth_oversample_by,
                       smooth_max_oversampled_img_size)
    if isinstance(interpolation, int):
        kx = ky = interpolation
    else:
        kx, ky = interpolation
    nans = np.isnan(img)
    if np.all(nans):
        return projected
    cleaned = replace_nans_with_interpolated_values(img)
    interpolator = scipy.interpolate.RectBivariateSpline(
        np.arange(ny), np.arange(nx), cleaned, kx=kx, ky=ky, s=spline_smoothing)
    a_vals, b_vals, x_vals, y_vals = [], [], [], []
    for a in range(projected.shape[0]):
        for b in range(projected.shape[1]):
            x = x_map[a, b]
            if math.isnan(x):
                continue
            y = y_map[a, b]
            if propagate_nan and should_propagate_nan_to_map(x, y, nans, nx, ny):
                continue
            a_vals.append(a)
            b_vals.append(b)
            x_vals.append(x)
            y_vals.append(y)
    if a_vals:
        projected[a_vals, b_vals] = interpolator.ev(y_vals, x_vals)
    return projected


def smooth_grid(n, limits, oversample_by, max_size, limit_padding=5.0):
    """get_xy_pchip of BodyXY._do_smooth_interpolation (body_xy.py:1723-1741): the oversampled
    1-D grid.  Returns (first original index, last original index, grid)."""
    original = np.arange(n)
    original = original[(original >= limits[0] - limit_padding) & (original <= limits[1] + limit_padding)]
    old_size = len(original)
    for oversample_to_use in range(oversample_by, 1, -1):
        new_size = old_size * oversample_to_use - (oversample_to_use - 1)
        if new_size <= max_size:
            return original[0], original[-1], np.linspace(original[0], original[-1], new_size)
    return original[0], original[-1], original.astype(float)


def pchip_grid_interp2d(xs_original, ys_original, img, xs, ys, xlim, ylim, limit_padding):
    """BodyXY._pchip_grid_interp2d (body_xy.py:1792-1853) with the real scipy PCHIP."""
    intermediate = np.full((len(ys_original), len(xs)), np.nan, dtype=np.float64)
    x_mask = (xs_original >= xlim[0] - limit_padding) & (xs_original <= xlim[1] + limit_padding)
    for i, y in enumerate(ys_original):
        if y < ylim[0] - limit_padding or y > ylim[1] + limit_padding:
            continue
        mask = np.isfinite(img[i]) & x_mask
        if np.sum(mask) < 2:
            continue
        interpolator = scipy.interpolate.PchipInterpolator(xs_original[mask], img[i, mask], extrapolate=False)
        intermediate[i] = interpolator(xs)
    final = np.full((len(ys), len(xs)), np.nan, dtype=np.float64)
    y_mask = (ys_original >= ylim[0] - limit_padding) & (ys_original <= ylim[1] + limit_padding)
    for j, x in enumerate(xs):
        if x < xlim[0] - limit_padding or x > xlim[1] + limit_padding:
            continue
        mask = np.isfinite(intermediate[:, j]) & y_mask
        if np.sum(mask) < 2:
            continue
        interpolator = scipy.interpolate.PchipInterpolator(ys_original[mask], intermediate[mask, j],
                                                           extrapolate=False)
        final[:, j] = interpolator(ys)
    return final


def _smooth(img, x_map, y_map, projected, propagate_nan, oversample_by, max_size, limit_padding=5.0):
    """BodyXY._do_smooth_interpolation (body_xy.py:1704-1790): PCHIP oversampling on a regular
    grid followed by linear interpolation, with the real scipy calls."""
    ny, nx = img.shape
    nans = np.isnan(img)
    if np.all(nans) or not np.any(np.isfinite(x_map)):
        return projected
    xlim = (np.nanmin(x_map), np.nanmax(x_map))
    ylim = (np.nanmin(y_map), np.nanmax(y_map))
    xs_original, ys_original = np.arange(nx), np.arange(ny)
    _, _, xs_pchip = smooth_grid(nx, xlim, oversample_by, max_size, limit_padding)
    _, _, ys_pchip = smooth_grid(ny, ylim, oversample_by, max_size, limit_padding)
    pchip_img = pchip_grid_interp2d(xs_original, ys_original, img, xs_pchip, ys_pchip, xlim, ylim, limit_padding)
    interpolator = scipy.interpolate.RegularGridInterpolator(
        (ys_pchip, xs_pchip), pchip_img, bounds_error=False, fill_value=np.nan, method='linear')
    a_vals, b_vals, x_vals, y_vals = [], [], [], []
    for a in range(projected.shape[0]):
        for b in range(projected.shape[1]):
            x = x_map[a, b]
            if math.isnan(x):
                continue
            y = y_map[a, b]
            if propagate_nan and should_propagate_nan_to_map(x, y, nans, nx, ny):
                continue
            a_vals.append(a)
            b_vals.append(b)
            x_vals.append(x)
            y_vals.append(y)
    if a_vals:
        projected[a_vals, b_vals] = interpolator((y_vals, x_vals))
    return projected


def map_cube_fast(cube, x_map, y_map, interpolation, propagate_nan=True):
    """Vectorised equivalent of map_img over a cube, used for the bounded CPU
    baseline in bench.py: identical scipy calls per plane, numpy instead of the
    per-cell Python loop for the bookkeeping."""
    cube = np.asarray(cube, dtype=float)
    nl, ny, nx = cube.shape
    out = np.full((nl,) + x_map.shape, np.nan)
    valid = ~np.isnan(x_map)
    if interpolation == 'nearest':
        xi = np.round(x_map[valid]).astype(int)
        yi = np.round(y_map[valid]).astype(int)
        out[:, valid] = cube[:, yi, xi]
        return out
    k = {'linear': 1, 'cubic': 3}[interpolation]
    xv, yv = x_map[valid], y_map[valid]
    inside = ~((xv < 0.0) | (yv < 0.0) | (xv > nx - 1) | (yv > ny - 1))
    x0 = np.maximum(np.floor(xv), 0).astype(int)
    x1 = np.minimum(np.ceil(xv), nx - 1).astype(int)
    y0 = np.maximum(np.floor(yv), 0).astype(int)
    y1 = np.minimum(np.ceil(yv), ny - 1).astype(int)
    x0c, x1c, y0c, y1c = (np.clip(v, 0, m) for v, m in
                          ((x0, nx - 1), (x1, nx - 1), (y0, ny - 1), (y1, ny - 1)))
    for l in range(nl):
        img = cube[l]
        nans = np.isnan(img)
        if np.all(nans):
            continue
        cleaned = replace_nans_with_interpolated_values(img)
        interp = scipy.interpolate.RectBivariateSpline(np.arange(ny), np.arange(nx), cleaned,
                                                       kx=k, ky=k, s=0)
        keep = np.ones(xv.shape, dtype=bool)
        if propagate_nan:
            keep = inside & ~(nans[y0c, x0c] | nans[y0c, x1c] | nans[y1c, x0c] | nans[y1c, x1c])
        vals = np.full(xv.shape, np.nan)
        if np.any(keep):
            vals[keep] = interp.ev(yv[keep], xv[keep])
        plane = out[l]
        plane[valid] = vals
    return out
