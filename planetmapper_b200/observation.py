"""
``Observation``: data cube + BodyXY, mirroring the mapping half of the reference's
``planetmapper.Observation`` (planetmapper/observation.py:826-905:
``get_mapped_data`` / ``_get_mapped_data``) and its two consumers,
``save_observation`` / ``save_mapped_observation`` (:1185-1474).  READING FITS / PNG
files, WCS fitting and disc fitting are out of scope (north_star: "FITS I/O ...
untouched"), so the cube is passed in as an array; the save methods write the
reference's on-disc format (one float64 HDU per backplane, EXTNAME = backplane name)
from a file image assembled on the device (``fits_stage.py``).

Where the reference loops ``map_img`` over wavelength planes in Python
(observation.py:892-905), this class hands the whole cube to ONE gather launch
(plus the NaN-repair / B-spline prefilter launches for spline modes).
"""
from __future__ import annotations

import datetime
import os
from typing import Any, Collection

import numpy as np

from . import _lib as L
from .body_xy import BodyXY
from .fits_stage import Header, ImageHDU, write_hdus
from .progress import CLIProgressHook, progress_decorator

_REFERENCE_URL = 'https://github.com/ortk95/planetmapper'


class Observation(BodyXY):
    FITS_KEYWORD = 'PLANMAP'  # observation.py:139

    def __init__(self, *, data, header=None, **kwargs) -> None:
        data = np.asarray(data)
        if data.ndim == 2:
            data = data[None]
        if data.ndim != 3:
            raise ValueError('data must be a 2D image or 3D cube')
        self.data = data
        # observation.py:190-200: the header of the input file, or an empty one
        if header is None:
            header = Header()
        elif not isinstance(header, Header):
            header = Header([(k, v, None) for k, v in dict(header).items()])
        self.header = header
        self.path = None  # loading from a file is out of scope
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

    @progress_decorator
    def _get_mapped_data(self, interpolation, *, spline_smoothing, propagate_nan, warn_nan,
                         smooth_oversample_by=5, smooth_max_oversampled_img_size=10_000,
                         **map_kwargs) -> np.ndarray:
        # alt-keyed clearable cache (observation.py:876-890)
        key = ('mapped_data', repr(interpolation), spline_smoothing, propagate_nan, smooth_oversample_by,
               smooth_max_oversampled_img_size, self._map_key(map_kwargs), self._alt_adjustment)
        if key not in self._cache:
            # the reference's loop over wavelength planes (observation.py:892-905), in chunks of as many planes
            # as fit on the device next to their own double buffer
            src = self._map_source(self._get_data_device(), interpolation=interpolation,
                                   spline_smoothing=spline_smoothing, propagate_nan=propagate_nan,
                                   warn_nan=warn_nan, smooth_oversample_by=smooth_oversample_by,
                                   smooth_max_oversampled_img_size=smooth_max_oversampled_img_size, **map_kwargs)
            host = L.empty_host((src.n_planes,) + src.map_shape)
            for first, count, _ in self._stream_mapped_chunks(src, None, into=host):
                self._update_progress_hook((first + count) / src.n_planes)   # observation.py:903
            self._cache[key] = host.numpy()
        return self._cache[key]

    @staticmethod
    def _planes_per_chunk(src, planes_per_chunk=None, budget_bytes=None) -> int:
        """Planes per gather launch: the caller's choice, else what fits twice (double buffer) into 40 % of
        the free device memory; a multiple of 4 (the spline operands are stored as plane quads)."""
        torch = L._torch()
        n_cells = 1
        for d in src.map_shape:
            n_cells *= int(d)
        if planes_per_chunk is None:
            if budget_bytes is None:
                free, _total = torch.cuda.mem_get_info()
                budget_bytes = int(0.4 * free)
            planes_per_chunk = budget_bytes // max(1, 2 * 8 * n_cells)
        planes_per_chunk = int(min(max(planes_per_chunk, 4), max(src.n_planes, 4)))
        return max(4, planes_per_chunk // 4 * 4)

    def _stream_mapped_chunks(self, src, planes_per_chunk, into=None, host_buffers=None):
        """Generator over (first_plane, count, host_view): maps the cube chunk by chunk and moves every chunk
        to the host while the NEXT chunk is being gathered (two device buffers, a copy stream, events).
        ``into``: a CPU tensor (n_planes, ...) receiving the whole result (host_view is then its slice);
        otherwise two pinned staging buffers are cycled and host_view is only valid until the next step."""
        torch = L._torch()
        n = src.n_planes
        if n == 0:
            return
        per = self._planes_per_chunk(src, planes_per_chunk)
        n_buf = 2 if n > per else 1
        if into is None and host_buffers is None:
            # the library's cached pinned pair (pinning 1 GB costs ~1 s: never inside the stream)
            with L.staging_buffers((min(per, n),) + src.map_shape, n_buf) as bufs:
                yield from self._stream_mapped_chunks(src, per, host_buffers=bufs)
            return
        dev = [torch.empty((min(per, n),) + src.map_shape, dtype=torch.float64, device='cuda') for _ in range(n_buf)]
        compute = torch.cuda.current_stream()
        copy_stream = torch.cuda.Stream()
        gathered = [torch.cuda.Event() for _ in dev]
        copied = [torch.cuda.Event() for _ in dev]
        pending = None   # (first, count, host_view, event) of the chunk whose copy is in flight
        for k, first in enumerate(range(0, n, per)):
            count = min(per, n - first)
            b = k % len(dev)
            if k >= len(dev):
                compute.wait_event(copied[b])       # the buffer's previous contents have left the device
            src.gather(first, count, out=dev[b][:count])
            gathered[b].record(compute)
            view = into[first:first + count] if into is not None else host_buffers[b][:count]
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(gathered[b])
                view.copy_(dev[b][:count], non_blocking=True)
                copied[b].record(copy_stream)
            if pending is not None:
                pending[3].synchronize()
                yield pending[:3]
            pending = (first, count, view, copied[b])
        pending[3].synchronize()
        yield pending[:3]

    def iter_mapped_data(self, interpolation='linear', *, planes_per_chunk: int | None = None,
                         spline_smoothing: float = 0, propagate_nan: bool = True, warn_nan: bool = False,
                         smooth_oversample_by: int = 5, smooth_max_oversampled_img_size: int = 10_000,
                         **map_kwargs):
        """``get_mapped_data`` for cubes whose mapped output does not fit in host memory (BASELINE config C4:
        3000 x 1800 x 3600 float64 = 155 GB): yields ``(first_plane, mapped)`` with ``mapped`` a float64 array
        (count,) + map shape holding planes first .. first + count.  The array is a view of a pinned staging
        buffer that is REUSED two iterations later: consume or copy it before advancing twice.  Chunk k + 1 is
        gathered while chunk k crosses PCIe."""
        src = self._map_source(self._get_data_device(), interpolation=interpolation,
                               spline_smoothing=spline_smoothing, propagate_nan=propagate_nan, warn_nan=warn_nan,
                               smooth_oversample_by=smooth_oversample_by,
                               smooth_max_oversampled_img_size=smooth_max_oversampled_img_size, **map_kwargs)
        for first, count, view in self._stream_mapped_chunks(src, planes_per_chunk):
            yield first, view.numpy()

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

    # ---- FITS header metadata (observation.py:908-1157) ------------------------------------
    def append_to_header(self, keyword: str, value, comment: str | None = None,
                         hierarch_keyword: bool = True, header: Header | None = None,
                         truncate_strings: bool = True, remove_existing: bool = True) -> None:
        """Add a card to ``header`` (default: :attr:`header`), observation.py:908-950."""
        if header is None:
            header = self.header
        if hierarch_keyword:
            keyword = self._make_fits_kw(keyword)
        if truncate_strings and isinstance(value, str):
            if len(keyword) + len(value) + 4 > 80:
                n = 80 - len(keyword) - 4 - 3
                value = value[:n] + '...'
        if remove_existing:
            header.remove(keyword, ignore_missing=True, remove_all=True)
        header.append(keyword, value, comment)

    @classmethod
    def _make_fits_kw(cls, keyword: str) -> str:
        return f'HIERARCH {cls.FITS_KEYWORD} {keyword}'

    def add_header_metadata(self, header: Header | None = None) -> None:
        """The PLANMAP metadata cards, in the reference's order (observation.py:956-1157)."""
        from . import __version__

        bc = self._bc
        cards = [
            ('VERSION', __version__, 'PlanetMapper version.'),
            ('URL', _REFERENCE_URL, 'Webpage.'),
            ('DATE', datetime.datetime.now().strftime('%Y-%m-%dT%H:%M:%S'), 'File generation datetime.'),
        ]
        if self.path is not None:
            cards.append(('INFILE', os.path.split(self.path)[1], 'Input file name.'))
        cards += [
            ('DISC X0', self.get_x0(), '[pixels] x coordinate of disc centre.'),
            ('DISC Y0', self.get_y0(), '[pixels] y coordinate of disc centre.'),
            ('DISC R0', self.get_r0(), '[pixels] equatorial radius of disc.'),
            ('DISC ROT', self.get_rotation(), '[degrees] rotation of image.'),
            ('DISC METHOD', self.get_disc_method(), 'Method used to find disc.'),
            ('ALTITUDE-ADJUSTMENT', self._alt_adjustment, '[km] Adjustment to surface altitude.'),
            ('UTC-OBS', self.utc, 'UTC date of observation'),
            ('ET-OBS', self.et, 'J2000 ephemeris seconds of observation.'),
            ('TARGET', self.target, 'Target body name used in SPICE.'),
            ('TARGET-ID', self.target_body_id, 'Target body ID from SPICE.'),
            ('SUBPOINT LAT', self.subpoint_lat, '[degrees] Sub-observer pgr latitude.'),
            ('SUBPOINT LON', self.subpoint_lon, '[degrees] Sub-observer pgr longitude.'),
            ('SUBSOL LAT', self.subsol_lat, '[degrees] Sub-solar pgr latitude.'),
            ('SUBSOL LON', self.subsol_lon, '[degrees] Sub-solar pgr longitude.'),
            ('LON-DIRECTION', self.positive_longitude_direction, 'Positive pgr longitude direction.'),
            ('NP-ANGLE', self.north_pole_angle(), '[degrees] North pole angle.'),
            ('TARGET RA', self.target_ra, '[degrees] RA of target centre.'),
            ('TARGET DEC', self.target_dec, '[degrees] Dec of target centre.'),
            ('TARGET DIAMETER', self.target_diameter_arcsec, '[arcsec] Equatorial angular diameter of target.'),
            ('R EQ', self.r_eq, '[km] Target equatorial radius from SPICE.'),
            ('R POLAR', self.r_polar, '[km] Target polar radius from SPICE.'),
            ('FLATTENING', self.flattening, 'Flattening of target body.'),
            ('LIGHT-TIME', self.target_light_time, '[seconds] Light time to target from SPICE.'),
            ('DISTANCE', self.target_distance, '[km] Distance to target from SPICE.'),
            ('OBSERVER', self.observer, 'Observer name used in SPICE.'),
            ('TARGET-FRAME', 'IAU_' + bc.target, 'Target frame used in SPICE.'),
            ('OBSERVER-FRAME', 'J2000', 'Observer frame used in SPICE.'),
            ('ILLUMINATION', 'SUN', 'Illumination source used in SPICE.'),
            ('ABCORR', 'CN', 'Aberration correction used in SPICE.'),
            ('SUBPOINT-METHOD', 'INTERCEPT/ELLIPSOID', 'Subpoint method used in SPICE.'),
            ('SURFACE-METHOD', 'ELLIPSOID', 'Surface intercept method used in SPICE.'),
            ('OPTIMIZATION-USED', self._optimize_speed, 'Speed optimizations used.'),
        ]
        for k, v, c in cards:
            self.append_to_header(k, v, c, header=header)

    def make_filename(self, extension: str = '.fits', prefix: str = '', suffix: str = '') -> str:
        """e.g. ``'JUPITER_2000-01-01T123456.fits'`` (observation.py:1159-1182)."""
        return f'{prefix}{self.target}_{self.dtm.strftime("%Y-%m-%dT%H%M%S")}{suffix}{extension}'

    def _add_map_header_metadata(self, header: Header, *, interpolation, spline_smoothing: float,
                                 propagate_nan: bool, smooth_oversample_by: int,
                                 smooth_max_oversampled_img_size: int, **map_kwargs) -> None:
        """observation.py:1476-1571."""
        *_, info = self.generate_map_coordinates(**map_kwargs)
        put = lambda k, v, c: self.append_to_header(k, v, c, header=header)  # noqa: E731
        put('MAP INTERPOLATION', str(interpolation) if isinstance(interpolation, tuple) else interpolation,
            'Interpolation method used in mapping.')
        if interpolation not in {'nearest', 'smooth'}:
            put('MAP SPLINE-SMOOTHING', spline_smoothing, 'Interpolation spline smoothing factor used in mapping.')
            put('MAP PROPAGATE-NAN', propagate_nan, 'Propagate NaN pixels to map when mapping.')
        if interpolation == 'smooth':
            put('MAP SMOOTH-OVERSAMPLE-BY', smooth_oversample_by, 'Oversampling factor used in map interpolation.')
            put('MAP SMOOTH-MAX-OVERSAMPLED-IMG-SIZE', smooth_max_oversampled_img_size,
                'Maximum oversampled image size allowed map interpolation.')
        put('MAP PROJECTION', info['projection'], 'Projection used for mapping.')
        for key, name, comment in (('degree_interval', 'MAP DEGREE-INTERVAL', '[deg] Degree interval in output map.'),
                                   ('lon', 'MAP LON', 'Central longitude of map projection.'),
                                   ('lat', 'MAP LAT', 'Central latitude of map projection.'),
                                   ('size', 'MAP SIZE', 'Size of output map.')):
            if key in info:
                put(name, info[key], comment)

    def _add_map_wcs_to_header(self, header: Header, **map_kwargs) -> None:
        """observation.py:1573-1612."""
        lons, lats, *_, info = self.generate_map_coordinates(**map_kwargs)
        if info['projection'] == 'rectangular':
            header['CTYPE1'] = f'Planetographic longitude, positive {self.positive_longitude_direction}'
            header['CUNIT1'] = 'deg'
            header['CRPIX1'] = 1
            header['CRVAL1'] = float(lons[0][0])
            header['CDELT1'] = float(lons[0][1] - lons[0][0])
            header['CTYPE2'] = 'Planetographic latitude'
            header['CUNIT2'] = 'deg'
            header['CRPIX2'] = 1
            header['CRVAL2'] = float(lats[0][0])
            header['CDELT2'] = float(lats[1][0] - lats[0][0])
        else:
            for n in '12':
                for key in (f'CTYPE{n}', f'CUNIT{n}', f'CRPIX{n}', f'CRVAL{n}', f'CDELT{n}'):
                    header.remove(key, ignore_missing=True, remove_all=True)
        for a in '12':
            for b in '123':
                for key in (f'PC{a}_{b}', f'PC{b}_{a}', f'CD{a}_{b}', f'CD{b}_{a}'):
                    header.remove(key, ignore_missing=True, remove_all=True)

    # ---- saving (observation.py:1185-1474) -------------------------------------------------
    def _get_backplane_names_to_save(self, backplanes_to_save: Collection[str] | None,
                                     backplanes_to_skip: Collection[str]) -> set[str]:
        if backplanes_to_save is None:
            backplanes_to_save = self.backplanes.keys()
        return ({self.standardise_backplane_name(n) for n in backplanes_to_save}
                - {self.standardise_backplane_name(n) for n in backplanes_to_skip})

    def _backplane_hdus(self, names: set[str], *, mapped: bool, print_info: bool, map_kwargs: dict) -> list[ImageHDU]:
        """One ImageHDU per requested backplane, in registration order.  Built-in backplanes
        come from ONE fused launch and stay on the device; user-registered ones are
        called like the reference does and uploaded."""
        wanted = [n for n in self.backplanes if n in names]
        builtin = [n for n in wanted if self._is_builtin_backplane(n, mapped)]
        planes = {}
        if builtin:
            mask = L.mask_from_names(builtin)
            if mapped:
                for have, dev in self.get_backplanes_map_device(mask, **map_kwargs).values():
                    for n in builtin:
                        pid = L.PLANE_ID[n]
                        if have >> pid & 1:
                            planes[n] = dev[L.popcount(have & ((1 << pid) - 1))]
            else:
                have, dev = self.get_backplanes_img_device(mask)
                for n in builtin:
                    pid = L.PLANE_ID[n]
                    planes[n] = dev[L.popcount(have & ((1 << pid) - 1))]
        hdus = []
        for name in wanted:
            backplane = self.backplanes[name]
            if print_info:
                print(' Creating backplane:', name)
            data = planes.get(name)
            if data is None:
                data = backplane.get_map(**map_kwargs) if mapped else backplane.get_img()
            header = Header([('ABOUT', backplane.description, None)])
            header.add_comment('Backplane generated by PlanetMapper software.')
            if mapped:
                self._add_map_wcs_to_header(header, **map_kwargs)
            hdus.append(ImageHDU(data=data, header=header, name=name))
        return hdus

    @staticmethod
    def _reject_wireframe(include_wireframe: bool) -> None:
        if include_wireframe:
            raise NotImplementedError(
                'the WIREFRAME HDU is drawn with matplotlib (plotting is out of scope of the '
                'accelerated path); pass include_wireframe=False')

    @progress_decorator
    def save_observation(self, path, *, backplanes_to_save: Collection[str] | None = None,
                         backplanes_to_skip: Collection[str] = frozenset(), include_wireframe: bool = False,
                         wireframe_kwargs: dict[str, Any] | None = None, show_progress: bool = False,
                         print_info: bool = True, alt: float = 0.0) -> None:
        """Save ``data`` + the generated backplanes as a FITS file (observation.py:1185-1303):
        primary HDU = data and header with the PLANMAP metadata, then one float64 image
        extension per backplane.  ``include_wireframe`` defaults to False here."""
        self._reject_wireframe(include_wireframe)
        path = os.fspath(path)
        names = self._get_backplane_names_to_save(backplanes_to_save, backplanes_to_skip)
        own_hook = show_progress and self._get_progress_hook() is None     # observation.py:1250-1254
        if own_hook:
            self._set_progress_hook(CLIProgressHook())
            print_info = False
        try:
            self._save_observation(path, names, print_info, alt)
        finally:
            if own_hook:
                self._remove_progress_hook()

    def _save_observation(self, path, names, print_info, alt) -> None:
        from .body_xy import _AltitudeScope

        if print_info:
            print('Saving observation to', path)
        with _AltitudeScope(self, self._check_alt(alt)):
            header = self.header.copy()
            self.add_header_metadata(header)
            hdus = [ImageHDU(data=self._get_data_device(), header=header)]
            hdus += self._backplane_hdus(names, mapped=False, print_info=print_info, map_kwargs={})
            if print_info:
                print(' Saving file...')
            write_hdus(path, hdus)
        if print_info:
            print('File saved')

    @progress_decorator
    def save_mapped_observation(self, path, *, interpolation='linear', propagate_nan: bool = True,
                                spline_smoothing: float = 0, smooth_oversample_by: int = 5,
                                smooth_max_oversampled_img_size: int = 10_000, include_backplanes: bool = True,
                                backplanes_to_save: Collection[str] | None = None,
                                backplanes_to_skip: Collection[str] = frozenset(), include_wireframe: bool = False,
                                wireframe_kwargs: dict[str, Any] | None = None, show_progress: bool = False,
                                print_info: bool = True, **map_kwargs) -> None:
        """Save the mapped cube + mapped backplanes as a FITS file (observation.py:1315-1474).
        The mapped cube never visits the host as native-endian data: gather output ->
        ``pm_fits_stage`` -> one copy.  ``include_wireframe`` defaults to False here."""
        self._reject_wireframe(include_wireframe)
        path = os.fspath(path)
        names = self._get_backplane_names_to_save(backplanes_to_save, backplanes_to_skip)
        own_hook = show_progress and self._get_progress_hook() is None     # observation.py:1398-1402
        if own_hook:
            self._set_progress_hook(CLIProgressHook())
            print_info = False
        try:
            self._save_mapped_observation(path, names, print_info, interpolation, propagate_nan, spline_smoothing,
                                          smooth_oversample_by, smooth_max_oversampled_img_size, include_backplanes,
                                          map_kwargs)
        finally:
            if own_hook:
                self._remove_progress_hook()

    def _save_mapped_observation(self, path, names, print_info, interpolation, propagate_nan, spline_smoothing,
                                 smooth_oversample_by, smooth_max_oversampled_img_size, include_backplanes,
                                 map_kwargs) -> None:
        if print_info:
            print('Saving map to', path)
            print(' Projecting mapped data...')
        mapping = dict(interpolation=interpolation, spline_smoothing=spline_smoothing, propagate_nan=propagate_nan,
                       smooth_oversample_by=smooth_oversample_by,
                       smooth_max_oversampled_img_size=smooth_max_oversampled_img_size)
        data = self.map_img_device(self._get_data_device(), warn_nan=False, **mapping, **map_kwargs)
        header = self.header.copy()
        self.add_header_metadata(header)
        self._add_map_header_metadata(header, **mapping, **map_kwargs)
        self._add_map_wcs_to_header(header, **map_kwargs)
        hdus = [ImageHDU(data=data, header=header)]
        if include_backplanes:
            hdus += self._backplane_hdus(names, mapped=True, print_info=print_info, map_kwargs=map_kwargs)
        if print_info:
            print(' Saving file...')
        write_hdus(path, hdus)
        if print_info:
            print('File saved')
