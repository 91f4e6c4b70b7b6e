"""
``NativeSpice``: MiniSpice with its two hot primitives - ``ssb_state`` and ``orientation`` -
evaluated by the host functions of libpm_b200.so (``pm_host_ssb_state`` /
``pm_host_orientation``, planetmapper_b200/csrc/host_ephem.cu) instead of Python.

Same tables, same formulas in the same order: the results agree with the Python reader to the last
bit or two (tests/test_minispice.py), a frame's constants cost ~0.3 ms instead of ~1 ms.  Everything
else (kernel loading, constants pool, time conversion, error messages) is inherited.
"""
from __future__ import annotations

import ctypes

import numpy as np

from .. import _lib
from .core import MiniSpice

MAX_NUT = 64


class _Segment(ctypes.Structure):
    _fields_ = [('target', ctypes.c_int32), ('center', ctypes.c_int32), ('spk_type', ctypes.c_int32),
                ('rsize', ctypes.c_int32), ('n', ctypes.c_int32), ('first_record', ctypes.c_int32),
                ('n_kept', ctypes.c_int32), ('pad', ctypes.c_int32), ('et_start', ctypes.c_double),
                ('et_end', ctypes.c_double), ('init', ctypes.c_double), ('intlen', ctypes.c_double),
                ('rec_offset', ctypes.c_int64)]


class _Orientation(ctypes.Structure):
    _fields_ = [('pole_ra', ctypes.c_double * 3), ('pole_dec', ctypes.c_double * 3), ('pm', ctypes.c_double * 3),
                ('n_ra', ctypes.c_int32), ('n_dec', ctypes.c_int32), ('n_pm', ctypes.c_int32),
                ('n_nut', ctypes.c_int32), ('n_nut_ra', ctypes.c_int32), ('n_nut_dec', ctypes.c_int32),
                ('n_nut_pm', ctypes.c_int32), ('pad', ctypes.c_int32), ('nut_ra', ctypes.c_double * MAX_NUT),
                ('nut_dec', ctypes.c_double * MAX_NUT), ('nut_pm', ctypes.c_double * MAX_NUT),
                ('nut_angles', ctypes.c_double * (2 * MAX_NUT))]


class NativeSpice(MiniSpice):
    name = 'minispice-native'

    def __init__(self, segments, pool):
        super().__init__(segments, pool)
        self._lib = _lib.load_library()
        # precedence order = load order reversed (later kernels win), J2000 segments only: the
        # Python reader raises for other frames, here they are simply not offered
        ordered = [s for s in reversed(segments) if s.frame == 1 and s.spk_type in (2, 3)]
        self._other_frames = {s.target for s in segments if s.frame != 1}
        self._segs = (_Segment * max(len(ordered), 1))()
        blobs, offset = [], 0
        for i, s in enumerate(ordered):
            rec = np.ascontiguousarray(s.records, dtype=np.float64)
            self._segs[i] = _Segment(s.target, s.center, s.spk_type, s.rsize, s.n, s.first_record, rec.shape[0], 0,
                                     s.et_start, s.et_end, s.init, s.intlen, offset)
            blobs.append(rec.reshape(-1))
            offset += rec.size
        self._n_segs = len(ordered)
        self._records = np.concatenate(blobs) if blobs else np.zeros(1)
        self._rec_ptr = self._records.ctypes.data_as(ctypes.c_void_p)
        self._models: dict[int, _Orientation | None] = {}
        self._state = (ctypes.c_double * 6)()
        self._rmat = (ctypes.c_double * 9)()
        self._omega = (ctypes.c_double * 3)()

    @classmethod
    def from_minispice(cls, ms: MiniSpice) -> 'NativeSpice':
        return cls(ms.segments, ms.pool)

    # ---- ephemeris ---------------------------------------------------------------------
    def ssb_state(self, body: int, et: float) -> np.ndarray:
        if self._other_frames:   # segments in other frames: precedence and errors are the Python reader's business
            return super().ssb_state(body, et)
        rc = self._lib.pm_host_ssb_state(self._segs, self._n_segs, self._rec_ptr, int(body), float(et), self._state)
        if rc != 0:
            return super().ssb_state(body, et)   # raises the reader's own LookupError / NotImplementedError
        return np.array(self._state)

    # ---- orientation -------------------------------------------------------------------
    def _model(self, body: int):
        if body not in self._models:
            try:
                m = _Orientation()
                for field, count, key in (('pole_ra', 'n_ra', 'POLE_RA'), ('pole_dec', 'n_dec', 'POLE_DEC'),
                                          ('pm', 'n_pm', 'PM')):
                    vals = [float(v) for v in self.pool[f'BODY{body}_{key}']]
                    if len(vals) > 3:
                        raise ValueError('polynomial too long')
                    setattr(m, count, len(vals))
                    for i, v in enumerate(vals):
                        getattr(m, field)[i] = v
                series = {k: self.pool.get(f'BODY{body}_NUT_PREC_{k}') for k in ('RA', 'DEC', 'PM')}
                if any(v is not None for v in series.values()):
                    bary = body // 100 if body >= 100 else body
                    ang = [float(v) for v in self.pool[f'BODY{bary}_NUT_PREC_ANGLES']]
                    m.n_nut = len(ang) // 2
                    if m.n_nut > MAX_NUT:
                        raise ValueError('too many nutation terms')
                    for i, v in enumerate(ang[:2 * m.n_nut]):
                        m.nut_angles[i] = v
                    for field, count, key in (('nut_ra', 'n_nut_ra', 'RA'), ('nut_dec', 'n_nut_dec', 'DEC'),
                                              ('nut_pm', 'n_nut_pm', 'PM')):
                        vals = [float(v) for v in (series[key] or [])]
                        if len(vals) > m.n_nut:
                            raise ValueError('series longer than the angle list')
                        setattr(m, count, len(vals))
                        for i, v in enumerate(vals):
                            getattr(m, field)[i] = v
                self._models[body] = m
            except (KeyError, ValueError):
                self._models[body] = None    # let the Python reader handle (and report) it
        return self._models[body]

    def orientation(self, body: int, et: float):
        m = self._model(int(body))
        if m is None:
            return super().orientation(body, et)
        rc = self._lib.pm_host_orientation(ctypes.byref(m), float(et), self._rmat, self._omega)
        if rc != 0:
            return super().orientation(body, et)
        return np.array(self._rmat).reshape(3, 3), np.array(self._omega)
