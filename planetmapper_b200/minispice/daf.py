"""
Minimal reader for NAIF DAF/SPK ephemeris files (segment types 2 and 3) and the
compact "ephemeris extract" this repo ships for machines without the SPK files.

This is HOST-side, once-per-frame scalar code (the part of the path north_star
keeps on the host). It stands in for ``spiceypy.spkssb`` / ``spkezr`` when
spiceypy is not installed; the reference gets the same numbers from CSPICE
(reference call sites: planetmapper/base.py:828 ``spice.spkezr``).

Formats follow the public NAIF "DAF Required Reading" / "SPK Required Reading":
type 2 = Chebyshev position only (velocity by differentiation), type 3 =
Chebyshev position and velocity. Records are ``MID, RADIUS, coeffs...``; the
segment trailer is ``INIT, INTLEN, RSIZE, N``.
"""

from __future__ import annotations

import os
import struct
from dataclasses import dataclass

import numpy as np


@dataclass
class Segment:
    target: int
    center: int
    frame: int
    spk_type: int
    et_start: float
    et_end: float
    init: float
    intlen: float
    rsize: int
    n: int
    records: np.ndarray  # (n_kept, rsize) float64, native endian
    first_record: int = 0  # index (in the full segment) of records[0]
    source: str = ''

    def covers(self, et: float) -> bool:
        if not self.et_start <= et <= self.et_end:
            return False
        idx = self._record_index(et)
        return self.first_record <= idx < self.first_record + len(self.records)

    def _record_index(self, et: float) -> int:
        idx = int((et - self.init) // self.intlen)
        return min(max(idx, 0), self.n - 1)

    def state(self, et: float) -> np.ndarray:
        """Return the 6-state (km, km/s) of target relative to center at et."""
        idx = self._record_index(et) - self.first_record
        rec = self.records[idx]
        mid, radius = rec[0], rec[1]
        s = (et - mid) / radius
        if self.spk_type == 2:
            ncoef = (self.rsize - 2) // 3
            coefs = rec[2:].reshape(3, ncoef)
            pos, dpos = _cheby_with_derivative(coefs, s)
            return np.concatenate([pos, dpos / radius])
        if self.spk_type == 3:
            ncoef = (self.rsize - 2) // 6
            coefs = rec[2:].reshape(6, ncoef)
            val, _ = _cheby_with_derivative(coefs, s)
            return val
        raise NotImplementedError(f'SPK type {self.spk_type} is not supported')


def _cheby_with_derivative(coefs: np.ndarray, s: float) -> tuple[np.ndarray, np.ndarray]:
    """Evaluate sum_k c_k T_k(s) and its derivative d/ds for each row of coefs."""
    n = coefs.shape[1]
    s = float(s)
    # the recurrences run on Python floats (same IEEE arithmetic as float64 scalars, a fraction of
    # the per-element cost of indexing an ndarray); only the two matrix-vector products use numpy
    t = [1.0, s][:n]
    dt = [0.0, 1.0][:n]
    two_s = 2.0 * s
    for k in range(2, n):
        t.append(two_s * t[k - 1] - t[k - 2])
        dt.append(2.0 * t[k - 1] + two_s * dt[k - 1] - dt[k - 2])
    return coefs @ np.array(t), coefs @ np.array(dt)


def read_spk(path: str, keep_types: tuple[int, ...] = (2, 3)) -> list[Segment]:
    """Read every supported segment of a DAF/SPK file (either endianness)."""
    with open(path, 'rb') as f:
        data = f.read()
    if data[:7] != b'DAF/SPK' and data[:8] != b'NAIF/DAF':
        raise ValueError(f'{path!r} is not a DAF/SPK file')
    bff = data[88:96]
    if bff.startswith(b'LTL-IEEE'):
        e = '<'
    elif bff.startswith(b'BIG-IEEE'):
        e = '>'
    else:  # pre-N0052 files carry no format string: pick the order where ND == 2
        e = '<' if struct.unpack('<i', data[8:12])[0] == 2 else '>'
    nd, ni = struct.unpack(e + 'ii', data[8:16])
    fward, _bward, _free = struct.unpack(e + 'iii', data[76:88])
    if (nd, ni) != (2, 6):
        raise ValueError(f'unexpected DAF summary format ND={nd} NI={ni}')
    dbl = np.dtype(e + 'f8')
    segments: list[Segment] = []
    rec = fward
    while rec:
        off = (rec - 1) * 1024
        nxt, _prv, nsum = struct.unpack(e + 'ddd', data[off : off + 24])
        for i in range(int(nsum)):
            s = data[off + 24 + i * 40 : off + 24 + (i + 1) * 40]
            et0, et1 = struct.unpack(e + 'dd', s[:16])
            target, center, frame, spk_type, a0, a1 = struct.unpack(e + '6i', s[16:40])
            if spk_type not in keep_types:
                continue
            # addresses are 1-based double-precision word indices
            trailer = np.frombuffer(data, dtype=dbl, count=4, offset=(a1 - 4) * 8)
            init, intlen, rsize, n = (float(trailer[0]), float(trailer[1]),
                                      int(trailer[2]), int(trailer[3]))
            recs = np.frombuffer(data, dtype=dbl, count=rsize * n, offset=(a0 - 1) * 8)
            recs = recs.astype(np.float64).reshape(n, rsize)
            segments.append(
                Segment(target, center, frame, spk_type, et0, et1, init, intlen,
                        rsize, n, recs, 0, os.path.basename(path))
            )
        rec = int(nxt)
    return segments


def window_segment(seg: Segment, et_lo: float, et_hi: float) -> Segment | None:
    """Cut a segment down to the records that cover [et_lo, et_hi]."""
    lo = max(et_lo, seg.et_start)
    hi = min(et_hi, seg.et_end)
    if lo > hi:
        return None
    i0 = seg._record_index(lo)
    i1 = seg._record_index(hi)
    return Segment(seg.target, seg.center, seg.frame, seg.spk_type, seg.et_start,
                   seg.et_end, seg.init, seg.intlen, seg.rsize, seg.n,
                   seg.records[i0 - seg.first_record : i1 - seg.first_record + 1].copy(),
                   i0, seg.source)


def save_extract(path: str, segments: list[Segment]) -> None:
    """Serialise segments (in load order) to a compact .npz."""
    meta = np.array(
        [[s.target, s.center, s.frame, s.spk_type, s.et_start, s.et_end, s.init,
          s.intlen, s.rsize, s.n, s.first_record, len(s.records)] for s in segments],
        dtype=np.float64,
    )
    arrays = {f'rec{i}': s.records for i, s in enumerate(segments)}
    sources = np.array([s.source for s in segments])
    np.savez_compressed(path, meta=meta, sources=sources, **arrays)


def load_extract(path: str) -> list[Segment]:
    z = np.load(path, allow_pickle=False)
    meta = z['meta']
    sources = z['sources']
    out = []
    for i, m in enumerate(meta):
        out.append(
            Segment(int(m[0]), int(m[1]), int(m[2]), int(m[3]), float(m[4]), float(m[5]),
                    float(m[6]), float(m[7]), int(m[8]), int(m[9]), z[f'rec{i}'],
                    int(m[10]), str(sources[i]))
        )
    return out
