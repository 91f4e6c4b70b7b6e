"""
FITS staging for ``Observation.save_observation`` / ``save_mapped_observation``
(planetmapper/observation.py:1185-1303, :1315-1474).

The reference builds an ``astropy.io.fits.HDUList`` -- a primary HDU holding the data
cube plus one float64 ``ImageHDU`` per backplane -- and lets astropy serialise it.  Here
the arrays are already in HBM, so the file is assembled on the device:

1. the host lays the file out (header blocks and data units are multiples of 2880 bytes)
   and formats the header cards (:class:`Header`, same card text astropy writes for the
   value types PlanetMapper uses: str, bool, int, float, COMMENT, ``HIERARCH`` keywords);
2. ONE ``pm_fits_stage`` launch byte-swaps every array into its big-endian data unit of a
   device-resident image of the file (``planetmapper_b200/csrc/stage_kernels.cu``);
3. ONE device-to-host copy moves the image into pinned memory, which is written to disk.

astropy is not a dependency: only the subset of the FITS standard (4.0, sections 3-5)
that those two methods produce is implemented.  Reading FITS files is out of scope.
"""
from __future__ import annotations

import os
from typing import Any, Iterable

import numpy as np

BLOCK = 2880
CARD = 80


class VerifyError(ValueError):
    """A header card cannot be represented (mirrors astropy.io.fits.VerifyError)."""


def _format_float(value: float) -> str:
    # astropy.io.fits.card._format_float: shortest repr, upper-case exponent, <= 20 chars
    s = str(float(value)).replace('e', 'E')
    if s in ('nan', 'inf', '-inf'):
        raise VerifyError(f'floating point value {s!r} cannot be represented in a FITS header')
    if len(s) > 20:
        idx = s.find('E')
        s = s[:20] if idx < 0 else s[:20 - (len(s) - idx)] + s[idx:]
    return s


def _format_value(value: Any) -> str:
    """The value field before any padding (astropy.io.fits.card._format_value)."""
    if isinstance(value, str):
        if value == '':
            return "''"
        return "'{:8}'".format(value.replace("'", "''"))
    if isinstance(value, (bool, np.bool_)):
        return 'T' if value else 'F'
    if isinstance(value, (int, np.integer)):
        return str(int(value))
    if isinstance(value, (float, np.floating)):
        return _format_float(float(value))
    raise VerifyError(f'unsupported FITS header value type {type(value).__name__}')


class Header:
    """Ordered list of (keyword, value, comment) cards with astropy's card formatting.

    Keywords longer than 8 characters or containing spaces are written as ``HIERARCH``
    cards (what ``Observation.append_to_header`` produces, observation.py:908-954).
    """

    def __init__(self, cards: Iterable[tuple] = ()) -> None:
        self.cards: list[tuple[str, Any, str | None]] = []
        for c in cards:
            self.append(*c)

    def copy(self) -> 'Header':
        h = Header()
        h.cards = list(self.cards)
        return h

    @staticmethod
    def _norm(keyword: str) -> str:
        k = keyword.strip()
        if k.upper().startswith('HIERARCH '):
            k = k[9:].strip()
        return k.upper()

    def append(self, keyword: str, value: Any = None, comment: str | None = None) -> None:
        self.cards.append((self._norm(keyword), value, comment))

    def add_comment(self, text: str) -> None:
        self.cards.append(('COMMENT', text, None))

    def remove(self, keyword: str, ignore_missing: bool = True, remove_all: bool = True) -> None:
        k = self._norm(keyword)
        hits = [i for i, c in enumerate(self.cards) if c[0] == k]
        if not hits and not ignore_missing:
            raise KeyError(keyword)
        for i in reversed(hits if remove_all else hits[:1]):
            del self.cards[i]

    def __setitem__(self, keyword: str, value: Any) -> None:
        # header[key] = value: update in place if present, else append (astropy semantics)
        k = self._norm(keyword)
        comment = None
        if isinstance(value, tuple):
            value, comment = value
        for i, c in enumerate(self.cards):
            if c[0] == k:
                self.cards[i] = (k, value, comment if comment is not None else c[2])
                return
        self.cards.append((k, value, comment))

    def __getitem__(self, keyword: str) -> Any:
        k = self._norm(keyword)
        for c in self.cards:
            if c[0] == k:
                return c[1]
        raise KeyError(keyword)

    def __contains__(self, keyword: str) -> bool:
        k = self._norm(keyword)
        return any(c[0] == k for c in self.cards)

    def keys(self) -> list[str]:
        return [c[0] for c in self.cards]

    @property
    def comments(self) -> '_Comments':
        """``header.comments[keyword]`` like astropy.io.fits.Header.comments."""
        return _Comments(self)

    def __len__(self) -> int:
        return len(self.cards)

    # ---- card text ---------------------------------------------------------------
    @staticmethod
    def format_card(keyword: str, value: Any, comment: str | None) -> list[str]:
        """80-character card image(s) of one header entry."""
        if keyword in ('COMMENT', 'HISTORY', ''):
            text = '' if value is None else str(value)
            out = []
            for i in range(0, max(len(text), 1), 72):
                out.append('{:8}{}'.format(keyword, text[i:i + 72]).ljust(CARD))
            return out
        hierarch = len(keyword) > 8 or ' ' in keyword
        if hierarch:
            head = f'HIERARCH {keyword} '
            delim = '= '
        else:
            if not all(ch.isalnum() or ch in '-_' for ch in keyword):
                raise VerifyError(f'illegal keyword name {keyword!r}')
            head = '{:8}'.format(keyword)
            delim = '= '
        v = _format_value(value)
        if not hierarch:
            # fixed format: numbers and logicals right-justified to column 30
            if not isinstance(value, str):
                v = '{:>20}'.format(v)
            elif comment:
                v = '{:20}'.format(v)
        text = head + delim + v
        if len(text) > CARD:
            if hierarch and len(text) == CARD + 1:
                text = head[:-1] + delim + v  # astropy drops the space before '=' to make it fit
            else:
                raise VerifyError(f'card for {keyword!r} is longer than 80 characters')
        if comment:
            text = (text + ' / ' + comment)[:CARD]  # astropy truncates the comment (with a warning)
        return [text.ljust(CARD)]

    def card_images(self) -> list[str]:
        out = []
        for k, v, c in self.cards:
            out += self.format_card(k, v, c)
        return out


class _Comments:
    def __init__(self, header: Header) -> None:
        self._header = header

    def __getitem__(self, keyword: str) -> str:
        k = Header._norm(keyword)
        for c in self._header.cards:
            if c[0] == k:
                return c[2] or ''
        raise KeyError(keyword)


def _structural_cards(shape: tuple[int, ...], primary: bool, extend: bool) -> list[tuple]:
    naxis = len(shape)
    cards = [('SIMPLE', True, 'conforms to FITS standard')] if primary else [('XTENSION', 'IMAGE', 'Image extension')]
    cards += [('BITPIX', -64, 'array data type'), ('NAXIS', naxis, 'number of array dimensions')]
    cards += [(f'NAXIS{i + 1}', int(n), None) for i, n in enumerate(reversed(shape))]
    if primary:
        if extend:
            cards.append(('EXTEND', True, None))
    else:
        cards += [('PCOUNT', 0, 'number of parameters'), ('GCOUNT', 1, 'number of groups')]
    return cards


_STRUCTURAL = {'SIMPLE', 'XTENSION', 'BITPIX', 'NAXIS', 'EXTEND', 'PCOUNT', 'GCOUNT', 'EXTNAME', 'END'}


def hdu_header_bytes(shape: tuple[int, ...], header: Header | None, *, primary: bool, extend: bool = True,
                     name: str | None = None) -> bytes:
    """Header blocks of a float64 image HDU, cards ordered the way astropy orders them:
    structural keywords, the user's cards, EXTNAME (extensions), commentary cards, END."""
    full = Header(_structural_cards(shape, primary, extend))
    tail = Header()
    for k, v, c in (header.cards if header is not None else []):
        if k in _STRUCTURAL or (k.startswith('NAXIS') and k[5:].isdigit()):
            continue
        (tail if k in ('COMMENT', 'HISTORY') else full).cards.append((k, v, c))
    if name is not None and not primary:
        full.append('EXTNAME', name, 'extension name')
    full.cards += tail.cards
    images = full.card_images() + ['END'.ljust(CARD)]
    text = ''.join(images)
    text += ' ' * (-len(text) % BLOCK)
    return text.encode('ascii')


class ImageHDU:
    """A float64 image HDU whose data is a CUDA tensor (or anything ``to_device`` accepts)."""

    def __init__(self, data, header: Header | None = None, name: str | None = None) -> None:
        self.data = data
        self.header = header if header is not None else Header()
        self.name = name


def file_layout(hdus: list[ImageHDU]) -> tuple[list[bytes], list[int], list[int], int]:
    """(header bytes per HDU, header offsets, data-unit offsets, file size)."""
    headers, hoff, doff = [], [], []
    pos = 0
    for i, h in enumerate(hdus):
        shape = tuple(int(n) for n in h.data.shape)
        hb = hdu_header_bytes(shape, h.header, primary=(i == 0), extend=len(hdus) > 1, name=h.name)
        headers.append(hb)
        hoff.append(pos)
        pos += len(hb)
        doff.append(pos)
        n = int(np.prod(shape)) if len(shape) else 0
        pos += (n * 8 + BLOCK - 1) // BLOCK * BLOCK
    return headers, hoff, doff, pos


def stage_file_image(hdus: list[ImageHDU]):
    """Assemble the whole FITS file in device memory and return it as a pinned uint8 CPU
    tensor (one launch + one device-to-host copy, both on the current stream, synchronised)."""
    from . import _lib as L

    torch = L._torch()
    arrays = []
    for h in hdus:
        d = h.data
        if not isinstance(d, torch.Tensor):
            d = L.to_device(np.ascontiguousarray(d, dtype=np.float64))
        if d.dtype != torch.float64:
            d = d.to(torch.float64)
        h.data = d.contiguous()
        arrays.append(h.data)
    headers, hoff, doff, size = file_layout(hdus)
    image = torch.empty(size, dtype=torch.uint8, device=arrays[0].device)
    host = torch.empty(size, dtype=torch.uint8, pin_memory=True)
    L.fits_stage(arrays, doff, image)
    host.copy_(image, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    # header blocks are host data: written straight into the pinned file image (the
    # device image's header regions are never read)
    view = host.numpy()
    for hb, off in zip(headers, hoff):
        view[off:off + len(hb)] = np.frombuffer(hb, dtype=np.uint8)
    return host


def write_hdus(path: str | os.PathLike, hdus: list[ImageHDU], overwrite: bool = True) -> int:
    """Write the HDUs to ``path``; returns the file size in bytes."""
    path = os.fspath(path)
    if not overwrite and os.path.exists(path):
        raise OSError(f'File {path!r} already exists.')
    host = stage_file_image(hdus)
    directory = os.path.dirname(path)
    if directory:
        os.makedirs(directory, exist_ok=True)  # utils.check_path (planetmapper/utils.py)
    with open(path, 'wb') as f:
        f.write(memoryview(host.numpy()))
    return int(host.numel())
