"""
TEST INFRASTRUCTURE ONLY (like everything under oracle/): CPU statement of how the reference's
save_observation / save_mapped_observation (planetmapper/observation.py:1185-1474) lay a
float64 image HDU out on disc.  The serialisation itself lives in astropy.io.fits (third
party, not vendored in /root/reference; requirements.txt pins astropy<=7.2.0), which follows
the FITS standard 4.0: header = 80-character ASCII cards in 2880-byte blocks, data unit =
the array in C order as big-endian IEEE-754 doubles, zero-padded to a 2880-byte multiple.

Pinned by tests/test_fits_stage.py against the golden files the reference's own tests compare
(tests/data/outputs/*.fits, read with tests/fits_min.py; raw cards in
tests/golden/ref_cards.json).
"""
import numpy as np

BLOCK = 2880


def data_unit(arr) -> bytes:
    """Data unit of ``fits.ImageHDU(data=arr)`` / ``fits.PrimaryHDU(data=arr)`` for float64 data."""
    raw = np.ascontiguousarray(arr, dtype=np.float64).astype('>f8').tobytes()
    return raw + b'\0' * (-len(raw) % BLOCK)


def assemble(header_blocks, arrays) -> bytes:
    """The whole file: header blocks and data units alternate, HDU by HDU."""
    out = []
    for hb, arr in zip(header_blocks, arrays):
        assert len(hb) % BLOCK == 0
        out.append(hb)
        out.append(data_unit(arr))
    return b''.join(out)


def parse_card(card: str):
    """(keyword, value, comment) of one 80-character value card; commentary cards give
    (keyword, text, None).  Inverse of astropy's card formatting for the value types the
    reference writes (str, bool, int, float)."""
    if card.startswith('HIERARCH '):
        key, _, rest = card[9:].partition('=')
        key = key.strip()
    elif card[8:10] == '= ':
        key, rest = card[:8].strip(), card[10:]
    else:
        return card[:8].strip(), card[8:].rstrip(), None
    rest = rest.strip()
    if rest.startswith("'"):
        end = 1
        while True:  # closing quote = a quote not followed by another quote
            end = rest.index("'", end)
            if rest[end:end + 2] == "''":
                end += 2
                continue
            break
        value = rest[1:end].replace("''", "'").rstrip()
        tail = rest[end + 1:]
    else:
        text, sep, tail = rest.partition('/')
        tail = sep + tail
        text = text.strip()
        if text in ('T', 'F'):
            value = text == 'T'
        else:
            try:
                value = int(text)
            except ValueError:
                value = float(text)
    tail = tail.strip()
    comment = tail[1:].strip() if tail.startswith('/') else None
    # astropy writes ' / ' + comment; keep the exact (possibly truncated) comment text
    if comment is not None:
        idx = card.index(' / ' + comment[:1]) if comment else card.index(' /')
        comment = card[idx + 3:].rstrip()
    return key, value, comment
