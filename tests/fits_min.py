"""Minimal FITS reader (numpy only) for the golden files; astropy is not installed."""
import numpy as np


def read_fits(path):
    data = open(path, 'rb').read()
    pos = 0
    hdus = []
    while pos < len(data):
        hdr = {}
        cards = []
        done = False
        while not done:
            block = data[pos:pos + 2880]
            pos += 2880
            for i in range(36):
                c = block[i * 80:(i + 1) * 80].decode('ascii')
                if c.startswith('END '):
                    done = True
                    break
                cards.append(c)
        for c in cards:
            if c.startswith('HIERARCH'):
                k, _, v = c[9:].partition('=')
            elif c[8:10] == '= ':
                k, v = c[:8], c[10:]
            else:
                continue
            v = v.strip()
            if v.startswith("'"):
                v = v[1:v.index("'", 1)].rstrip()
            else:
                v = v.split('/')[0].strip()
                try:
                    v = int(v)
                except ValueError:
                    try:
                        v = float(v)
                    except ValueError:
                        v = {'T': True, 'F': False}.get(v, v)
            hdr[k.strip()] = v
        bitpix = int(hdr.get('BITPIX', 8))
        naxis = int(hdr.get('NAXIS', 0))
        shape = [int(hdr['NAXIS%d' % (i + 1)]) for i in range(naxis)]
        n = abs(bitpix) // 8 * int(np.prod(shape)) if naxis else 0
        arr = None
        if n:
            dt = {-64: '>f8', -32: '>f4', 16: '>i2', 32: '>i4', 64: '>i8', 8: 'u1'}[bitpix]
            arr = np.frombuffer(data[pos:pos + n], dtype=dt).reshape(shape[::-1])
            arr = arr.astype(arr.dtype.newbyteorder('='))
        pos += (n + 2879) // 2880 * 2880
        hdus.append((hdr, arr))
    return hdus


def read_cards(path, max_hdus=None):
    """Raw 80-character header cards (END excluded) of every HDU, as a list of lists."""
    data = open(path, 'rb').read()
    pos = 0
    out = []
    while pos < len(data) and (max_hdus is None or len(out) < max_hdus):
        cards = []
        done = False
        while not done:
            block = data[pos:pos + 2880]
            pos += 2880
            for i in range(36):
                c = block[i * 80:(i + 1) * 80].decode('ascii')
                if c.startswith('END '):
                    done = True
                    break
                cards.append(c)
        kv = {c[:8].strip(): c[10:30].strip() for c in cards if c[8:10] == '= '}
        n = abs(int(kv['BITPIX'])) // 8
        for i in range(int(kv['NAXIS'])):
            n *= int(kv['NAXIS%d' % (i + 1)])
        if int(kv['NAXIS']) == 0:
            n = 0
        pos += (n + 2879) // 2880 * 2880
        out.append(cards)
    return out
