"""PMFrame layout, C-ABI surface and 'fails loudly' behaviour (no GPU needed)."""
import ctypes
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from planetmapper_b200 import _lib as L
from planetmapper_b200 import frame as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'pm_b200.h')


def test_pmframe_python_layout_matches_c_struct():
    """Compile a tiny C program that prints offsetof() for every PMFrame field."""
    fields = [n for n, _ in F.PMFRAME_FIELDS]
    prog = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', 'int main(void){']
    for n in fields:
        prog.append(f'printf("{n} %zu\\n", offsetof(PMFrame, {n}) / sizeof(double));')
    prog.append('printf("TOTAL %zu %d\\n", sizeof(PMFrame) / sizeof(double), PM_FRAME_NDOUBLES);')
    prog.append('return 0;}')
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, 'layout.c')
        exe = os.path.join(td, 'layout')
        open(src, 'w').write('\n'.join(prog))
        cc = '/usr/bin/gcc' if os.path.exists('/usr/bin/gcc') else 'gcc'
        subprocess.run([cc, src, '-o', exe], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split('\n')
    got = dict(line.split()[:2] for line in out if line and not line.startswith('TOTAL'))
    for n in fields:
        assert int(got[n]) == F.PMFRAME_OFFSETS[n][0], n
    total = [line for line in out if line.startswith('TOTAL')][0].split()
    assert int(total[1]) == int(total[2]) == F.PMFRAME_NDOUBLES


def test_library_loads_and_exports_every_declared_symbol():
    text = open(HEADER).read()
    declared = set(re.findall(r'\b(pm_[a-z0-9_]+)\s*\(', text))
    assert declared == set(L.EXPORTED_SYMBOLS)
    lib = L.load_library()
    for sym in sorted(declared):
        assert hasattr(lib, sym), f'{sym} not exported by libpm_b200.so'
    assert lib.pm_abi_version() == 10
    assert lib.pm_error_string(-1) == b'bad argument'
    # argument validation happens before any device work
    assert lib.pm_backplanes_img(None, 1, 4, 4, 1, None, None) == -1
    assert lib.pm_gather(None, None, None, 1, 4, 4, 0, 1, None, None, 4, 0, 0, 0, None, None) == -1
    assert lib.pm_spline_work_bytes(3, 8, 8, 3) > 0
    # empty inputs are valid and carry null pointers: accepted before any device work
    import ctypes

    frame = (ctypes.c_double * F.PMFRAME_NDOUBLES)()
    assert lib.pm_gather(None, None, None, 1, 4, 4, 0, 1, None, None, 0, 0, 0, 0, None, None) == 0
    assert lib.pm_xy2lonlat(frame, None, None, 0, None, None, None, None) == 0
    assert lib.pm_lonlat2xy(frame, None, None, 0, 0, None, None, None) == 0
    assert lib.pm_backplanes_map(frame, None, None, 0, 1, None, None) == 0
    assert lib.pm_xy2lonlat(frame, None, None, 3, None, None, None, None) == -1


def test_oracle_library_exports():
    from oracle import oracle as O

    lib = O.lib()
    for sym in ('pmo_backplanes_img', 'pmo_backplanes_map', 'pmo_xy2lonlat', 'pmo_lonlat2xy',
                'pmo_proj_inverse', 'pmo_gather_nearest'):
        assert hasattr(lib, sym)


def test_product_never_imports_the_oracle():
    """The product path must not route through oracle/ (no CPU fallback)."""
    for top in ('planetmapper_b200', 'tools', 'include'):   # only tests/, smoke() and bench.py's CPU legs may use it
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for fn in files:
                if fn.endswith(('.py', '.cu', '.cuh', '.h', '.sh')):
                    text = open(os.path.join(dirpath, fn), encoding='utf-8').read()
                    assert 'import oracle' not in text and 'from oracle' not in text, fn
                    assert 'pm_oracle' not in text, fn


@pytest.mark.skipif(__import__('torch').cuda.is_available(), reason='CPU-only check')
def test_compute_fails_loudly_without_gpu(bc_hst):
    import planetmapper_b200 as pm

    body = pm.BodyXY(constants=bc_hst, nx=8, ny=8)
    with pytest.raises(L.PMLibraryError):
        body.get_backplane_img('EMISSION')
    with pytest.raises(L.PMLibraryError):
        body.xy2lonlat(np.arange(3.0), np.arange(3.0))


def test_pack_frame_alt_changes_only_shape_fields(bc_hst):
    a = F.pack_frame(bc_hst, nx=7, ny=10, x0=2.5, y0=3.1, r0=3.9, rotation_radians=0.3)
    b = F.pack_frame(bc_hst, nx=7, ny=10, x0=2.5, y0=3.1, r0=3.9, rotation_radians=0.3, alt=1234.5)
    changed = {n for n, (o, k) in F.PMFRAME_OFFSETS.items() if not np.array_equal(a[o:o + k], b[o:o + k])}
    # (r_cut2 depends on max(radii) / r_eq, which is 1 for Jupiter)
    assert {'radii', 're', 'f', 'r_eq'} <= changed <= {'radii', 're', 'f', 'r_cut2', 'r_eq'}
    assert np.allclose(F.frame_field(b, 'radii') - F.frame_field(a, 'radii'), 1234.5)
