#!/usr/bin/env python
"""Runs tests/analysis/count_flops.cpp (the CPU oracle with a counting scalar type) on the C2
frame and writes profiles/flops_per_pixel_oracle.json: the libm-weighted operation count of the
CSPICE-shaped oracle, kept for context next to the kernel's own executed-flop table (SURVEY.md 8(d):
"re-derive by instrumenting the CPU oracle").  Test-side analysis: it compiles oracle/ and therefore
lives under tests/, not tools/."""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from planetmapper_b200 import _lib as L  # noqa: E402


def run(exe, frame, sz, mask):
    with tempfile.NamedTemporaryFile(suffix='.bin', delete=False) as f:
        f.write(frame.tobytes())
        path = f.name
    out = subprocess.run([exe, path, str(sz), str(sz), str(mask)], check=True, capture_output=True, text=True)
    os.unlink(path)
    return json.loads(out.stdout)


def main():
    exe = os.path.join(tempfile.gettempdir(), 'pm_count_flops')
    cxx = '/usr/bin/g++' if os.path.exists('/usr/bin/g++') else 'g++'
    subprocess.run([cxx, '-O1', '-w', '-D_Thread_local=thread_local', '-o', exe, os.path.join(ROOT, 'tests', 'analysis', 'count_flops.cpp')], check=True)
    bc = bench.load_bc()
    sz = 512
    fr = bench.c2_frame(bc, sz)
    table = {
        'weights': {'add/sub/mul': 1, 'div': 18, 'sqrt': 15, 'sin': 45, 'cos': 45, 'atan2': 80, 'asin/acos': 70,
                    'fmod': 20, 'hypot': 18, 'floor/fmax/fmin/rint': 1, 'compare/fabs/neg': 0},
        'source': 'tests/analysis/count_flops.cpp: oracle/pm_oracle.c compiled with a counting scalar type, '
                  f'C2 frame geometry at {sz}x{sz} (per-class means are size independent)',
        'raw_columns': ['add/sub/mul', 'div', 'sqrt', 'sin+cos', 'atan2', 'asin/acos', 'other'],
    }
    table['c2_12plane'] = run(exe, fr, sz, L.mask_from_names(bench.C2_NAMES))
    table['all_26plane'] = run(exe, fr, sz, L.ALL_PLANES)
    os.makedirs(os.path.join(ROOT, 'profiles'), exist_ok=True)
    with open(os.path.join(ROOT, 'profiles', 'flops_per_pixel_oracle.json'), 'w') as f:
        json.dump(table, f, indent=1)
    print(json.dumps(table, indent=1))


if __name__ == '__main__':
    main()
