#!/usr/bin/env python
"""
Executed FP64 work of backplanes_img_kernel per pixel class, measured with ncu counters
(run on the GPU box):

    python tools/count_flops_ncu.py            -> profiles/flops_per_pixel.json

The script profiles three synthetic 1024 x 1024 frames of the C2 body / plane stack with
    smsp__sass_thread_inst_executed_op_{dfma,dmul,dadd}_pred_on.sum
and divides by the pixel count:
    on_disc          r0 so large that every pixel hits the body,
    in_circle_miss   body moved out of the frame, early-out circle disabled (every pixel
                     runs the first intercept pass and misses),
    outside_circle   body out of the frame, early-out circle enabled.
flop = 2 DFMA + DMUL + DADD (comparisons and MUFU seeds count 0).  These are the FP64
operations THIS kernel executes (its own, leanest-known formulation of the path), which
is what bench.py multiplies by the per-class pixel counts of the timed frame.
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SZ = 1024
METRICS = ['smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum',
           'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum', 'smsp__inst_executed.sum',
           'sm__inst_executed_pipe_fp64.sum']


def child(kind):
    import torch

    import bench
    from planetmapper_b200 import _lib as L
    from planetmapper_b200 import frame as F

    bc = bench.load_bc()
    c = (SZ - 1) / 2
    if kind == 'on_disc':
        fr = F.pack_frame(bc, nx=SZ, ny=SZ, x0=c, y0=c, r0=3.0 * SZ, rotation_radians=0.0)
    elif kind == 'in_circle_miss':
        fr = F.pack_frame(bc, nx=SZ, ny=SZ, x0=c + 40.0 * SZ, y0=c, r0=0.45 * SZ, rotation_radians=0.0,
                          optimize_speed=False)
    else:
        fr = F.pack_frame(bc, nx=SZ, ny=SZ, x0=c + 40.0 * SZ, y0=c, r0=0.45 * SZ, rotation_radians=0.0)
    mask = L.mask_from_names(bench.C2_NAMES)
    out = L.backplanes_img_host(fr, SZ, SZ, mask)   # the single-frame launch the bench times
    torch.cuda.synchronize()
    frac = float(torch.isfinite(out[0]).double().mean())
    print('ON_DISC_FRACTION', kind, frac)


def main():
    if len(sys.argv) > 2 and sys.argv[1] == '--child':
        return child(sys.argv[2])
    res = {}
    for kind in ('on_disc', 'in_circle_miss', 'outside_circle'):
        cmd = ['ncu', '--metrics', ','.join(METRICS), '--clock-control', 'none', '-k', 'regex:backplanes_img', '--csv',
               sys.executable, os.path.abspath(__file__), '--child', kind]
        p = subprocess.run(cmd, capture_output=True, text=True)
        text = p.stdout
        frac = [l for l in text.split('\n') if l.startswith('ON_DISC_FRACTION')]
        start = text.index('"ID"')
        rows = list(csv.DictReader(io.StringIO(text[start:])))
        vals = {r['Metric Name']: float(r['Metric Value'].replace(',', '')) for r in rows}
        npx = SZ * SZ
        dfma, dmul, dadd = (vals[METRICS[i]] / npx for i in range(3))
        res[kind] = {'dfma': dfma, 'dmul': dmul, 'dadd': dadd, 'flop': 2 * dfma + dmul + dadd,
                     'warp_inst': vals[METRICS[3]] * 32 / npx / 32, 'fp64_pipe_inst': vals[METRICS[4]] / npx * 32 / 32,
                     'on_disc_fraction': float(frac[0].split()[-1]) if frac else None}
        print(kind, res[kind])
    out = {
        'source': 'tools/count_flops_ncu.py: ncu thread-instruction counters of backplanes_img_kernel on three '
                  'synthetic 1024x1024 frames (C2 body, 12-plane stack); per-pixel means',
        'flop_definition': '2*DFMA + DMUL + DADD executed by this kernel (predicated-on thread instructions)',
        'c2_12plane': {'on_disc': res['on_disc']['flop'], 'in_circle_miss': res['in_circle_miss']['flop'],
                       'outside_circle': res['outside_circle']['flop']},
        'detail': res,
        'reference_algorithm_flops': 'profiles/flops_per_pixel_oracle.json (libm-weighted count of the CSPICE-shaped '
                                     'oracle: 5315 / 501 / 298 per pixel); kept for context, not used for the roofline',
    }
    with open(os.path.join(ROOT, 'profiles', 'flops_per_pixel.json'), 'w') as f:
        json.dump(out, f, indent=1)
    print('wrote profiles/flops_per_pixel.json')


if __name__ == '__main__':
    main()
