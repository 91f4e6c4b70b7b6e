#!/usr/bin/env python
"""Per-step times of the drop-in end-to-end call (variance probe)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import planetmapper_b200 as pm
from planetmapper_b200 import _lib as L
bc = bench.load_bc()
SZ = bench.SZ
def dropin():
    body = pm.BodyXY(constants=bc, nx=SZ, ny=SZ)
    planes = {n: body.get_backplane_img(n) for n in bench.C2_NAMES}
    return float(planes['EMISSION'][SZ // 2, SZ // 2])
ts = []
for i in range(40):
    t0 = time.perf_counter(); dropin(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
print('steps ms:', ' '.join(f'{t:.1f}' for t in ts))
print('host stats', {k: v for k, v in torch.cuda.host_memory_stats().items() if 'current' in k or 'num_host_alloc' in k or 'host_alloc_time' in k})
# phases of one step
body = pm.BodyXY(constants=bc, nx=SZ, ny=SZ)
for n in bench.C2_NAMES:
    t0 = time.perf_counter(); a = body.get_backplane_img(n); t1 = time.perf_counter()
    print(n, f'{(t1 - t0) * 1e3:.2f} ms')
