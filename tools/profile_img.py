#!/usr/bin/env python
"""ncu driver for the single-frame image kernel (C2 shape): a few launches of the host-frame
(kernel-parameter) path.  PM_B200_LIBRARY selects a tuning variant."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from planetmapper_b200 import _lib as L  # noqa: E402

bc = bench.load_bc()
fr = bench.c2_frame(bc)
mask = L.mask_from_names(bench.C2_NAMES)
out = torch.empty((12, bench.SZ, bench.SZ), dtype=torch.float64, device='cuda')
for _ in range(6):
    L.backplanes_img_host(fr, bench.SZ, bench.SZ, mask, out=out)
torch.cuda.synchronize()
print('done')
