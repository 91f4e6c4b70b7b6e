#!/usr/bin/env python
"""Where the set-up time of the C4 end-to-end stream goes (cProfile of the path up to the first chunk)."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import planetmapper_b200 as pm  # noqa: E402

bc = bench.load_bc()
cube_h, lo, la = bench.c4_inputs(256)
for rep in range(2):
    obs = pm.Observation(data=cube_h, constants=bc)
    obs.set_disc_params(31.5, 31.5, 28.0, 0.0)
    prof = cProfile.Profile()
    t0 = time.perf_counter()
    prof.enable()
    it = obs.iter_mapped_data('linear', planes_per_chunk=32, degree_interval=0.1)
    first = next(it)
    prof.disable()
    print('rep', rep, 'first chunk after', time.perf_counter() - t0)
    for _ in it:
        pass
    del it, obs
    if rep == 1:
        pstats.Stats(prof).sort_stats('cumulative').print_stats(22)
