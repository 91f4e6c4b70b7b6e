"""
spiceypy adaptor for the once-per-frame host constants (the reference's own SPICE
setup: planetmapper/base.py:554-586 kernel loading, :815-837 ``str2et``/``spkezr``;
planetmapper/body.py:522-535 ``bodvar``).

Exposes the five primitives :func:`planetmapper_b200.frame.build_body_constants`
needs, so the derived quantities (light time, sub-point, ring plane, LST Sun
longitude) come from the same tested code path as with MiniSpice.  spiceypy is not
installed in the authoring container or on the GPU box, so this module is exercised
only where a maintainer has it; nothing per-pixel happens here.
"""
from __future__ import annotations

import glob
import os
from pathlib import Path

import numpy as np


class SpiceProvider:
    name = 'spiceypy'

    def __init__(self, kernel_path: str | None = None, load_kernels: bool = True) -> None:
        import spiceypy as spice

        self.spice = spice
        if load_kernels and kernel_path:
            pattern = os.path.join(os.path.expanduser(kernel_path), '**', '*')
            kernels = {p for p in glob.glob(pattern, recursive=True) if os.path.isfile(p)}
            # reference load order: deepest first, then alphabetical (base.py:968-977)
            for k in sorted(kernels, key=lambda p: (-len(Path(p).resolve().parts),
                                                    os.path.dirname(p), os.path.basename(p),
                                                    os.path.normpath(p), p)):
                spice.furnsh(k)

    def clight(self) -> float:
        return float(self.spice.clight())

    def bods2c(self, name) -> int:
        if isinstance(name, (int, np.integer)):
            return int(name)
        return int(self.spice.bods2c(str(name).strip().upper()))

    def bodc2n(self, code: int) -> str:
        return str(self.spice.bodc2n(int(code)))

    def bodvar(self, body: int, item: str) -> np.ndarray:
        return np.array(self.spice.bodvrd(str(int(body)), item, 32)[1], dtype=float)

    def utc2et(self, utc: str) -> float:
        return float(self.spice.str2et(utc))

    def ssb_state(self, body: int, et: float) -> np.ndarray:
        return np.array(self.spice.spkssb(int(body), float(et), 'J2000'), dtype=float)

    def orientation(self, body: int, et: float):
        frame = 'IAU_' + self.bodc2n(body)
        xf = np.array(self.spice.sxform('J2000', frame, float(et)))
        rmat = xf[:3, :3]
        drmat = xf[3:, :3]
        om = -drmat @ rmat.T  # dR/dt = -[omega]x R
        omega = np.array([om[2, 1], om[0, 2], om[1, 0]])
        return rmat, omega
