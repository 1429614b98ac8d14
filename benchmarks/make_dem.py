#!/usr/bin/env python
"""Synthetic detector error models in stim's text format for BASELINE configs[3] (no stim in this image).

Phenomenological memory-Z experiment on the d x d rotated surface code, `rounds` noisy syndrome-measurement rounds
followed by a perfect data read-out: detectors D[t*nz + s] compare Z-stabilizer s between rounds t-1 and t
(t = 0 .. rounds), one bit-flip mechanism per (data qubit, round) with probability p flips the detectors of the
adjacent Z stabilizers in layer t and, if the qubit lies on the logical Z support, observable L0; one measurement
mechanism per (stabilizer, round) with probability q flips D[t][s] and D[t+1][s].  This is the standard 3-D matching
graph of the surface code written as an (undecomposed, loop-free) DEM: what `parse_dem_string` accepts
(src/stim_parser/stim_parser.jl:342-375).  Circuit-level DEMs additionally contain hook / correlated mechanisms; those
need stim or the reference's circuit -> DEM generator.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def surface_memory_dem(d: int, rounds: int, p: float, q: float) -> str:
    import tensorqec.jl_b200 as tq
    t = tq.CSSTannerGraph(tq.SurfaceCode(d, d))
    _, lz = tq.logical_operator(t)
    zs = t.stgz.s2q
    nz = len(zs)
    q2s = t.stgz.q2s
    lines = []
    for r in range(rounds + 1):                       # data errors before round r (the last layer is the read-out)
        if r == rounds:
            break
        for qb in range(d * d):
            tg = [f"D{r * nz + s}" for s in q2s[qb]]
            if lz[0, qb]:
                tg.append("L0")
            lines.append(f"error({p!r}) " + " ".join(tg))
    for r in range(rounds):
        for s in range(nz):
            lines.append(f"error({q!r}) D{r * nz + s} D{(r + 1) * nz + s}")
    for r in range(rounds + 1):
        for s in range(nz):
            lines.append(f"detector({s}, 0, {r}) D{r * nz + s}")
    lines.append("logical_observable L0")
    return "\n".join(lines) + "\n"


if __name__ == "__main__":
    d = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    r = int(sys.argv[2]) if len(sys.argv) > 2 else d
    sys.stdout.write(surface_memory_dem(d, r, 0.01, 0.01))
