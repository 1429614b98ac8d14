"""Philox4x32-10 and the reference's sampling rule (oracle; see oracle/__init__.py).

The reference draws one uniform per qubit from Julia's task-local Xoshiro stream (error_model.jl:69-71, 97-117),
which cannot be reproduced outside Julia.  The rebuild fixes a counter-based generator instead
(SURVEY 8d): Philox4x32-10 (Salmon et al., SC'11), key = (seed_lo, seed_hi), counter = (shot_lo, shot_hi, site, 0),
u = ((x0 << 32 | x1) >> 11) * 2^-53 in [0, 1).  What IS restated from the reference is the threshold rule:
depolarizing, Y tested first:  u < py -> Y ; u < px+py -> X ; u < px+py+pz -> Z ; else I  (error_model.jl:101-115);
flip model: u < p[i] (error_model.jl:69-71).
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & MASK for c in (c0, c1, c2, c3))
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)) & MASK, lo1, (hi0 ^ c3 ^ np.uint64(k1)) & MASK, lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def uniforms(seed, shot0, B, nsites):
    """(B, nsites) float64 uniforms for shots shot0..shot0+B-1."""
    shots = (np.arange(B, dtype=np.uint64) + np.uint64(shot0))[:, None]
    sites = np.arange(nsites, dtype=np.uint64)[None, :]
    sl = np.broadcast_to(shots & MASK, (B, nsites))
    sh = np.broadcast_to(shots >> np.uint64(32), (B, nsites))
    st = np.broadcast_to(sites, (B, nsites))
    x0, x1, _, _ = philox4x32_10(sl, sh, st, np.zeros_like(st), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    bits = ((x0 << np.uint64(32)) | x1) >> np.uint64(11)
    return bits.astype(np.float64) * 2.0 ** -53


def sample_depolarizing(px, py, pz, seed, shot0, B):
    """error_model.jl:97-117 on Philox uniforms -> (ex, ez) each (B, n) uint8."""
    px, py, pz = (np.asarray(p, dtype=np.float64) for p in (px, py, pz))
    u = uniforms(seed, shot0, B, len(px))
    isY = u < py
    isX = ~isY & (u < px + py)
    isZ = ~isY & ~isX & (u < px + py + pz)
    return (isX | isY).astype(np.uint8), (isZ | isY).astype(np.uint8)


def sample_flips(p, seed, shot0, B):
    """error_model.jl:69-71 on Philox uniforms -> (B, n) uint8."""
    p = np.asarray(p, dtype=np.float64)
    return (uniforms(seed, shot0, B, len(p)) < p).astype(np.uint8)
