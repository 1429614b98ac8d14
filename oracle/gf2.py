"""GF(2) kernels of the hot path, restated with numpy (oracle; see oracle/__init__.py).

Reference: src/decoding/error_model.jl:131-146 (`syndrome_extraction`: s = H e; CSS: sx = Hx ez, sz = Hz ex),
:161-163 and :179-181 (`check_logical_error`), src/codes/mod2.jl:44-71 (`bitmul!`).
"""
import numpy as np


def syndrome_extraction(e, H):
    """error_model.jl:131-133.  e: (..., nq) 0/1, H: (ns, nq) -> (..., ns)."""
    return (np.asarray(e, dtype=np.int64) @ np.asarray(H, dtype=np.int64).T & 1).astype(np.uint8)


def css_syndrome(ex, ez, Hx, Hz):
    """error_model.jl:144-146: X stabilizers see Z errors, Z stabilizers see X errors -> (sx, sz)."""
    return syndrome_extraction(ez, Hx), syndrome_extraction(ex, Hz)


def check_logical_error(e1, e2, L):
    """error_model.jl:161-163: any_i L[i,:].(e1 - e2)."""
    d = (np.asarray(e1, dtype=np.int64) ^ np.asarray(e2, dtype=np.int64))
    return ((d @ np.asarray(L, dtype=np.int64).T) & 1).any(axis=-1)


def check_logical_error_css(x1, z1, x2, z2, lx, lz):
    """error_model.jl:179-181: check(z1, z2, lx) || check(x1, x2, lz)."""
    return check_logical_error(z1, z2, lx) | check_logical_error(x1, x2, lz)


def bitmul(A, B):
    """mod2.jl:44-57 semantics by plain integer arithmetic (the packed form lives in the product's mod2.py)."""
    return ((np.asarray(A, dtype=np.int64) @ np.asarray(B, dtype=np.int64)) & 1).astype(np.uint8)
