"""Where the time of the user-level call goes: tq.decode(compiled, CSSSyndrome(sx, sz)) on 2e6 shots, d = 9."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import tensorqec.jl_b200 as tq
from tensorqec.jl_b200 import _cabi
from tensorqec.jl_b200.mod2 import as_bits

t = tq.CSSTannerGraph(tq.SurfaceCode(9, 9))
em = tq.iid_error(0.05, t)
ct = tq.compile(tq.TNMAP(), t, em)
B = 2_000_000
ep = tq.random_error_pattern(em, seed=3, shots=B)
syn = tq.syndrome_extraction(ep, t)
sx, sz = np.ascontiguousarray(syn.sx), np.ascontiguousarray(syn.sz)
def T(f, n=3):
    best = 1e9
    for _ in range(n):
        t0 = time.perf_counter(); r = f(); best = min(best, time.perf_counter() - t0)
    return best * 1e3, r
ms, _ = T(lambda: tq.decode(ct, tq.CSSSyndrome(sx, sz))); print("tq.decode total ms", round(ms, 1))
ms, _ = T(lambda: (as_bits(sx), as_bits(sz))); print("as_bits x2 ms", round(ms, 1))
ms, both = T(lambda: np.concatenate([sx, sz], axis=-1)); print("concatenate ms", round(ms, 1))
plan = ct.cd.plan
ms, _ = T(lambda: plan.decode_map_bits(both, 162)); print("decode_map_bits (C call + output alloc) ms", round(ms, 1))
words = tq.pack_bits(both)
ms, _ = T(lambda: plan.decode_map(words)); print("decode_map packed words (pageable) ms", round(ms, 1))
ms, _ = T(lambda: tq.CSSSyndrome(sx, sz)); print("CSSSyndrome() ms", round(ms, 1))
