"""TNMAP throughput of even-distance / rectangular rotated surface codes: k_sweep (fresh pins) vs the general kernels."""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tensorqec.jl_b200 as tq
from tensorqec.jl_b200 import _cabi

def time_map(plan, words, ncw, reps=3):
    B = words.shape[0]
    d_syn = torch.from_numpy(words.view(np.int64)).cuda()
    d_cor = torch.empty((B, ncw), dtype=torch.int64, device="cuda")
    d_lp = torch.empty((B,), dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream()
    for _ in range(2):
        plan.decode_map_dev(d_syn.data_ptr(), B, d_cor.data_ptr(), d_lp.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps):
        plan.decode_map_dev(d_syn.data_ptr(), B, d_cor.data_ptr(), d_lp.data_ptr(), st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, d_cor.cpu().numpy(), d_lp.cpu().numpy()

for (dx, dz), B in (((6, 6), 4_000_000), ((8, 8), 2_000_000), ((5, 7), 4_000_000), ((9, 9), 2_000_000)):
    t = tq.CSSTannerGraph(tq.SurfaceCode(dx, dz))
    em = tq.iid_error(0.05, t)
    n = dx * dz
    ref = None
    for path in ("sweep", "general"):
        if path == "general":
            os.environ["TQEC_NO_SWEEP"] = "1"
        ct = tq.compile(tq.TNMAP(), t, em)
        os.environ.pop("TQEC_NO_SWEEP", None)
        words = _cabi.sample_errors(_cabi.MODEL_DEPOL, [em.px, em.py, em.pz], 3, 0, B, 0)
        H = np.zeros((t.stgx.ns + t.stgz.ns, 2 * n), dtype=np.uint8)
        H[:t.stgx.ns, n:] = t.stgx.H
        H[t.stgx.ns:, :n] = t.stgz.H
        syn = _cabi.GF2Matrix(H).apply(words)
        ms, cor, lp = time_map(ct.cd.plan, syn, max(1, (2 * n + 63) // 64))
        same = None if ref is None else bool(np.array_equal(lp, ref))
        ref = lp if ref is None else ref
        print(json.dumps({"case": f"TNMAP SurfaceCode({dx},{dz})", "path": path, "sweep": ct.cd.plan.query(_cabi.Q_SWEEP), "shots": B,
                          "ms": ms, "syndromes_per_s": B / ms * 1e3, "same_logp_as_sweep": same}), flush=True)
