"""ctypes front end of oracle/csrc/oracle.c (oracle; see oracle/__init__.py).

`DensePlan`     : lowers a reference network (networks.py) + greedy tree (dense.py) to the index tables of the C dense
                  executor -- the timed stand-in for the reference's per-shot OMEinsum / TensorInference contraction.
`FrontierPlan`  : wraps a lowered frontier schedule for the C port of the recurrence.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import dense

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "csrc", "liboracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            subprocess.run(["make", "-C", os.path.join(_HERE, "csrc")], check=True)
        _lib = C.CDLL(_SO)
        _lib.oracle_dense_run.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int]
        _lib.oracle_frontier_run.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int]
        _lib.oracle_frontier_wide_run.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int]
    return _lib


def max_threads():
    return lib().oracle_max_threads()


class _DenseC(C.Structure):
    _fields_ = [("n_leaves", C.c_int32), ("n_steps", C.c_int32), ("n_checks", C.c_int32), ("n_vars", C.c_int32),
                ("node_size", C.c_void_p), ("leaf_off0", C.c_void_p), ("leaf_off1", C.c_void_p), ("leaf_ev", C.c_void_p),
                ("leaf_data", C.c_void_p), ("st_a", C.c_void_p), ("st_b", C.c_void_p), ("st_no", C.c_void_p),
                ("st_nk", C.c_void_p), ("st_tab", C.c_void_p), ("tabs", C.c_void_p), ("var_leaf", C.c_void_p)]


class _FrontierC(C.Structure):
    _fields_ = [("semiring", C.c_int32), ("n_vars", C.c_int32), ("n_checks", C.c_int32), ("n_obs", C.c_int32),
                ("n_steps", C.c_int32), ("w_max", C.c_int32), ("hdr", C.c_void_p), ("ints", C.c_void_p),
                ("tables", C.c_void_p), ("obs_slot", C.c_void_p)]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _pack(bits):
    bits = np.atleast_2d(np.asarray(bits, dtype=np.uint8))
    B, n = bits.shape
    W = max(1, (n + 63) // 64)
    pad = np.zeros((B, W * 64), dtype=np.uint8)
    pad[:, :n] = bits
    return np.ascontiguousarray(np.packbits(pad, axis=1, bitorder="little")).view("<u8").reshape(B, W)


def _unpack(words, n):
    by = np.ascontiguousarray(words, dtype="<u8").view(np.uint8).reshape(words.shape[0], -1)
    return np.unpackbits(by, axis=1, bitorder="little")[:, :n].copy()


class DensePlan:
    """net: networks.Network built for the ALL-ZERO syndrome; n_checks syndrome bits select leaf variants."""

    def __init__(self, net, n_checks, n_vars, maxplus, tree=None):
        self.maxplus = bool(maxplus)
        self.n_checks, self.n_vars = n_checks, n_vars
        sizes = {l: 1 for l in net.evidence}
        self.tree = tree or dense.greedy_tree(net.ixs, net.iy, sizes)
        tree = self.tree
        ev = set(net.evidence)
        eff = [[l for l in ls if l not in ev] for ls in tree.labels]
        nl = tree.n_leaves
        data, off0, off1, levs = [], [], [], []
        pos = 0

        def dom(x):
            with np.errstate(divide="ignore"):
                return np.log(x) if self.maxplus else x

        for i, (ix, t) in enumerate(zip(net.ixs, net.tensors)):
            t = np.asarray(t, dtype=np.float64)
            evl = [l for l in ix if l in ev]
            if len(evl) > 1:
                raise ValueError("a leaf with more than one evidence label is not supported")
            if evl:
                ax = list(ix).index(evl[0])
                v0 = np.take(t, 0, axis=ax).reshape(-1, order="F")
                v1 = np.take(t, 1, axis=ax).reshape(-1, order="F")
                bit = net.ev_bit[evl[0]]
            elif i in net.syn_leaf:
                v0, v1, bit = np.array([1.0, 0.0]), np.array([0.0, 1.0]), net.syn_leaf[i]
            else:
                v0 = v1 = t.reshape(-1, order="F")
                bit = -1
            off0.append(pos)
            data.append(dom(v0))
            pos += v0.size
            if bit >= 0:
                off1.append(pos)
                data.append(dom(v1))
                pos += v1.size
            else:
                off1.append(off0[-1])
            levs.append(bit)
        tabs, st_tab, st_no, st_nk, st_a, st_b = [], [], [], [], [], []
        tpos = 0
        for s, (a, b, out, con) in enumerate(tree.steps):
            oe = [l for l in out if l not in ev]
            ce = [l for l in con if l not in ev]
            no, nk = 1 << len(oe), 1 << len(ce)

            def table(node_labels, idx_labels, n):
                idx = np.arange(n, dtype=np.int64)
                o = np.zeros(n, dtype=np.int64)
                for p_, l in enumerate(idx_labels):
                    if l in node_labels:
                        o |= ((idx >> p_) & 1) << node_labels.index(l)
                return o.astype(np.int32)

            parts = [table(eff[a], oe, no), table(eff[b], oe, no), table(eff[a], ce, nk), table(eff[b], ce, nk)]
            st_tab.append(tpos)
            for prt in parts:
                tabs.append(prt)
                tpos += prt.size
            st_no.append(no); st_nk.append(nk); st_a.append(a); st_b.append(b)
        self._keep = dict(
            node_size=np.array([1 << len(e) for e in eff], dtype=np.int64),
            leaf_off0=np.array(off0, dtype=np.int64), leaf_off1=np.array(off1, dtype=np.int64),
            leaf_ev=np.array(levs, dtype=np.int32), leaf_data=np.concatenate(data).astype(np.float64),
            st_a=np.array(st_a, dtype=np.int32), st_b=np.array(st_b, dtype=np.int32),
            st_no=np.array(st_no, dtype=np.int64), st_nk=np.array(st_nk, dtype=np.int64),
            st_tab=np.array(st_tab, dtype=np.int64), tabs=np.concatenate(tabs).astype(np.int32),
            var_leaf=np.arange(max(n_vars, 1), dtype=np.int32))
        k = self._keep
        self.root_size = int(k["node_size"][-1])
        # output axis order of the root (sum-product): element index bits follow eff[root]; remember the permutation
        self.root_labels = eff[-1]
        self.iy = list(net.iy)
        self.c = _DenseC(nl, len(tree.steps), n_checks, n_vars, _p(k["node_size"]), _p(k["leaf_off0"]), _p(k["leaf_off1"]),
                         _p(k["leaf_ev"]), _p(k["leaf_data"]), _p(k["st_a"]), _p(k["st_b"]), _p(k["st_no"]), _p(k["st_nk"]),
                         _p(k["st_tab"]), _p(k["tabs"]), _p(k["var_leaf"]))
        self.ops_per_shot = float(sum(n * m for n, m in zip(st_no, st_nk)))

    def run(self, syndromes, threads=0, want_config=True):
        words = _pack(syndromes)
        B = words.shape[0]
        if self.maxplus:
            cw = max(1, (self.n_vars + 63) // 64)
            cfg = np.zeros((B, cw), dtype=np.uint64) if want_config else None
            lp = np.zeros(B)
            rc = lib().oracle_dense_run(C.byref(self.c), 1, _p(words), B, _p(cfg) if want_config else None, _p(lp), threads)
            assert rc == 0
            return lp, (_unpack(cfg, self.n_vars) if want_config else None)
        out = np.zeros((B, self.root_size))
        rc = lib().oracle_dense_run(C.byref(self.c), 0, _p(words), B, None, _p(out), threads)
        assert rc == 0
        # reorder root element bits (root_labels order) into iy order, first iy label fastest
        idx = np.arange(self.root_size)
        src = np.zeros_like(idx)
        for i, l in enumerate(self.iy):
            src |= ((idx >> i) & 1) << self.root_labels.index(l)
        return out[:, src]


class FrontierPlan:
    def __init__(self, sch):
        self.sch = sch
        self._keep = dict(hdr=np.ascontiguousarray(sch.hdr, dtype=np.int32), ints=np.ascontiguousarray(sch.ints, dtype=np.int32),
                          tables=np.ascontiguousarray(sch.tables, dtype=np.float64),
                          obs=np.ascontiguousarray(sch.obs_slot if sch.obs_slot else [0], dtype=np.int32))
        k = self._keep
        self.c = _FrontierC(sch.semiring, sch.n_vars, sch.n_checks, sch.n_obs, len(sch.steps), sch.w_max, _p(k["hdr"]),
                            _p(k["ints"]), _p(k["tables"]), _p(k["obs"]))

    def run_wide(self, syndromes, threads=0):
        """Sum-product plans of any width up to 31 bits lowered with the stable layout: one shot at a time, threads
        over the state entries (oracle_frontier_wide_run)."""
        words = _pack(syndromes)
        B = words.shape[0]
        out = np.zeros((B, 1 << self.sch.n_obs))
        rc = lib().oracle_frontier_wide_run(C.byref(self.c), _p(words), B, _p(out), threads)
        assert rc == 0, f"oracle_frontier_wide_run failed ({rc})"
        return np.ldexp(out, getattr(self.sch, "log2_scale", 0))

    def run(self, syndromes, threads=0, want_config=True):
        words = _pack(syndromes)
        B = words.shape[0]
        sch = self.sch
        if sch.semiring == 0:
            cw = max(1, (sch.n_vars + 63) // 64)
            cfg = np.zeros((B, cw), dtype=np.uint64) if want_config else None
            lp = np.zeros(B)
            rc = lib().oracle_frontier_run(C.byref(self.c), _p(words), B, _p(cfg) if want_config else None, _p(lp), threads)
            assert rc == 0
            return lp, (_unpack(cfg, sch.n_vars) if want_config else None)
        out = np.zeros((B, 1 << sch.n_obs))
        rc = lib().oracle_frontier_run(C.byref(self.c), _p(words), B, None, _p(out), threads)
        assert rc == 0
        return np.ldexp(out, getattr(sch, "log2_scale", 0))
