"""Minimal stabilizer-tableau simulator (Aaronson-Gottesman CHP) used to check that the detectors and observables of a
generated circuit are deterministic in the absence of noise (oracle; see oracle/__init__.py).  Gates: H, S, CX, CZ,
Z-basis measurement and reset, X-basis variants via H.  Noise instructions are ignored."""
import numpy as np


class Tableau:
    def __init__(self, n):
        self.n = n
        self.x = np.zeros((2 * n, n), dtype=np.uint8)
        self.z = np.zeros((2 * n, n), dtype=np.uint8)
        self.r = np.zeros(2 * n, dtype=np.uint8)
        for i in range(n):
            self.x[i, i] = 1            # destabilizers
            self.z[n + i, i] = 1        # stabilizers

    def h(self, a):
        self.r ^= self.x[:, a] & self.z[:, a]
        self.x[:, a], self.z[:, a] = self.z[:, a].copy(), self.x[:, a].copy()

    def s(self, a):
        self.r ^= self.x[:, a] & self.z[:, a]
        self.z[:, a] ^= self.x[:, a]

    def cx(self, a, b):
        self.r ^= self.x[:, a] & self.z[:, b] & (self.x[:, b] ^ self.z[:, a] ^ 1)
        self.x[:, b] ^= self.x[:, a]
        self.z[:, a] ^= self.z[:, b]

    def _rowsum(self, h, i):
        def g(x1, z1, x2, z2):
            return np.where((x1 == 0) & (z1 == 0), 0,
                            np.where((x1 == 1) & (z1 == 1), z2.astype(int) - x2.astype(int),
                                     np.where((x1 == 1) & (z1 == 0), z2.astype(int) * (2 * x2.astype(int) - 1),
                                              x2.astype(int) * (1 - 2 * z2.astype(int)))))
        tot = 2 * int(self.r[h]) + 2 * int(self.r[i]) + int(g(self.x[i], self.z[i], self.x[h], self.z[h]).sum())
        self.r[h] = (tot % 4) // 2
        self.x[h] ^= self.x[i]
        self.z[h] ^= self.z[i]

    def measure(self, a, rng):
        """-> (outcome, deterministic)"""
        n = self.n
        ps = [p for p in range(n, 2 * n) if self.x[p, a]]
        if ps:
            p = ps[0]
            for i in range(2 * n):
                if i != p and self.x[i, a]:
                    self._rowsum(i, p)
            self.x[p - n], self.z[p - n], self.r[p - n] = self.x[p].copy(), self.z[p].copy(), self.r[p]
            self.x[p] = 0
            self.z[p] = 0
            self.z[p, a] = 1
            self.r[p] = rng.integers(0, 2)
            return int(self.r[p]), False
        # deterministic: accumulate in a scratch row
        sx = np.zeros(n, dtype=np.uint8)
        sz = np.zeros(n, dtype=np.uint8)
        sr = 0
        self.x = np.vstack([self.x, sx])
        self.z = np.vstack([self.z, sz])
        self.r = np.append(self.r, sr)
        for i in range(n):
            if self.x[i, a]:
                self._rowsum(2 * n, i + n)
        out = int(self.r[2 * n])
        self.x, self.z, self.r = self.x[:2 * n], self.z[:2 * n], self.r[:2 * n]
        return out, True


def run_noiseless(circ, seed=0):
    """circ: tensorqec.jl_b200.circuit.StimCircuit -> (detector values, observable values) of one noiseless run."""
    rng = np.random.default_rng(seed)
    T = Tableau(circ.n_qubits)
    rec = []
    det, obs = [], {}
    for ins in circ.instructions:
        nm, t = ins.name, ins.targets
        if nm == "H":
            for q in t:
                T.h(q)
        elif nm in ("S", "SQRT_Z"):
            for q in t:
                T.s(q)
        elif nm in ("CX", "CNOT", "ZCX"):
            for i in range(0, len(t), 2):
                T.cx(t[i], t[i + 1])
        elif nm in ("CZ", "ZCZ"):
            for i in range(0, len(t), 2):
                T.h(t[i + 1]); T.cx(t[i], t[i + 1]); T.h(t[i + 1])
        elif nm in ("M", "MZ", "MR", "MRZ", "MX", "MRX"):
            q = t[0]
            xb = nm in ("MX", "MRX")
            if xb:
                T.h(q)
            v, _ = T.measure(q, rng)
            rec.append(v)
            if nm.startswith("MR") and v:
                T.h(q); T.s(q); T.s(q); T.h(q)           # X = H Z H flips the qubit back to |0>
            if xb:
                T.h(q)
        elif nm in ("R", "RZ", "RX"):
            for q in t:
                v, _ = T.measure(q, rng)
                if v:
                    T.h(q); T.s(q); T.s(q); T.h(q)
                if nm == "RX":
                    T.h(q)
        elif nm == "DETECTOR":
            det.append(sum(rec[k] for k in t) % 2)
        elif nm == "OBSERVABLE_INCLUDE":
            o = int(ins.args[0])
            obs[o] = (obs.get(o, 0) + sum(rec[k] for k in t)) % 2
        elif nm in ("X_ERROR", "Y_ERROR", "Z_ERROR", "DEPOLARIZE1", "DEPOLARIZE2", "I", "TICK"):
            pass
        else:
            raise ValueError(f"chp: unsupported instruction {nm}")
    return det, obs
