"""Stim-format Clifford circuits and their detector error models (the generator of the
circuit-level inputs of the TNMMAP decoder).

Reference: src/stim_parser/stim_parser.jl:10-335 (`parse_stim_file`, `parse_stim_string`: instruction set, REPEAT
flattening, `rec[-k]` look-back), src/decoding/dem.jl:22-145 (`detector_error_model`, `forward_analysis`,
`push_to_dict!`), golden answers in test/decoding/dem.jl:5-109.

The reference propagates every fault FORWARD through the rest of the circuit (one Pauli-frame walk per fault and Pauli
component, O(faults x gates)).  Here the same map fault -> flipped detectors is obtained by ONE BACKWARD sweep that
keeps, for every qubit, the set of detectors / observables an X or a Z error at the current time would flip (two
bitsets per qubit, conjugated through each gate): O(gates), exact, no per-fault work.  A fault site then just reads its
sets.  This is compile-time work on the host (once per circuit), not part of the per-shot path.

Numbering (0-based): detectors in order of appearance (after REPEAT flattening); observable L# becomes id
n_detectors + #, the convention of `parse_dem_string` (dem.py).  The reference numbers DETECTOR and
OBSERVABLE_INCLUDE lines with one shared counter and ignores the observable's argument; the two agree whenever the
observables come last and are listed once each, which holds for every reference test and stim-generated circuit.

Noise channels.  X_ERROR / Y_ERROR / Z_ERROR(p): one mechanism.  DEPOLARIZE1(p): the reference's rule (dem.jl:40-66) --
X, Y, Z components as three independent mechanisms of probability (1 - sqrt(1 - 4p/3)) / 2 each, or one mechanism of
probability 2p/3 when a component is invisible.  DEPOLARIZE2(p) (rejected by the reference's analysis, dem.jl:41) is
expanded the same way into its 15 two-qubit Pauli components with stim's independent-channel probability
1/2 - 1/2 (1 - 16p/15)^(1/8).  M(p): measurement flip.  Mechanisms with equal detector sets are merged with
p <- p1 (1 - p2) + p2 (1 - p1) (`push_to_dict!`, dem.jl:130-141); mechanisms that flip nothing are dropped.
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

from .dem import DetectorErrorModel

_GATES_1Q = {"I", "X", "Y", "Z", "H", "S", "S_DAG", "SQRT_Z", "SQRT_Z_DAG", "SQRT_X", "SQRT_X_DAG", "C_XYZ", "C_ZYX"}
_GATES_2Q = {"CX", "CNOT", "ZCX", "CZ", "ZCZ", "CY", "ZCY", "SWAP"}
_MEASURE = {"M": "Z", "MZ": "Z", "MX": "X", "MY": "Y", "MR": "Z", "MRZ": "Z", "MRX": "X", "MRY": "Y"}
_RESET = {"R", "RZ", "RX", "RY"}
_NOISE = {"X_ERROR", "Y_ERROR", "Z_ERROR", "DEPOLARIZE1", "DEPOLARIZE2"}
_ANNOT = {"TICK", "QUBIT_COORDS", "SHIFT_COORDS"}
_FEEDBACK = {"CX": "X", "CNOT": "X", "ZCX": "X", "CZ": "Z", "ZCZ": "Z", "CY": "Y", "ZCY": "Y"}


@dataclass
class Instruction:
    name: str
    args: Tuple[float, ...]
    targets: Tuple[int, ...]                    # qubits, or absolute measurement-record indices for DETECTOR / OBSERVABLE


@dataclass
class StimCircuit:
    n_qubits: int
    instructions: List[Instruction] = field(default_factory=list)     # REPEAT blocks flattened
    n_measurements: int = 0
    n_detectors: int = 0
    n_observables: int = 0

    def count(self, name: str) -> int:
        return sum(1 for i in self.instructions if i.name == name)


_LINE = re.compile(r"^([A-Za-z_][A-Za-z_0-9]*)\s*(?:\(([^)]*)\))?\s*(.*)$")


def _flatten(lines: List[str]) -> List[str]:
    """Expand `REPEAT n { ... }` blocks (nested blocks allowed)."""
    out: List[str] = []
    i = 0
    while i < len(lines):
        ln = lines[i]
        m = re.match(r"^REPEAT\s+(\d+)\s*\{\s*$", ln)
        if m:
            depth, j = 1, i + 1
            while j < len(lines) and depth:
                if re.match(r"^REPEAT\s+\d+\s*\{\s*$", lines[j]):
                    depth += 1
                elif lines[j] == "}":
                    depth -= 1
                j += 1
            if depth:
                raise ValueError("unterminated REPEAT block")
            body = _flatten(lines[i + 1:j - 1])
            out += body * int(m.group(1))
            i = j
        elif ln == "}":
            raise ValueError("unmatched '}'")
        else:
            out.append(ln)
            i += 1
    return out


def parse_stim_string(content: str, n_qubits: Optional[int] = None) -> StimCircuit:
    """stim_parser.jl:10-200.  `n_qubits` defaults to the largest qubit index + 1."""
    lines = []
    for raw in content.split("\n"):
        ln = raw.split("#", 1)[0].strip()
        if ln:
            lines.append(ln)
    circ = StimCircuit(0)
    max_q = -1
    for ln in _flatten(lines):
        m = _LINE.match(ln)
        if not m:
            raise ValueError(f"cannot parse stim line: {ln!r}")
        name = m.group(1).upper()
        args = tuple(float(x) for x in m.group(2).split(",")) if m.group(2) and m.group(2).strip() else ()
        toks = m.group(3).split()
        if name in _ANNOT:
            continue
        if name in ("DETECTOR", "OBSERVABLE_INCLUDE"):
            recs = []
            for t in toks:
                mm = re.match(r"^rec\[(-\d+)\]$", t)
                if not mm:
                    raise ValueError(f"{name}: unsupported target {t!r}")
                k = circ.n_measurements + int(mm.group(1))
                if k < 0:
                    raise ValueError(f"{name}: {t} looks back before the first measurement")
                recs.append(k)
            if name == "DETECTOR":
                circ.instructions.append(Instruction(name, (float(circ.n_detectors),), tuple(recs)))
                circ.n_detectors += 1
            else:
                obs = int(args[0]) if args else 0
                circ.instructions.append(Instruction(name, (float(obs),), tuple(recs)))
                circ.n_observables = max(circ.n_observables, obs + 1)
            continue
        if name not in _GATES_1Q | _GATES_2Q | set(_MEASURE) | _RESET | _NOISE:
            raise ValueError(f"Unknown instruction: {name}")
        if name in _FEEDBACK and any(t.startswith("rec[") for t in toks):
            # classically controlled Pauli (stim_parser.jl:236-263: `condition(measure, X | Z, nothing)`): pairs of
            # (rec[-k], qubit); stored with the ABSOLUTE record index as first target
            if len(toks) % 2:
                raise ValueError(f"{name} needs pairs of targets")
            for a, b in zip(toks[0::2], toks[1::2]):
                mm = re.match(r"^rec\[(-\d+)\]$", a)
                if not mm or not re.match(r"^\d+$", b):
                    raise ValueError(f"{name}: unsupported target pair {a!r} {b!r}")
                k = circ.n_measurements + int(mm.group(1))
                if k < 0:
                    raise ValueError(f"{name}: {a} looks back before the first measurement")
                circ.instructions.append(Instruction("FEEDBACK_" + _FEEDBACK[name], (), (k, int(b))))
                max_q = max(max_q, int(b))
            continue
        if any(not re.match(r"^!?\d+$", t) for t in toks):
            raise ValueError(f"{name}: unsupported target in {ln!r}")
        qs = tuple(int(t.lstrip("!")) for t in toks)
        if name in _GATES_2Q | {"DEPOLARIZE2"} and len(qs) % 2:
            raise ValueError(f"{name} needs pairs of qubits")
        if name in _NOISE and len(args) != 1:
            raise ValueError(f"{name} needs one probability")
        if qs:
            max_q = max(max_q, max(qs))
        if name in _MEASURE:
            for q in qs:                                     # one record per target, in order
                circ.instructions.append(Instruction(name, args, (q,)))
                circ.n_measurements += 1
        else:
            circ.instructions.append(Instruction(name, args, qs))
    circ.n_qubits = n_qubits if n_qubits is not None else max_q + 1
    if max_q >= circ.n_qubits:
        raise ValueError(f"qubit {max_q} used but the circuit has {circ.n_qubits} qubits")
    return circ


def parse_stim_file(path: str, n_qubits: Optional[int] = None) -> StimCircuit:
    with open(path, "r") as fh:
        return parse_stim_string(fh.read(), n_qubits)


# ------------------------------------------------------------------------------------------------------------------
def _merge(table: Dict[int, float], mask: int, p: float):
    """push_to_dict! (dem.jl:130-141)."""
    if mask == 0 or p == 0.0:
        return
    if mask in table:
        q = table[mask]
        table[mask] = (1.0 - q) * p + q * (1.0 - p)
    else:
        table[mask] = p


def detector_error_model(circ: StimCircuit) -> DetectorErrorModel:
    """dem.jl:22-92 by one backward sweep (module docstring).  Mechanisms are returned sorted by detector set."""
    nq = circ.n_qubits
    sx = [0] * nq                                   # detectors / observables flipped by an X error on qubit q from here on
    sz = [0] * nq
    n_det = circ.n_detectors
    rec: Dict[int, int] = {}                        # measurement record -> detectors / observables that include it
    table: Dict[int, float] = {}
    m_idx = circ.n_measurements
    for ins in reversed(circ.instructions):
        nm, t = ins.name, ins.targets
        if nm == "DETECTOR":
            bit = 1 << int(ins.args[0])
            for k in t:
                rec[k] = rec.get(k, 0) ^ bit
        elif nm == "OBSERVABLE_INCLUDE":
            bit = 1 << (n_det + int(ins.args[0]))
            for k in t:
                rec[k] = rec.get(k, 0) ^ bit
        elif nm.startswith("FEEDBACK_"):
            # a Pauli applied iff an earlier measurement read 1: a flipped record applies it wrongly, so the record
            # inherits what that Pauli would flip from here on
            k, q = t
            pa = nm[-1]
            eff = (sx[q] if pa in "XY" else 0) ^ (sz[q] if pa in "ZY" else 0)
            rec[k] = rec.get(k, 0) ^ eff
        elif nm in _MEASURE:
            m_idx -= 1
            q, r = t[0], rec.get(m_idx, 0)
            basis = _MEASURE[nm]
            if nm.startswith("MR"):                  # measure, then reset: nothing before it survives the reset
                sx[q] = sz[q] = 0
            if basis in ("Z", "Y"):
                sx[q] ^= r
            if basis in ("X", "Y"):
                sz[q] ^= r
            if ins.args and ins.args[0] > 0.0:
                _merge(table, r, float(ins.args[0]))
        elif nm in _RESET:
            for q in t:
                sx[q] = sz[q] = 0
        elif nm in ("H",):
            for q in t:
                sx[q], sz[q] = sz[q], sx[q]
        elif nm in ("S", "S_DAG", "SQRT_Z", "SQRT_Z_DAG"):
            for q in t:
                sx[q] ^= sz[q]
        elif nm in ("SQRT_X", "SQRT_X_DAG"):
            for q in t:
                sz[q] ^= sx[q]
        elif nm == "C_XYZ":                          # X -> Y -> Z -> X
            for q in t:
                sx[q], sz[q] = sx[q] ^ sz[q], sx[q]
        elif nm == "C_ZYX":                          # X -> Z -> Y -> X
            for q in t:
                sx[q], sz[q] = sz[q], sx[q] ^ sz[q]
        elif nm in ("CX", "CNOT", "ZCX"):
            for i in range(len(t) - 2, -1, -2):      # pairs act in order; undo them last to first
                c, x = t[i], t[i + 1]
                sx[c] ^= sx[x]
                sz[x] ^= sz[c]
        elif nm in ("CZ", "ZCZ"):
            for i in range(len(t) - 2, -1, -2):
                a, b = t[i], t[i + 1]
                sx[a], sx[b] = sx[a] ^ sz[b], sx[b] ^ sz[a]
        elif nm in ("CY", "ZCY"):
            for i in range(len(t) - 2, -1, -2):
                c, x = t[i], t[i + 1]
                zc = sz[c]
                sx[c] ^= sx[x] ^ sz[x]
                sz[x] ^= zc
                sx[x] ^= zc
        elif nm == "SWAP":
            for i in range(len(t) - 2, -1, -2):
                a, b = t[i], t[i + 1]
                sx[a], sx[b] = sx[b], sx[a]
                sz[a], sz[b] = sz[b], sz[a]
        elif nm in ("X_ERROR", "Y_ERROR", "Z_ERROR"):
            p = float(ins.args[0])
            for q in reversed(t):
                _merge(table, sx[q] if nm[0] == "X" else (sz[q] if nm[0] == "Z" else sx[q] ^ sz[q]), p)
        elif nm == "DEPOLARIZE1":
            p = float(ins.args[0])
            if p > 0.75:
                raise ValueError("Can't analyze single-qubit over-mixing depolarizing errors (probability > 3/4)")
            for q in reversed(t):
                x, z = sx[q], sz[q]
                y = x ^ z
                if x == 0 or y == 0 or z == 0:       # one component is invisible: the other two coincide (dem.jl:55-63)
                    _merge(table, x or y or z, p * 2.0 / 3.0)
                else:
                    pc = (1.0 - math.sqrt(1.0 - 4.0 * p / 3.0)) / 2.0
                    for msk in (x, y, z):
                        _merge(table, msk, pc)
        elif nm == "DEPOLARIZE2":
            p = float(ins.args[0])
            if p > 15.0 / 16.0:
                raise ValueError("Can't analyze two-qubit over-mixing depolarizing errors (probability > 15/16)")
            pc = 0.5 - 0.5 * (1.0 - 16.0 * p / 15.0) ** 0.125
            for i in range(len(t) - 2, -1, -2):
                a, b = t[i], t[i + 1]
                pa = (0, sx[a], sx[a] ^ sz[a], sz[a])
                pb = (0, sx[b], sx[b] ^ sz[b], sz[b])
                for ia in range(4):
                    for ib in range(4):
                        if ia or ib:
                            _merge(table, pa[ia] ^ pb[ib], pc)
        elif nm in _GATES_1Q:
            pass                                     # Paulis and identity only change signs
        else:                                        # pragma: no cover - parser admits nothing else
            raise ValueError(f"Unknown instruction: {nm}")
    n_obs = circ.n_observables
    rates, flipped = [], []
    for mask in sorted(table, key=lambda m: [b for b in range(n_det + n_obs) if (m >> b) & 1]):
        rates.append(table[mask])
        flipped.append([b for b in range(n_det + n_obs) if (mask >> b) & 1])
    return DetectorErrorModel(rates, flipped, list(range(n_det)), list(range(n_det, n_det + n_obs)))


def dem_to_string(dem: DetectorErrorModel) -> str:
    """stim DEM text (`error(p) D.. L..`), readable by `parse_dem_string`."""
    n_det = len(dem.detector_list)
    out = []
    for p, fl in zip(dem.error_rates, dem.flipped_detectors):
        toks = [f"D{d}" if d < n_det else f"L{d - n_det}" for d in fl]
        out.append(f"error({p!r}) " + " ".join(toks))
    return "\n".join(out) + "\n"


def circuit_to_string(circ: StimCircuit) -> str:
    """Stim text of a parsed circuit (REPEAT blocks stay flattened; `rec[-k]` look-backs are recomputed from the absolute
    record indices).  `parse_stim_string(circuit_to_string(c))` reproduces `c` instruction for instruction."""
    out: List[str] = []
    n_meas = 0
    for ins in circ.instructions:
        nm = ins.name
        if nm in ("DETECTOR", "OBSERVABLE_INCLUDE"):
            recs = " ".join(f"rec[{k - n_meas}]" for k in ins.targets)
            head = "DETECTOR" if nm == "DETECTOR" else f"OBSERVABLE_INCLUDE({int(ins.args[0])})"
            out.append((head + " " + recs).rstrip())
        elif nm.startswith("FEEDBACK_"):
            k, q = ins.targets
            out.append(f"C{nm[-1]} rec[{k - n_meas}] {q}")
        else:
            arg = "(" + ", ".join(repr(a) for a in ins.args) + ")" if ins.args else ""
            out.append((nm + arg + " " + " ".join(str(q) for q in ins.targets)).rstrip())
            if nm in _MEASURE:
                n_meas += len(ins.targets)
    return "\n".join(out) + "\n"


def dump_stim_file(circ: StimCircuit, filename: str) -> None:
    """stim_parser.jl:378-437 (`dump_stim_file`), for this package's circuit container: every instruction the parser
    reads is written back (the reference writes H / X / Y / Z / M / CX / DETECTOR / OBSERVABLE_INCLUDE only)."""
    with open(filename, "w") as fh:
        fh.write(circuit_to_string(circ))


# ------------------------------------------------------------------------------------------------------------------
def surface_memory_circuit(d: int, rounds: int, basis: str = "Z", after_clifford_depolarization: float = 0.0,
                           before_round_data_depolarization: float = 0.0, before_measure_flip_probability: float = 0.0,
                           after_reset_flip_probability: float = 0.0) -> str:
    """Stim text of a rotated-surface-code memory experiment built from `SurfaceCode(d, d)` (codes.py): the layout,
    noise knobs and detector structure of stim's `surface_code:rotated_memory_z/x` (there is no stim in this image).
    Data qubit (i, j) = i*d + j; one ancilla per stabilizer after the data qubits, in the code's generator order.
    Every round: reset-free repeated MR of the ancillas; X-type ancillas are conjugated by H; the four CX layers visit
    the corners of a plaquette in the order NW, NE, SW, SE for X-type and NW, SW, NE, SE for Z-type checks, so that
    every data qubit is touched once per layer and hook errors run perpendicular to the logical operators."""
    from .codes import SurfaceCode
    from .tanner import CSSTannerGraph
    if basis not in ("Z", "X"):
        raise ValueError("basis must be 'Z' or 'X'")
    t = CSSTannerGraph(SurfaceCode(d, d))
    n = d * d
    xs = [list(c) for c in t.stgx.s2q]
    zs = [list(c) for c in t.stgz.s2q]
    anc_x = [n + k for k in range(len(xs))]
    anc_z = [n + len(xs) + k for k in range(len(zs))]
    rc = lambda q: (q // d, q % d)

    def corner_slots(sup, kind):
        """time slot (0..3) of every data qubit of a check."""
        rows = sorted({rc(q)[0] for q in sup})
        cols = sorted({rc(q)[1] for q in sup})
        if len(sup) == 4:
            r0, c0 = rows[0], cols[0]
        elif len(rows) == 2:                       # vertical pair: X boundary, left (c = 0) or right (c = d-1) column
            r0 = rows[0]
            c0 = cols[0] - 1 if cols[0] == 0 else cols[0]
        else:                                      # horizontal pair: Z boundary, top (r = 0) or bottom (r = d-1) row
            c0 = cols[0]
            r0 = rows[0] - 1 if rows[0] == 0 else rows[0]
        order = {"X": [(0, 0), (0, 1), (1, 0), (1, 1)], "Z": [(0, 0), (1, 0), (0, 1), (1, 1)]}[kind]
        return {q: order.index((rc(q)[0] - r0, rc(q)[1] - c0)) for q in sup}

    layers: List[List[int]] = [[], [], [], []]
    for a, sup in zip(anc_x, xs):
        for q, s in corner_slots(sup, "X").items():
            layers[s] += [a, q]                    # X check: ancilla controls
    for a, sup in zip(anc_z, zs):
        for q, s in corner_slots(sup, "Z").items():
            layers[s] += [q, a]                    # Z check: data controls
    for ly in layers:
        used = ly
        if len(set(used)) != len(used):
            raise AssertionError("a qubit is used twice in one CX layer")
    data = list(range(n))
    anc = anc_x + anc_z
    p1, p2, pm, pr = after_clifford_depolarization, before_round_data_depolarization, before_measure_flip_probability, after_reset_flip_probability
    L: List[str] = []
    j = lambda qs: " ".join(str(q) for q in qs)
    for q in data:
        L.append(f"QUBIT_COORDS({2 * rc(q)[1] + 1}, {2 * rc(q)[0] + 1}) {q}")
    L.append(("R " if basis == "Z" else "RX ") + j(data))
    L.append("R " + j(anc))
    if pr > 0:
        L.append(("X_ERROR" if basis == "Z" else "Z_ERROR") + f"({pr}) " + j(data))
        L.append(f"X_ERROR({pr}) " + j(anc))

    def round_body(first: bool) -> List[str]:
        B = ["TICK"]
        if p2 > 0:
            B.append(f"DEPOLARIZE1({p2}) " + j(data))
        B.append("H " + j(anc_x))
        if p1 > 0:
            B.append(f"DEPOLARIZE1({p1}) " + j(anc_x))
        for ly in layers:
            B.append("TICK")
            B.append("CX " + j(ly))
            if p1 > 0:
                B.append(f"DEPOLARIZE2({p1}) " + j(ly))
        B.append("TICK")
        B.append("H " + j(anc_x))
        if p1 > 0:
            B.append(f"DEPOLARIZE1({p1}) " + j(anc_x))
        B.append("TICK")
        if pm > 0:
            B.append(f"X_ERROR({pm}) " + j(anc))
        B.append("MR " + j(anc))
        if pr > 0:
            B.append(f"X_ERROR({pr}) " + j(anc))
        na = len(anc)
        same = anc_z if basis == "Z" else anc_x     # checks that are deterministic in the first round
        for k, a in enumerate(anc):
            back = k - na
            if first:
                if a in same:
                    B.append(f"DETECTOR rec[{back}]")
            else:
                B.append(f"DETECTOR rec[{back}] rec[{back - na}]")
        return B

    L += round_body(True)
    if rounds > 1:
        L.append(f"REPEAT {rounds - 1} {{")
        L += ["    " + s for s in round_body(False)]
        L.append("}")
    if pm > 0:
        L.append(("X_ERROR" if basis == "Z" else "Z_ERROR") + f"({pm}) " + j(data))
    L.append(("M " if basis == "Z" else "MX ") + j(data))
    na = len(anc)
    checks = list(zip(anc_z, zs)) if basis == "Z" else list(zip(anc_x, xs))
    for a, sup in checks:
        k = anc.index(a)
        recs = [f"rec[{q - n}]" for q in sup] + [f"rec[{k - na - n}]"]
        L.append("DETECTOR " + " ".join(recs))
    # logical operator of the measured basis, as chosen by `logical_operator` (code_distance.jl:24-100)
    from .tanner import logical_operator
    lx, lz = logical_operator(t)
    logical = [int(q) for q in (lz[0] if basis == "Z" else lx[0]).nonzero()[0]]
    L.append("OBSERVABLE_INCLUDE(0) " + " ".join(f"rec[{q - n}]" for q in logical))
    return "\n".join(L) + "\n"
