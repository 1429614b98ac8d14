"""Stim circuit parser and circuit -> detector error model (tensorqec.jl_b200/circuit.py) against the reference's
known answers (test/decoding/dem.jl:5-109, ids shifted to 0-based) and a stabilizer-tableau check of the generated
surface-code memory circuits."""
import os

import numpy as np
import pytest

from oracle import chp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _as_dict(dem):
    return {tuple(sorted(f)): p for p, f in zip(dem.error_rates, dem.flipped_detectors)}


def test_dem_golden_bit_flip_and_depolarizing(tq):
    from tensorqec.jl_b200 import circuit as C
    # test/decoding/dem.jl:5-24: error_rates [0.212, 0.212], flipped_detectors [[1,3,4],[1,2]] (1-based)
    qc = C.parse_stim_string("""
        X_ERROR(0.2) 0 1
        DEPOLARIZE1(0.03) 0 1
        CX 0 1
        M 0 1
        DETECTOR rec[-1]
        DETECTOR rec[-2]
        DETECTOR rec[-1] rec[-2]
        OBSERVABLE_INCLUDE(0) rec[-1] rec[-2]
    """, 2)
    got = _as_dict(C.detector_error_model(qc))
    assert set(got) == {(0, 2, 3), (0, 1)}
    assert all(abs(p - 0.212) < 1e-10 for p in got.values())
    # :26-37
    got = _as_dict(C.detector_error_model(C.parse_stim_string("DEPOLARIZE1(0.3) 0\nM 0\nDETECTOR rec[-1]\n", 1)))
    assert set(got) == {(0,)} and abs(got[(0,)] - 0.2) < 1e-10
    # :39-56
    qc = C.parse_stim_string("""
        H 0
        CX 0 1
        DEPOLARIZE1(0.0297) 0
        H 3
        CX 0 2 1 2 3 0 3 1
        H 3
        M 2 3
        DETECTOR rec[-2]
        DETECTOR rec[-1]
    """, 4)
    got = _as_dict(C.detector_error_model(qc))
    assert set(got) == {(0,), (0, 1), (1,)} and all(abs(p - 0.01) < 1e-10 for p in got.values())


def test_dem_golden_with_reset(tq):
    from tensorqec.jl_b200 import circuit as C
    # test/decoding/dem.jl:58-82: error_rates [0.2, 0.32], flipped_detectors [[1],[2]]
    qc = C.parse_stim_string("""
        X_ERROR(0.2) 0 1
        CX 0 1
        M 0 1
        DETECTOR rec[-1] rec[-2]
        R 1
        X_ERROR(0.2) 0 1
        CX 1 0
        M 0 1
        DETECTOR rec[-1] rec[-2]
    """, 2)
    got = _as_dict(C.detector_error_model(qc))
    assert set(got) == {(0,), (1,)}
    assert abs(got[(0,)] - 0.2) < 1e-10 and abs(got[(1,)] - 0.32) < 1e-10


def test_dem_golden_stim_generated_color_code(tq):
    """test/decoding/dem.jl:84-109: color_code:memory_xyz, 2 rounds, distance 3 (REPEAT, C_XYZ, MR, MY) against the DEM
    stim itself produced for it (21 mechanisms, atol 1e-10)."""
    from tensorqec.jl_b200 import circuit as C
    qc = C.parse_stim_file(os.path.join(GOLD, "color_memory_xyz_d3_r2.stim"), 10)
    assert qc.n_detectors == 6 and qc.n_observables == 1 and qc.n_measurements == 13
    d1 = C.detector_error_model(qc)
    d2 = tq.parse_dem_file(os.path.join(GOLD, "color_memory_xyz_d3_r2.dem"))
    assert d1.detector_list == d2.detector_list and d1.logical_list == d2.logical_list
    a, b = _as_dict(d1), _as_dict(d2)
    assert set(a) == set(b) and len(a) == 21
    assert all(abs(a[k] - b[k]) < 1e-10 for k in a)
    # text round trip
    d3 = tq.parse_dem_string(C.dem_to_string(d1))
    assert _as_dict(d3) == a


def test_parser_errors_and_repeat(tq):
    from tensorqec.jl_b200 import circuit as C
    with pytest.raises(ValueError, match="Unknown instruction"):
        C.parse_stim_string("FOO 0\n")
    with pytest.raises(ValueError, match="looks back"):
        C.parse_stim_string("M 0\nDETECTOR rec[-2]\n")
    with pytest.raises(ValueError, match="over-mixing"):
        C.detector_error_model(C.parse_stim_string("DEPOLARIZE1(0.8) 0\nM 0\nDETECTOR rec[-1]\n"))
    c = C.parse_stim_string("R 0 1\nREPEAT 3 {\n  CX 0 1\n  REPEAT 2 {\n    MR 1\n  }\n}\nM 0\n")
    assert c.n_measurements == 7 and c.count("CX") == 3 and c.n_qubits == 2


@pytest.mark.parametrize("d,rounds,basis", [(3, 3, "Z"), (3, 2, "X"), (5, 2, "Z")])
def test_surface_memory_circuit_is_deterministic_and_detects_single_faults(tq, d, rounds, basis):
    from tensorqec.jl_b200 import circuit as C
    txt = C.surface_memory_circuit(d, rounds, basis, after_clifford_depolarization=1e-3,
                                   before_round_data_depolarization=1e-3, before_measure_flip_probability=1e-3,
                                   after_reset_flip_probability=1e-3)
    c = C.parse_stim_string(txt)
    na = d * d - 1
    assert c.n_qubits == 2 * d * d - 1 and c.n_detectors == (na // 2) * 2 + na * (rounds - 1)
    for seed in range(3):                                        # random outcomes of the first X/Z-check round
        det, obs = chp.run_noiseless(c, seed)
        assert not any(det) and obs == {0: 0}
    dem = C.detector_error_model(c)
    nd = c.n_detectors
    assert all(0.0 < p < 0.5 for p in dem.error_rates)
    # no single fault flips the observable without tripping a detector (fault distance > 1), and every mechanism
    # flips at most 2 detectors of each basis plus hook pairs: <= 6 here
    assert all(any(x < nd for x in f) for f in dem.flipped_detectors)
    assert max(len(f) for f in dem.flipped_detectors) <= 6
    # the noiseless DEM is empty, phenomenological knobs give the textbook mechanism count
    c0 = C.parse_stim_string(C.surface_memory_circuit(d, rounds, basis))
    assert len(C.detector_error_model(c0).error_rates) == 0


def test_generated_dem_lowers_to_a_tnmmap_schedule(tq):
    from tensorqec.jl_b200 import circuit as C, decoding as D
    c = C.parse_stim_string(C.surface_memory_circuit(3, 3, "Z", after_clifford_depolarization=1e-3,
                                                     before_round_data_depolarization=1e-3,
                                                     before_measure_flip_probability=1e-3, after_reset_flip_probability=1e-3))
    dem = C.detector_error_model(c)
    tanner, l2q, sch, R, L, FIX = D.tnmmap_dem_schedule(tq.TNMMAP(), dem)
    assert sch.n_obs == 1 and sch.n_checks == 24 and sch.w_max <= 13
    # the coset-representative matrix solves H e = s for every detector pattern the mechanisms can produce
    H = tanner.H
    assert np.array_equal((H @ R) % 2 @ H % 2, H % 2)


@pytest.mark.parametrize("name,nq,n_meas,n_det", [("teleportation", 100, 2, 0), ("repetition", 7, 3007, 3006),
                                                   ("noisy_repetition", 7, 3004, 3003), ("noisy_surface", 26, 8009, 8000)])
def test_reference_stim_fixtures_parse_and_round_trip(tq, golden_dir, tmp_path, name, nq, n_meas, n_det):
    """The circuits of test/stim_parser/stim_parser.jl:52-75 (stim's own documentation examples): REPEAT blocks of 1000
    rounds flattened, `rec[-k]` look-back, classically controlled Paulis (teleportation); `dump_stim_file`
    (stim_parser.jl:378-437) writes a file that parses back to the same instruction list."""
    circ = tq.parse_stim_file(str(golden_dir / "stim" / f"{name}.stim"), nq)
    assert (circ.n_qubits, circ.n_measurements, circ.n_detectors) == (nq, n_meas, n_det)
    out = tmp_path / "dump.stim"
    tq.dump_stim_file(circ, str(out))
    back = tq.parse_stim_file(str(out), nq)
    assert back.instructions == circ.instructions and back.n_observables == circ.n_observables


def test_feedback_pauli_enters_the_error_model(tq):
    """A measurement whose outcome steers a later Pauli: flipping the record applies the Pauli wrongly, so the record
    flip inherits the detectors that Pauli flips.  Here M(0.1) on qubit 0 controls an X on qubit 1, which is then
    measured and compared with a detector: the mechanism {D0} has probability 0.1."""
    circ = tq.parse_stim_string("R 0 1\nM(0.1) 0\nCX rec[-1] 1\nM 1\nDETECTOR rec[-1]\n")
    dem = tq.detector_error_model(circ)
    assert dem.flipped_detectors == [[0]] and abs(dem.error_rates[0] - 0.1) < 1e-15
    # without the feedback the measurement error is invisible
    dem0 = tq.detector_error_model(tq.parse_stim_string("R 0 1\nM(0.1) 0\nM 1\nDETECTOR rec[-1]\n"))
    assert dem0.flipped_detectors == []


def test_noisy_repetition_fixture_gives_a_chain_dem(tq, golden_dir):
    """stim's noisy repetition-code example (1000 rounds): every mechanism of its detector error model flips at most two
    detectors (a matching graph), and the observable is touched."""
    circ = tq.parse_stim_file(str(golden_dir / "stim" / "noisy_repetition.stim"), 7)
    dem = tq.detector_error_model(circ)
    assert dem.n_detectors == 3003 and dem.n_observables == 1
    assert max(sum(1 for d in f if d < 3003) for f in dem.flipped_detectors) <= 2
    assert any(3003 in f for f in dem.flipped_detectors)
