"""tensorqec.jl_b200 -- B200-native TNMAP / TNMMAP decoding hot path of TensorQEC.jl behind the reference's API.

Host side (this package) mirrors the reference's operator interface for the path -- same names, same argument
meaning, 0-based indices -- and marshals bit-packed batches through the C ABI of `libtqec_cuda.so`
(include/tqec.h).  All arithmetic of the hot path runs in hand-written sm_100a CUDA kernels; there is no CPU
fallback (a missing library or GPU raises `TqecError`).
"""
from ._cabi import TqecError, LIB_PATH
from .codes import CSSQuantumCode, Color488, QuantumCode, SteaneCode, SurfaceCode, stabilizers
from .decoding import (TNMAP, TNMMAP, AbstractDecoder, AbstractGeneralDecoder, ClassicalDecodingProblem,
                       CompiledDecoder, CompiledDEMTNMMAP, CompiledGeneralDecoder, CompiledTNMAP, CompiledTNMMAP,
                       CSSToGeneralDecodingProblem, DecodingResult, GeneralDecodingProblem,
                       IndependentDepolarizingDecodingProblem, NoOptimizer, SimpleTensorNetwork, compile, decode,
                       extract_decoding, get_problem, reduce2general, single_qubit_tensor, tnmap_schedule,
                       tnmmap_css_schedule, tnmmap_dem_schedule)
from .circuit import (StimCircuit, circuit_to_string, dem_to_string, detector_error_model, dump_stim_file, parse_stim_file,
                      parse_stim_string, surface_memory_circuit)
from .dem import DetectorErrorModel, dem2tanner, parse_dem_file, parse_dem_string
from .bposd import BPDecoder
from .truthtable import TableDecoder, TruthTable, load_table, make_table, save_table
from .encoder import (CompiledInference, CSSBimatrix, clifford_network, correction_pauli_string, encode_circuit, encode_stabilizers,
                      generate_syndrome_dict, inference, pauli_string_map_iter, stabilizers2bimatrix, syndrome_inference,
                      syndrome_transform)
from .error_model import (CSSErrorPattern, CSSSyndrome, IndependentDepolarizingError, IndependentFlipError,
                          SimpleSyndrome, check_logical_error, iid_error, random_error_pattern, syndrome_extraction)
from .mod2 import Mod2, bitmul, pack_bits, unpack_bits
from .tanner import (CSSTannerGraph, SimpleTannerGraph, StabilizerList, gf2_right_inverse, logical_operator, nq, ns,
                     null_space, row_echelon_form, same_qubit_order)
from .threshold import multi_round_qec, MonteCarlo
from .sharding import shard_range, multiprocess_run
