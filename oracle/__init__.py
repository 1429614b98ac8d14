"""CPU oracle for the TNMAP / TNMMAP hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under `oracle/` is part of the product: only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import, link or execute it, and only as the checker or as the timed
CPU baseline.  The product path (`tensorqec.jl_b200`) never imports this package and fails loudly without the
CUDA library.

PARITY PINNING.  The reference (nzy1997/TensorQEC.jl v2.2.1) is pure Julia and its contraction arithmetic lives in
un-vendored third-party packages (TensorInference.jl compat "0.6", OMEinsum.jl compat "0.9", SCIP 0.12.3 /
JuMP 1.24; Project.toml:35-45, no Manifest).  No Julia toolchain exists in the build image or on the GPU box, so
the reference cannot be executed.  The oracle therefore restates the published algorithms and is pinned against
every known-answer value the reference's own tests hold for this path (tests/test_oracle_golden.py):
  * test/decoding/tndecoder.jl:8-14      parity_check_matrix(4) entries
  * test/decoding/tndecoder.jl:101-110   TNMMAP marginal of SurfaceCode(3,3), px=0.1 (atol 1e-10)   <- the only
                                          numeric contraction golden
  * test/codes/code_distance.jl:54-59    logical_operator(SurfaceCode(3,3))
  * test/codes/ldpc.jl:7-23, 34-39       SimpleTannerGraph fields, syndrome_extraction known answer
  * test/decoding/error_model.jl:21-33   check_logical_error known answers
  * test/codes/mod2.jl:3-31              Mod2 algebra, bitmul! == A*B
  * test/codes/codes.jl:176-181          Color488(5): 16 stabilizers
  * test/stim_parser/test_circuits/dem.dem   DEM text fixture (21 mechanisms, 6 detectors, 1 observable)
TNMAP output patterns, MAP tie-breaking and DEM marginals are NOT pinned by any reference test ("parity
unpinned" for those): there the oracle is additionally cross-checked by exhaustive enumeration (bruteforce.py).

Modules
  gf2.py          GF(2) kernels: syndrome extraction, logical check, packed product (error_model.jl, mod2.jl)
  philox.py       Philox4x32-10 counter-based sampler + the reference's threshold rule (error_model.jl:97-117)
  networks.py     the reference's tensor networks, label for label (tndecoder.jl:33-50, 97-146, 186-238)
  dense.py        greedy pairwise contraction of those networks: sum-product and max-plus + traceback
  bruteforce.py   exhaustive enumeration: exact marginals, MAP value and the full set of maximisers
  frontier.py     the frontier recurrence the CUDA kernels execute, written independently over named axes
  emulator.py     table-level emulator of the lowered schedule (checks the lowering itself on CPU)
  csrc/           C restatements (dense executor, frontier recurrence) used as the timed CPU baselines
"""
