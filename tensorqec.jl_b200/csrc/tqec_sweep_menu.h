// Register-level shapes of a sweep super-step compiled into k_sweep.  Mirror of MENU in tensorqec.jl_b200/sweep.py
// (tests/test_sweep_cpu.py parses this list and compares).  One line per shape:
//   X(id, M, NL,  NP0, P00, P01, NF0, F00, F01,   NP1, P10, P11, NF1, F10, F11)
// M = patch bits, NL = layers; per layer: NP pinned variables whose value is patch bit P.., NF free variables with
// patch-local flip masks F.. (unused entries: -1 for bits, 0 for masks).
#pragma once
#define TQEC_SWEEP_MENU(X) \
  X(0, 2, 1, 0, -1, -1, 2, 1, 2, 0, -1, -1, 0, 0, 0) \
  X(1, 2, 1, 1, 0, -1, 1, 2, 0, 0, -1, -1, 0, 0, 0) \
  X(2, 2, 2, 1, 0, -1, 1, 2, 0, 0, -1, -1, 2, 1, 2) \
  X(3, 2, 2, 1, 0, -1, 1, 2, 0, 0, -1, -1, 2, 2, 1) \
  X(4, 2, 2, 1, 0, -1, 1, 2, 0, 1, 0, -1, 1, 2, 0) \
  X(5, 3, 2, 0, -1, -1, 2, 1, 2, 1, 0, -1, 1, 4, 0) \
  X(6, 3, 1, 0, -1, -1, 2, 3, 4, 0, -1, -1, 0, 0, 0) \
  X(7, 3, 1, 1, 0, -1, 1, 6, 0, 0, -1, -1, 0, 0, 0) \
  X(8, 3, 2, 1, 0, -1, 1, 2, 0, 0, -1, -1, 2, 5, 2) \
  X(9, 3, 2, 1, 0, -1, 1, 2, 0, 1, 1, -1, 1, 5, 0) \
  X(10, 3, 2, 1, 0, -1, 1, 6, 0, 1, 1, -1, 1, 1, 0) \
  X(11, 4, 2, 1, 0, -1, 1, 6, 0, 1, 1, -1, 1, 9, 0)
#define TQEC_SWEEP_MENU_SIZE 12
