// Register-level shapes of a sweep super-step compiled into k_sweep.  Mirror of MENU in tensorqec.jl_b200/sweep.py
// (tests/test_sweep_cpu.py parses this list and compares).  One line per shape:
//   X(id, M, NL,  NP0, P00, P01, NF0, F00, F01, K00, K01,   NP1, P10, P11, NF1, F10, F11, K10, K11)
// M = patch bits, NL = layers; per layer: NP pinned variables whose value is patch bit P.. and which flip the other
// patch bits K.. when set, NF free variables with patch-local flip masks F.. (unused entries: -1 for bits, 0 for masks).
#pragma once
#define TQEC_SWEEP_MENU(X) \
  X(0, 2, 1, 0, -1, -1, 2, 1, 2, 0, 0, 0, -1, -1, 0, 0, 0, 0, 0) \
  X(1, 2, 1, 1, 0, -1, 1, 2, 0, 0, 0, 0, -1, -1, 0, 0, 0, 0, 0) \
  X(2, 2, 2, 1, 0, -1, 1, 2, 0, 0, 0, 0, -1, -1, 2, 1, 2, 0, 0) \
  X(3, 2, 2, 1, 0, -1, 1, 2, 0, 0, 0, 0, -1, -1, 2, 2, 1, 0, 0) \
  X(4, 2, 2, 1, 0, -1, 1, 2, 0, 0, 0, 1, 0, -1, 1, 2, 0, 0, 0) \
  X(5, 3, 2, 0, -1, -1, 2, 1, 2, 0, 0, 1, 0, -1, 1, 4, 0, 0, 0) \
  X(6, 3, 1, 0, -1, -1, 2, 3, 4, 0, 0, 0, -1, -1, 0, 0, 0, 0, 0) \
  X(7, 3, 1, 1, 0, -1, 1, 6, 0, 0, 0, 0, -1, -1, 0, 0, 0, 0, 0) \
  X(8, 3, 2, 1, 0, -1, 1, 2, 0, 0, 0, 0, -1, -1, 2, 5, 2, 0, 0) \
  X(9, 3, 2, 1, 0, -1, 1, 2, 0, 0, 0, 1, 1, -1, 1, 5, 0, 0, 0) \
  X(10, 3, 2, 1, 0, -1, 1, 6, 0, 0, 0, 1, 1, -1, 1, 1, 0, 0, 0) \
  X(11, 4, 2, 1, 0, -1, 1, 6, 0, 0, 0, 1, 1, -1, 1, 9, 0, 0, 0) \
  /* 12..19: shapes with FRESH pins (K contains the pinned variable's own bit: the opened check takes a dead slot, output \
     bit 1 reads the live half) and their neighbours -- even-distance and rectangular rotated surface codes */ \
  X(12, 3, 1, 2, 0, 1, 0, 0, 0, 5, 2, 0, -1, -1, 0, 0, 0, 0, 0) \
  X(13, 3, 1, 0, -1, -1, 2, 1, 6, 0, 0, 0, -1, -1, 0, 0, 0, 0, 0) \
  X(14, 4, 2, 1, 0, -1, 1, 6, 0, 0, 0, 1, 3, -1, 1, 2, 0, 0, 0) \
  X(15, 3, 2, 2, 0, 1, 0, 0, 0, 5, 2, 0, -1, -1, 2, 1, 2, 0, 0) \
  X(16, 4, 2, 1, 0, -1, 1, 2, 0, 0, 0, 1, 2, -1, 1, 10, 0, 0, 0) \
  X(17, 3, 2, 2, 0, 1, 0, 0, 0, 5, 2, 1, 0, -1, 1, 2, 0, 0, 0) \
  X(18, 3, 2, 0, -1, -1, 2, 1, 2, 0, 0, 1, 1, -1, 1, 4, 0, 0, 0) \
  X(19, 3, 2, 1, 0, -1, 1, 2, 0, 0, 0, 0, -1, -1, 2, 2, 5, 0, 0) \
  X(20, 4, 1, 1, 0, -1, 1, 14, 0, 0, 0, 0, -1, -1, 0, 0, 0, 0, 0) \
  X(21, 4, 1, 1, 0, -1, 1, 6, 0, 8, 0, 0, -1, -1, 0, 0, 0, 0, 0) \
  X(22, 4, 2, 1, 0, -1, 1, 2, 0, 0, 0, 1, 1, -1, 1, 5, 0, 8, 0) \
  X(23, 4, 2, 1, 0, -1, 1, 2, 0, 0, 0, 1, 2, -1, 1, 10, 0, 1, 0) \
  X(24, 4, 2, 1, 0, -1, 1, 6, 0, 0, 0, 0, -1, -1, 2, 9, 6, 0, 0) \
  /* 25..31: sum-product only, extended instantiation: TNMMAP plans of even-distance and rectangular codes */ \
  X(25, 4, 1, 0, -1, -1, 2, 3, 12, 0, 0, 0, -1, -1, 0, 0, 0, 0, 0) \
  X(26, 3, 1, 2, 0, 1, 0, 0, 0, 1, 6, 0, -1, -1, 0, 0, 0, 0, 0) \
  X(27, 4, 2, 2, 0, 1, 0, 0, 0, 1, 6, 1, 2, -1, 1, 9, 0, 0, 0) \
  X(28, 4, 2, 2, 0, 1, 0, 0, 0, 1, 6, 0, -1, -1, 2, 9, 6, 0, 0) \
  X(29, 3, 2, 1, 0, -1, 1, 6, 0, 0, 0, 1, 0, -1, 1, 6, 0, 0, 0) \
  X(30, 4, 2, 1, 0, -1, 1, 6, 0, 0, 0, 0, -1, -1, 2, 6, 9, 0, 0) \
  X(31, 3, 2, 1, 0, -1, 1, 6, 0, 0, 0, 0, -1, -1, 2, 6, 1, 0, 0)
#define TQEC_SWEEP_MENU_SIZE 32
#define TQEC_SWEEP_MENU_MAXPLUS 20 /* shapes 0..19 are compiled into the max-plus kernel; 20.. are sum-product only */
#define TQEC_SWEEP_MENU_BASE 12    /* shapes 12..19 only in the extended instantiation of k_sweep (plans with fresh pins) */
#define TQEC_SWEEP_MENU_SP_BASE 25 /* sum-product shapes 25.. only in the extended instantiation as well */
