// In-place patch sweep (k_sweep): the frontier recurrence with static slots, register patches and a tabulated head.
// Host side of the lowering and the full description: tensorqec.jl_b200/sweep.py.  Summary of what a warp team does for
// one pass (2^sg shots, 1024 FP64 state entries = 8 KiB of shared memory, updated in place):
//   1. copy the head-table row selected by the shots' head syndrome bits into the state (swizzled index, see sw_phys);
//   2. per super-step: every lane loads 2^M entries that differ in the step's patch bits (address = lane part ^ loop
//      part ^ closed-syndrome part ^ patch part, all XOR of host-made byte masks), absorbs one or two factors on them in
//      registers (sweep_layer: Out[j] = max_k R[j ^ F(k)] + T[p(j)][k], static register indices), stores them back to
//      the same addresses and emits the packed back-pointer bits of the patch;
//   3. after 32 shots (deferred traceback) lane q walks shot q backwards through the traceback records.
// Bit-identical to the unfused schedule of schedule.py (same IEEE adds in the same order, strict > on ascending
// candidates).  Shapes of a super-step come from tqec_sweep_menu.h.  Sum-product plans (TNMMAP) run the same code with
// multiply / add layers, no back-pointers and no traceback; their open observable slots index the output marginals.
// All teams of a CTA run the same number of rounds and meet at a CTA barrier once per group of 32 shots (sync_mode):
// passes have identical instruction streams, so teams that stay in step share instruction-cache lines.
#include <cstdlib>
#include <cstring>
#include <utility>

#include "tqec_common.h"
#include "tqec_sweep_menu.h"

namespace tqec {

#define SW_REC_INTS 32
#define SW_TB_INTS 64
#ifndef TQEC_SWEEP_STORE_MASK
#define TQEC_SWEEP_STORE_MASK 0x11   /* outputs (index mod 8) of a last layer that are stored as a predicated pair instead of selected: a
                                        predicated-off store still takes a shared-memory wavefront slot, so the two forms are mixed to load
                                        the ALU pipe and the LSU evenly (same-box A/B, benchmarks/ab_store.sh: 0x00 72.9, 0xff 73.1, 0x55 73.5,
                                        0x33 74.0, 0x11 74.1 M syndromes/s at d = 9) */
#endif

__device__ __forceinline__ double sw_lds(uint32_t addr) {
  double v;
#ifdef TQEC_DIAG_NOMEM   // diagnosis only (wrong results): no shared-memory traffic, the value depends on the address
  v = __hiloint2double((int)addr | 0xbff00000, (int)addr);
#else
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
#endif
  return v;
}
__device__ __forceinline__ void sw_sts(uint32_t addr, double v) {
#ifdef TQEC_DIAG_NOMEM
  if (v == 1.2345 && addr == 77) asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v));
#else
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v));
#endif
}
__device__ __forceinline__ uint32_t sw_xor3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ int32_t sw_lds32(uint32_t addr) {
  int32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ int4 sw_lds128(uint32_t addr) {
  int4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
// ---- TMA (1-D bulk copy) of a tabulated head row into the team's state, completion on an mbarrier ----------------------------
__device__ __forceinline__ void sw_mbar_init(uint32_t bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void sw_mbar_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sw_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void sw_mbar_wait(uint32_t bar, uint32_t phase) {
  asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}"
               ::"r"(bar), "r"(phase) : "memory");
}

// swizzled entry index: bits 4..7 XOR-ed into bits 0..3 (sweep.py:phys)
__host__ __device__ __forceinline__ uint32_t sw_phys(uint32_t x) { return x ^ ((x >> 4) & 15u); }

// One factor absorbed on a register patch.  Output j reads R[j ^ F(k)] for the 2^NF assignments k of the free variables
// (ascending k = ascending assignment; strict > keeps the smallest on ties) with the table row picked by the pinned bits
// of j.  Back-pointer bits of output j: k at bit OFF + j * NF.  The output index J is a template parameter so that the
// back-pointer mask is an immediate of a predicated OR (one issue slot per bit; the compiler's own lowering of
// `if (p) bits |= m` costs a SEL plus a share of a LOP3).
template <int SEMI, int M, int NP, int P0, int P1, int NF, int F0, int F1, int K0, int K1, int OFF, int JO>
__device__ __forceinline__ void sweep_out(const double (&R)[1 << M], double (&O)[1 << M], const double *T, uint32_t &bits) {
  constexpr int pidx = (NP > 0 ? ((JO >> (P0 < 0 ? 0 : P0)) & 1) : 0) | (NP > 1 ? (((JO >> (P1 < 0 ? 0 : P1)) & 1) << 1) : 0);
  // output JO reads the inputs J ^ F(k), J = JO with the other bits flipped that its pinned variables (when 1) flip
  constexpr int J = JO ^ ((pidx & 1) ? K0 : 0) ^ ((pidx & 2) ? K1 : 0);
  const double *Tp = T + (pidx << NF);
  if (NF == 0) {
    O[JO] = SEMI == TQEC_SEMIRING_MAXPLUS ? R[J] + Tp[0] : R[J] * Tp[0];
  } else if (NF == 1) {
    if (SEMI == TQEC_SEMIRING_MAXPLUS) {
      constexpr uint32_t m = 1u << ((OFF + JO) & 31);
      // c0 / c1 in ascending assignment order, strict > keeps the smaller assignment on ties; the back-pointer bit is one
      // predicated OR with an immediate mask.  (A variant that recomputes the winner with a predicated add instead of
      // selecting it -- FP64 pipe instead of ALU pipe -- cannot be expressed: ptxas merges the second add with the first
      // and emits the same DSETP + 2 FSEL, also for a DFMA by an opaque 1.0, which it hoists out of the predicate.)
      const double c0 = R[J] + Tp[0], c1 = R[J ^ F0] + Tp[1];
      asm("{\n .reg .pred p;\n setp.gt.f64 p, %2, %3;\n selp.f64 %0, %2, %3, p;\n @p or.b32 %1, %1, %4;\n}"
          : "=d"(O[JO]), "+r"(bits) : "d"(c1), "d"(c0), "n"(m));
    } else {
      O[JO] = R[J] * Tp[0] + R[J ^ F0] * Tp[1];
    }
  } else {
    if (SEMI == TQEC_SEMIRING_MAXPLUS) {
      const double c0 = R[J] + Tp[0], c1 = R[J ^ F0] + Tp[1], c2 = R[J ^ F1] + Tp[2], c3 = R[J ^ F0 ^ F1] + Tp[3];
      const bool p01 = c1 > c0, p23 = c3 > c2;
      const double b01 = p01 ? c1 : c0, b23 = p23 ? c3 : c2;
      const bool pf = b23 > b01;
      O[JO] = pf ? b23 : b01;
      const uint32_t bk = pf ? (2u | (uint32_t)p23) : (uint32_t)p01;
      bits |= bk << ((OFF + 2 * JO) & 31);
    } else {
      O[JO] = ((R[J] * Tp[0] + R[J ^ F0] * Tp[1]) + R[J ^ F1] * Tp[2]) + R[J ^ F0 ^ F1] * Tp[3];
    }
  }
}

template <int SEMI, int M, int NP, int P0, int P1, int NF, int F0, int F1, int K0, int K1, int OFF, int... J>
__device__ __forceinline__ void sweep_layer_seq(double (&R)[1 << M], const double *T, uint32_t &bits, std::integer_sequence<int, J...>) {
  double O[1 << M];
  (sweep_out<SEMI, M, NP, P0, P1, NF, F0, F1, K0, K1, OFF, J>(R, O, T, bits), ...);
#pragma unroll
  for (int j = 0; j < (1 << M); ++j) R[j] = O[j];
}

template <int SEMI, int M, int NP, int P0, int P1, int NF, int F0, int F1, int K0, int K1, int OFF>
__device__ __forceinline__ void sweep_layer(double (&R)[1 << M], const double *T, uint32_t &bits) {
  sweep_layer_seq<SEMI, M, NP, P0, P1, NF, F0, F1, K0, K1, OFF>(R, T, bits, std::make_integer_sequence<int, (1 << M)>{});
}

// The LAST layer of a super-step writes its outputs straight to the state.  For the common max-plus form (one free
// variable) the winner is not selected into a register first: the two candidates are stored under complementary
// predicates (`@p st c1; @!p st c0`), which replaces two FSEL on the ALU pipe -- the busiest pipe of this kernel -- and
// the separate store by two predicated stores (ptxas keeps predicated stores as they are; it turns every predicated
// register write into a select).  Same values, same tie rule.
template <int SEMI, int M, int NP, int P0, int P1, int NF, int F0, int F1, int K0, int K1, int OFF, int JO>
__device__ __forceinline__ void sweep_out_store(const double (&R)[1 << M], const double *T, uint32_t &bits, uint32_t addr) {
  constexpr int pidx = (NP > 0 ? ((JO >> (P0 < 0 ? 0 : P0)) & 1) : 0) | (NP > 1 ? (((JO >> (P1 < 0 ? 0 : P1)) & 1) << 1) : 0);
  constexpr int J = JO ^ ((pidx & 1) ? K0 : 0) ^ ((pidx & 2) ? K1 : 0);
  const double *Tp = T + (pidx << NF);
  if (SEMI == TQEC_SEMIRING_MAXPLUS && NF == 1 && (((TQEC_SWEEP_STORE_MASK) >> (JO & 7)) & 1)) {
    constexpr uint32_t m = 1u << ((OFF + JO) & 31);
    const double c0 = R[J] + Tp[0], c1 = R[J ^ F0] + Tp[1];
#ifdef TQEC_DIAG_NOMEM
    if (c0 == 1.2345 && addr == 77) asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(c1));
    asm("{\n .reg .pred p;\n setp.gt.f64 p, %1, %2;\n @p or.b32 %0, %0, %3;\n}" : "+r"(bits) : "d"(c1), "d"(c0), "n"(m));
#else
    asm volatile("{\n .reg .pred p;\n setp.gt.f64 p, %1, %2;\n @p st.shared.f64 [%3], %1;\n @!p st.shared.f64 [%3], %2;\n @p or.b32 %0, %0, %4;\n}"
                 : "+r"(bits) : "d"(c1), "d"(c0), "r"(addr), "n"(m));
#endif
  } else {
    double O[1 << M];
    sweep_out<SEMI, M, NP, P0, P1, NF, F0, F1, K0, K1, OFF, JO>(R, O, T, bits);
    sw_sts(addr, O[JO]);
  }
}

template <int SEMI, int M, int NP, int P0, int P1, int NF, int F0, int F1, int K0, int K1, int OFF, int... J>
__device__ __forceinline__ void sweep_layer_store_seq(const double (&R)[1 << M], const double *T, uint32_t &bits, uint32_t outb,
                                                      const uint32_t (&lo)[4], const uint32_t (&hi)[4],
                                                      std::integer_sequence<int, J...>) {
  (sweep_out_store<SEMI, M, NP, P0, P1, NF, F0, F1, K0, K1, OFF, J>(R, T, bits, sw_xor3(outb, lo[J & 3], hi[(J >> 2) & 3])), ...);
}

template <int SEMI, int M, int NP, int P0, int P1, int NF, int F0, int F1, int K0, int K1, int OFF>
__device__ __forceinline__ void sweep_layer_store(const double (&R)[1 << M], const double *T, uint32_t &bits, uint32_t outb,
                                                  const uint32_t (&lo)[4], const uint32_t (&hi)[4]) {
  sweep_layer_store_seq<SEMI, M, NP, P0, P1, NF, F0, F1, K0, K1, OFF>(R, T, bits, outb, lo, hi, std::make_integer_sequence<int, (1 << M)>{});
}

template <int SEMI, int M, int NL, int NP0, int P00, int P01, int NF0, int F00, int F01, int K00, int K01, int NP1,
          int P10, int P11, int NF1, int F10, int F11, int K10, int K11>
__device__ __forceinline__ void sweep_step(const int32_t *__restrict__ rec, const double *__restrict__ tvals,
                                           uint32_t st_abs, uint32_t lt, const uint16_t *__restrict__ stab_row,
                                           const uint16_t *__restrict__ late_row, uint32_t *__restrict__ bpt, int lane) {
  constexpr int N = 1 << M;
  constexpr int NT0 = 1 << (NP0 + NF0), NT1 = NL > 1 ? (1 << (NP1 + NF1)) : 1;
  // back-pointers exist for max-plus only; shapes whose bits would not fit one word are sum-product-only shapes
  constexpr int BPP0 = SEMI == TQEC_SEMIRING_MAXPLUS ? N * (NF0 + (NL > 1 ? NF1 : 0)) : 0;
  constexpr int BPP = BPP0 <= 32 ? BPP0 : 0;
  constexpr int IPW = BPP ? 32 / BPP : 1;
  const int4 r0 = *reinterpret_cast<const int4 *>(rec);         // menu id, iterations, table offset, first bp word
  const int2 am = *reinterpret_cast<const int2 *>(rec + 4);     // byte masks of patch bits 0..3 (u16 each)
  const uint32_t a0 = (uint32_t)am.x & 0xffffu, a1 = (uint32_t)am.x >> 16, a2 = (uint32_t)am.y & 0xffffu, a3 = (uint32_t)am.y >> 16;
  const uint32_t lo[4] = {0u, a0, a1, a0 ^ a1};
  const uint32_t hi[4] = {0u, a2, a3, a2 ^ a3};
  double T0[NT0];
  const double *tv = tvals + r0.z;
#pragma unroll
  for (int i = 0; i < NT0; ++i) T0[i] = tv[i];
  const uint32_t laddr = st_abs ^ (lt & 0xffffu);
  const uint32_t lsub = lt >> 16;
  const uint16_t *la = reinterpret_cast<const uint16_t *>(rec + 8);
  const uint8_t *ls = reinterpret_cast<const uint8_t *>(rec + 12);
  uint32_t word = 0;
  uint32_t sidx = lsub | (uint32_t)ls[0];
  uint32_t base = laddr ^ (uint32_t)la[0];
  uint32_t inb = base ^ (uint32_t)stab_row[sidx];
#pragma unroll 1
  for (int it = 0; it < r0.y; ++it) {
    double R[N];
#pragma unroll
    for (int j = 0; j < N; ++j) R[j] = sw_lds(sw_xor3(inb, lo[j & 3], hi[(j >> 2) & 3]));
    // late syndrome bit of the shot (a check opened by layer 0 and closed by layer 1): byte mask for the store address,
    // bit 15 = use the row-swapped copy of layer 1's table
    const uint32_t late = NL > 1 ? (uint32_t)late_row[sidx] : 0u;
    double T1[NT1];
    const double *t1 = tv + NT0 + ((late >> 15) ? NT1 : 0);
#pragma unroll
    for (int i = 0; i < NT1; ++i) T1[i] = NL > 1 ? t1[i] : 0.0;
    const uint32_t outb = base ^ (late & 0x3fffu);
    // the next iteration's addresses (two dependent table reads) are looked up while this patch is in flight
    const int itn = it + 1 < r0.y ? it + 1 : it;
    const uint32_t sidx_n = lsub | (uint32_t)ls[itn];
    const uint32_t base_n = laddr ^ (uint32_t)la[itn];
    const uint32_t inb_n = base_n ^ (uint32_t)stab_row[sidx_n];
    uint32_t bits = 0;
#ifndef TQEC_DIAG_NOLAYERS   // diagnosis only (wrong results): load / store skeleton without the arithmetic
#ifdef TQEC_SWEEP_SELECT_THEN_STORE   // the round-1 form: every layer selects into registers, one store loop at the end
    sweep_layer<SEMI, M, NP0, P00, P01, NF0, F00, F01, K00, K01, 0>(R, T0, bits);
    if (NL > 1) sweep_layer<SEMI, M, NP1, P10, P11, NF1, F10, F11, K10, K11, N * NF0>(R, T1, bits);
#pragma unroll
    for (int j = 0; j < N; ++j) sw_sts(sw_xor3(outb, lo[j & 3], hi[(j >> 2) & 3]), R[j]);
#else
    if (NL > 1) {
      sweep_layer<SEMI, M, NP0, P00, P01, NF0, F00, F01, K00, K01, 0>(R, T0, bits);
      sweep_layer_store<SEMI, M, NP1, P10, P11, NF1, F10, F11, K10, K11, N * NF0>(R, T1, bits, outb, lo, hi);
    } else {
      sweep_layer_store<SEMI, M, NP0, P00, P01, NF0, F00, F01, K00, K01, 0>(R, T0, bits, outb, lo, hi);
    }
#endif
#else
#pragma unroll
    for (int j = 0; j < N; ++j) sw_sts(sw_xor3(outb, lo[j & 3], hi[(j >> 2) & 3]), R[j]);
#endif
    if (BPP) {
      if (IPW == 1) {
        bpt[(r0.w + it) * 32 + lane] = bits;
      } else {
        const int q = it % IPW;
        word |= bits << (BPP * q);
        if (q == IPW - 1 || it == r0.y - 1) {
          bpt[(r0.w + it / IPW) * 32 + lane] = word;
          word = 0;
        }
      }
    }
    base = base_n;
    inb = inb_n;
    sidx = sidx_n;
  }
}

// Deferred traceback of one shot (lane q walks shot q of its group backwards through the super-steps): per step the patch
// coordinates of the current entry, the packed back-pointers of its patch, the assignments they encode, the entry the winner came
// from.  A chain of dependent look-ups, so the records are read from the CTA's shared-memory copy when there is one (TBSM).
template <bool TBSM>
__device__ __forceinline__ void sw_traceback(const SweepDev &P, uint32_t tb_abs, const uint32_t *bp, int lane, const uint64_t (&syn)[4],
                                             int64_t myshot, int64_t B, uint64_t *__restrict__ corr) {
#define TBW(k) (TBSM ? sw_lds32(taddr + 4u * (uint32_t)(k)) : __ldg(t + (k)))
#define TBW4(k) (TBSM ? sw_lds128(taddr + 4u * (uint32_t)(k)) : __ldg(reinterpret_cast<const int4 *>(t + (k))))
  const int SG = 1 << P.sg, NE = 1 << P.W;
    const int f = lane >> P.sg, sub = lane & (SG - 1);
    const uint32_t *bpq = bp + (size_t)f * P.bp_words * 32;
    uint64_t cfg[4] = {0ull, 0ull, 0ull, 0ull};
    uint32_t x = (uint32_t)P.out_index[0] | ((uint32_t)sub << P.W);
    for (int i = P.n_ss - 1; i >= 0; --i) {
      const int32_t *t = P.tb + (size_t)i * SW_TB_INTS;
      const uint32_t taddr = tb_abs + (uint32_t)i * (SW_TB_INTS * 4);
      const int4 t0 = TBW4(0);          // M, layers, loop bits, bpp
      const int4 t1 = TBW4(4);      // wbase, ipw, closed, late (bit | patch bit << 16)
      const int4 tp = TBW4(8);      // positions of patch bits
      const int pos[4] = {tp.x, tp.y, tp.z, tp.w};
      uint32_t j = 0, ln = 0, it = 0, pm = 0;
#pragma unroll
      for (int b = 0; b < 4; ++b)
        if (b < t0.x) { j |= ((x >> pos[b]) & 1u) << b; pm |= 1u << pos[b]; }
#pragma unroll
      for (int q = 0; q < 5; ++q) ln |= ((x >> TBW(12 + q)) & 1u) << q;
      for (int q = 0; q < t0.z; ++q) it |= ((x >> TBW(17 + q)) & 1u) << q;
      uint32_t lsyn = 0;
      if (t1.w >= 0) {
        const int sb = t1.w & 0xffff, w = sb >> 6;
        const uint64_t sw = w == 0 ? syn[0] : (w == 1 ? syn[1] : (w == 2 ? syn[2] : syn[3]));
        lsyn = (uint32_t)((sw >> (sb & 63)) & 1ull);
        j ^= lsyn << (t1.w >> 16);
      }
      uint32_t pb = 0;
      if (t0.w) {
        const uint32_t wd = __ldcg(bpq + (size_t)(t1.x + it / t1.y) * 32 + ln);
        pb = wd >> (t0.w * (it % t1.y));
      }
      for (int li = t0.y - 1; li >= 0; --li) {
        const int lo = 30 + 14 * li;
        const int np = TBW(lo), nf = TBW(lo + 1), bo = TBW(lo + 2), flipm = TBW(lo + 11);
        const uint32_t k = nf ? ((pb >> (bo + j * nf)) & ((1u << nf) - 1u)) : 0u;
        uint32_t pflip = 0;
        for (int q = 0; q < np; ++q) {
          const int pbit = TBW(lo + 3 + 2 * q), v = TBW(lo + 4 + 2 * q);
          if ((j >> pbit) & 1u) pflip ^= (uint32_t)TBW(lo + 12 + q);
          const uint64_t bitv = (uint64_t)(((j >> pbit) & 1u) ^ (lsyn & ((uint32_t)flipm >> q) & 1u)) << (v & 63);
          const int w = v >> 6;
          cfg[0] |= w == 0 ? bitv : 0ull; cfg[1] |= w == 1 ? bitv : 0ull;
          cfg[2] |= w == 2 ? bitv : 0ull; cfg[3] |= w == 3 ? bitv : 0ull;
        }
        for (int q = 0; q < nf; ++q) {
          const int fm = TBW(lo + 7 + 2 * q), v = TBW(lo + 8 + 2 * q);
          const uint32_t kb = (k >> q) & 1u;
          const uint64_t bitv = (uint64_t)kb << (v & 63);
          const int w = v >> 6;
          cfg[0] |= w == 0 ? bitv : 0ull; cfg[1] |= w == 1 ? bitv : 0ull;
          cfg[2] |= w == 2 ? bitv : 0ull; cfg[3] |= w == 3 ? bitv : 0ull;
          if (kb) j ^= (uint32_t)fm;
        }
        j ^= pflip;
      }
      x &= ~pm;
#pragma unroll
      for (int b = 0; b < 4; ++b)
        if (b < t0.x) x |= ((j >> b) & 1u) << pos[b];
      for (int q = 0; q < t1.z; ++q) {
        const int sb = TBW(22 + 2 * q), ps = TBW(23 + 2 * q), w = sb >> 6;
        const uint64_t sw = w == 0 ? syn[0] : (w == 1 ? syn[1] : (w == 2 ? syn[2] : syn[3]));
        x ^= (uint32_t)((sw >> (sb & 63)) & 1ull) << ps;
      }
    }
    int hp = 0;
    for (int jb = 0; jb < P.nh; ++jb) {
      const int b = P.head_bits[jb], w = b >> 6;
      const uint64_t sw = w == 0 ? syn[0] : (w == 1 ? syn[1] : (w == 2 ? syn[2] : syn[3]));
      hp |= (int)((sw >> (b & 63)) & 1ull) << jb;
    }
    const uint64_t *hc = P.head_cfg + (((size_t)hp << P.W) + (x & (uint32_t)(NE - 1))) * P.ncw;
    if (myshot < B)
#pragma unroll
      for (int w = 0; w < 4; ++w)
        if (w < P.ncw) corr[myshot * P.ncw + w] = cfg[w] | __ldg(hc + w);
#undef TBW
#undef TBW4
}

// EXT = 1 additionally compiles the shapes with fresh pins (menu ids TQEC_SWEEP_MENU_BASE .. TQEC_SWEEP_MENU_MAXPLUS - 1: even-
// distance and rectangular codes); plans that do not use them run the EXT = 0 instantiation, whose code is unchanged
template <int SEMI, int MAXT, int EXT>
__global__ void __launch_bounds__(MAXT, 1)
k_sweep(const SweepDev P, const uint64_t *__restrict__ synd, const int64_t B, uint64_t *__restrict__ corr,
        double *__restrict__ out, int32_t *__restrict__ argmax_out, uint32_t *__restrict__ bp_all) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int NW = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int SG = 1 << P.sg, NE = 1 << P.W;
  if (((uint32_t)__cvta_generic_to_shared(smem_raw) + P.off_states) & 8191u) __trap();
  int32_t *sm_rec = reinterpret_cast<int32_t *>(smem_raw + P.off_rec);
  uint32_t *sm_lt = reinterpret_cast<uint32_t *>(smem_raw + P.off_lanetab);
  double *sm_tv = reinterpret_cast<double *>(smem_raw + P.off_tvals);
  for (int i = threadIdx.x; i < P.n_ss * SW_REC_INTS; i += blockDim.x) sm_rec[i] = P.rec[i];
  for (int i = threadIdx.x; i < P.n_ss * 32; i += blockDim.x) sm_lt[i] = P.lanetab[i];
  for (int i = threadIdx.x; i < P.n_tvals; i += blockDim.x) sm_tv[i] = P.tvals[i];
  // traceback records: the walk is a chain of dependent look-ups per super-step, so they sit next to the other tables
  int32_t *sm_tb = reinterpret_cast<int32_t *>(smem_raw + (P.off_tb >= 0 ? P.off_tb : 0));
  const uint32_t tb_abs = (uint32_t)__cvta_generic_to_shared(sm_tb);
  if (SEMI == TQEC_SEMIRING_MAXPLUS && P.off_tb >= 0)
    for (int i = threadIdx.x; i < P.n_ss * SW_TB_INTS; i += blockDim.x) sm_tb[i] = P.tb[i];
  __syncthreads();
  unsigned char *words = smem_raw + P.off_words + (size_t)P.words_bytes * warp;
  uint64_t *sh_syn = reinterpret_cast<uint64_t *>(words);
  uint16_t *stab = reinterpret_cast<uint16_t *>(sh_syn + SG * P.nsw);
  double *st = reinterpret_cast<double *>(smem_raw + P.off_states + (size_t)8192 * warp);
  const uint32_t st_abs = (uint32_t)__cvta_generic_to_shared(st);
  // head rows arrive by TMA when the table is stored pre-swizzled (W >= 8: the swizzle then stays inside one shot's row)
  const uint32_t bar = (uint32_t)__cvta_generic_to_shared(words + P.words_bytes - 16);
  uint32_t bar_phase = 0;
  if (P.head_tma && lane == 0) sw_mbar_init(bar);
  __syncwarp();
  const int G = P.grp;                                           // shots per deferred-traceback group (lane q < G walks shot q)
  const int NF = G >> P.sg;                                      // forward passes per group
  const int64_t n_groups = (B + G - 1) / G;
  const int sync_every = 32 / G;                                 // the CTA barrier stays at one per 32 shots
  const int64_t team = (int64_t)blockIdx.x * NW + warp, n_teams = (int64_t)gridDim.x * NW;
  uint32_t *bp = bp_all + (size_t)team * NF * P.bp_words * 32;

  // Every team of the CTA runs the same number of rounds and (optionally) meets the others at a CTA barrier before each
  // pass / group: passes have identical instruction streams, so teams that start together stay on the same code and
  // share instruction-cache lines; left alone they drift apart over a long launch (measured: -9 % on larger bodies).
  const int64_t rounds = (n_groups + n_teams - 1) / n_teams;
  for (int64_t rd = 0; rd < rounds; ++rd) {
    const int64_t g = team + rd * n_teams;
    const bool live = g < n_groups;
    if (P.sync_mode == 1 && rd % sync_every == 0) __syncthreads();
    const int64_t group0 = g * G, myshot = lane < G ? group0 + lane : B;
    uint64_t syn[4] = {0ull, 0ull, 0ull, 0ull};
    if (live && myshot < B)
#pragma unroll
      for (int w = 0; w < 4; ++w)
        if (w < P.nsw) syn[w] = synd[myshot * P.nsw + w];

    for (int f = 0; f < NF; ++f) {
      const int64_t shot0 = group0 + ((int64_t)f << P.sg);
      if (P.sync_mode == 2) __syncthreads();
      if (!live || shot0 >= B) continue;
      if ((lane >> P.sg) == f) {
        const int sub = lane & (SG - 1);
#pragma unroll
        for (int w = 0; w < 4; ++w)
          if (w < P.nsw) sh_syn[sub * P.nsw + w] = syn[w];
      }
      __syncwarp();
      // head: copy the tabulated state of the shots' head syndrome pattern
      if (P.head_tma) {
        // one bulk copy (TMA) per shot: rows are stored in the state's swizzled order, so a row lands as it is
        if (lane == 0) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the state was last written by ordinary stores
          sw_mbar_expect(bar, (uint32_t)(SG * NE * 8));
          for (int sub = 0; sub < SG; ++sub) {
            int hp = 0;
            for (int j = 0; j < P.nh; ++j) {
              const int b = P.head_bits[j];
              hp |= (int)((sh_syn[sub * P.nsw + (b >> 6)] >> (b & 63)) & 1ull) << j;
            }
            sw_bulk_g2s(st_abs + (uint32_t)((sub << P.W) * 8), P.head_state + ((size_t)hp << P.W), (uint32_t)(NE * 8), bar);
          }
        }
      } else {
        for (int sub = 0; sub < SG; ++sub) {
          int hp = 0;
          for (int j = 0; j < P.nh; ++j) {
            const int b = P.head_bits[j];
            hp |= (int)((sh_syn[sub * P.nsw + (b >> 6)] >> (b & 63)) & 1ull) << j;
          }
          const double *src = P.head_state + ((size_t)hp << P.W);
          for (int e = lane; e < NE; e += 32) st[sw_phys((uint32_t)(e | (sub << P.W)))] = __ldg(src + e);
        }
      }
      // closed-bit address masks of every (super-step, shot)
      for (int idx = lane; idx < (P.n_ss << P.sg); idx += 32) {
        const int i = idx >> P.sg, sub = idx & (SG - 1);
        const int32_t *r = sm_rec + i * SW_REC_INTS;
        uint32_t v = 0;
        for (int q = 0; q < r[14]; ++q) {
          const uint32_t c = (uint32_t)r[16 + q], sb = c & 0xffffu;
          if ((sh_syn[sub * P.nsw + (sb >> 6)] >> (sb & 63)) & 1ull) v ^= c >> 16;
        }
        stab[idx] = (uint16_t)v;
        uint32_t lv = 0;
        if (r[20] >= 0) {
          const uint32_t c = (uint32_t)r[20], sb = c & 0xffffu;
          if ((sh_syn[sub * P.nsw + (sb >> 6)] >> (sb & 63)) & 1ull) lv = ((c >> 16) & 0x3fffu) | ((c >> 30) & 1u ? 0x8000u : 0x4000u);
        }
        stab[(P.n_ss << P.sg) + idx] = (uint16_t)lv;
      }
      if (P.head_tma) {                                          // the copy overlapped the mask computation above
        sw_mbar_wait(bar, bar_phase);
        bar_phase ^= 1u;
      }
      __syncwarp();
      uint32_t *bpf = bp + (size_t)f * P.bp_words * 32;
      for (int i = 0; i < P.n_ss; ++i) {
        const int32_t *rec = sm_rec + i * SW_REC_INTS;
        const uint32_t lt = sm_lt[i * 32 + lane];
        const uint16_t *srow = stab + (i << P.sg), *lrow = stab + ((P.n_ss + i) << P.sg);
        switch (rec[0]) {
#define SW_CASE(ID, M, NL, NP0, P00, P01, NF0, F00, F01, K00, K01, NP1, P10, P11, NF1, F10, F11, K10, K11)               \
  case ID:                                                                                                             \
    if (ID < TQEC_SWEEP_MENU_BASE || (ID < TQEC_SWEEP_MENU_MAXPLUS ? EXT != 0                                              \
                                          : (SEMI == TQEC_SEMIRING_SUMPROD && (ID < TQEC_SWEEP_MENU_SP_BASE || EXT != 0))))  /* code size */ \
      sweep_step<SEMI, M, NL, NP0, P00, P01, NF0, F00, F01, K00, K01, NP1, P10, P11, NF1, F10, F11, K10, K11>(         \
          rec, sm_tv, st_abs, lt, srow, lrow, bpf, lane);                                                             \
    break;
          TQEC_SWEEP_MENU(SW_CASE)
#undef SW_CASE
          default: __trap();
        }
        __syncwarp();
      }
      if (SEMI == TQEC_SEMIRING_MAXPLUS) {
        if (out && lane < SG && shot0 + lane < B) out[shot0 + lane] = st[sw_phys((uint32_t)(P.out_index[0] | (lane << P.W)))];
      } else {
        // marginals over the open observable slots (observable 0 fastest) and their first maximal entry (findmax)
        const int NO = 1 << P.n_obs;
        for (int i = lane; i < (SG << P.n_obs); i += 32) {
          const int sub = i >> P.n_obs, idx = i & (NO - 1);
          if (shot0 + sub < B) out[(shot0 + sub) * NO + idx] = st[sw_phys((uint32_t)(P.out_index[idx] | (sub << P.W)))];
        }
        if (argmax_out && lane < SG && shot0 + lane < B) {
          double best = -1.0;
          int bi = 0;
          for (int idx = 0; idx < NO; ++idx) {
            const double v = st[sw_phys((uint32_t)(P.out_index[idx] | (lane << P.W)))];
            if (v > best) { best = v; bi = idx; }
          }
          argmax_out[shot0 + lane] = bi;
        }
      }
      __syncwarp();
    }

    // deferred traceback: lane q walks shot q of the group
    if (SEMI == TQEC_SEMIRING_MAXPLUS && live && lane < G) {
      if (P.off_tb >= 0) sw_traceback<true>(P, tb_abs, bp, lane, syn, myshot, B, corr);
      else sw_traceback<false>(P, 0u, bp, lane, syn, myshot, B, corr);
    }
    __syncwarp();
  }
}

template <int SEMI>
static const void *sweep_kernel(int maxt, int ext) {
  if (ext) return (const void *)k_sweep<SEMI, 512, 1>;          // (the register-budget variants exist for the base menu only)
  return maxt == 768 ? (const void *)k_sweep<SEMI, 768, 0> : maxt == 640 ? (const void *)k_sweep<SEMI, 640, 0> : (const void *)k_sweep<SEMI, 512, 0>;
}

static int launch_sweep_part(tqec_plan *plan, const SweepDev &P, const uint64_t *d_synd, int64_t B, uint64_t *d_corr, double *d_out,
                             int32_t *d_argmax, cudaStream_t stream) {
  const int64_t groups = (B + P.grp - 1) / P.grp;
  const int64_t ctas = (groups + plan->sw_teams - 1) / plan->sw_teams;
  const int grid = (int)(ctas < plan->sm_count ? ctas : plan->sm_count);
  const void *kern = plan->semiring == TQEC_SEMIRING_MAXPLUS ? sweep_kernel<TQEC_SEMIRING_MAXPLUS>(plan->sw_maxt, plan->sw_ext)
                                                             : sweep_kernel<TQEC_SEMIRING_SUMPROD>(plan->sw_maxt, plan->sw_ext);
  void *args[] = {(void *)&P, (void *)&d_synd, (void *)&B, (void *)&d_corr, (void *)&d_out, (void *)&d_argmax,
                  (void *)&plan->d_sw_bp};
  TQEC_CUDA(cudaLaunchKernel(kern, dim3(grid), dim3(32 * plan->sw_teams), args, (size_t)plan->sw_smem, stream));
  plan->launches += 1;
  return TQEC_OK;
}

// A launch runs whole rounds: every team takes one group of `grp` shots per round, so a batch that ends half-way through a
// round pays for the full round (1.25e6 shots per GPU = 16.5 rounds of 148 x 16 x 32 shots: 3 % of the step at 8 GPUs).
// The tail of a multi-round batch is therefore launched on its own with smaller groups (16 or 8 shots per team and
// round: less efficient per shot -- idle lanes in the traceback -- but the round is as short as the tail).  Groups only
// decide which team decodes which shots: results are unchanged.
int launch_sweep(tqec_plan *plan, const uint64_t *d_synd, int64_t B, uint64_t *d_corr, double *d_out, int32_t *d_argmax,
                 cudaStream_t stream) {
  const SweepDev &P = plan->sw;
  const int64_t per_round = (int64_t)plan->sm_count * plan->sw_teams * P.grp;
  const int64_t rounds = B / per_round, tail = B - rounds * per_round;
  int g_tail = P.grp;
  if (P.grp == 32 && tail > 0 && std::getenv("TQEC_SWEEP_NO_TAIL_SPLIT") == nullptr) {
    if (4 * tail <= per_round && (1 << P.sg) <= 8) g_tail = 8;
    else if (2 * tail <= per_round && (1 << P.sg) <= 16) g_tail = 16;
  }
  if (g_tail == P.grp) return launch_sweep_part(plan, P, d_synd, B, d_corr, d_out, d_argmax, stream);
  if (rounds == 0) {                                             // less than a round in all (the last chunk of a pipeline, small
    SweepDev T = P;                                              // batches): one launch with the smaller groups, every SM busy
    T.grp = g_tail;
    return launch_sweep_part(plan, T, d_synd, B, d_corr, d_out, d_argmax, stream);
  }
  const int64_t main_shots = rounds * per_round;
  const int64_t NO = plan->semiring == TQEC_SEMIRING_MAXPLUS ? 1 : ((int64_t)1 << P.n_obs);
  int rc = launch_sweep_part(plan, P, d_synd, main_shots, d_corr, d_out, d_argmax, stream);
  if (rc) return rc;
  SweepDev T = P;
  T.grp = g_tail;
  return launch_sweep_part(plan, T, d_synd + main_shots * P.nsw, tail, d_corr ? d_corr + main_shots * P.ncw : nullptr,
                           d_out ? d_out + main_shots * NO : nullptr, d_argmax ? d_argmax + main_shots : nullptr, stream);
}

template <typename T>
static int sw_upload(void **slot, const T *src, size_t n) {
  TQEC_CUDA(cudaMalloc(slot, (n ? n : 1) * sizeof(T)));
  if (n) TQEC_CUDA(cudaMemcpy(*slot, src, n * sizeof(T), cudaMemcpyHostToDevice));
  return TQEC_OK;
}

void sweep_destroy(tqec_plan *p) {
  for (int i = 0; i < 8; ++i) if (p->d_sw[i]) cudaFree(p->d_sw[i]);
  if (p->d_sw_bp) cudaFree(p->d_sw_bp);
}

// Validate and upload the sweep tables of a plan descriptor; chooses teams per CTA from shared memory and registers.
int sweep_create(tqec_plan *p, const tqec_plan_desc *d, const cudaDeviceProp &prop) {
  const tqec_sweep_desc *s = d->sweep;
  p->has_sweep = 0;
  if (!s || std::getenv("TQEC_NO_SWEEP")) return TQEC_OK;
  TQEC_REQUIRE(s->W >= 1 && s->sg >= 0 && s->sg <= 5 && s->W + s->sg == 10, "sweep: W=%d sg=%d must add up to 10 index bits", s->W, s->sg);
  TQEC_REQUIRE(s->n_ss > 0 && s->rec && s->tb && s->lanetab && s->tvals && s->head_state && s->head_cfg && s->out_index,
               "sweep: missing table");
  TQEC_REQUIRE(s->n_head_bits >= 0 && s->n_head_bits <= 16 && (s->n_head_bits == 0 || s->head_bits), "sweep: bad head bits");
  const int nsw = words_for(d->n_checks), ncw = words_for(d->n_vars);
  TQEC_REQUIRE(nsw <= 4 && ncw <= 4, "sweep: more than 256 checks / variables");
  for (int i = 0; i < s->n_ss; ++i) {
    const int32_t *r = s->rec + (size_t)i * SW_REC_INTS;
    TQEC_REQUIRE(r[0] >= 0 && r[0] < (d->semiring == TQEC_SEMIRING_MAXPLUS ? TQEC_SWEEP_MENU_MAXPLUS : TQEC_SWEEP_MENU_SIZE),
                 "sweep step %d: unknown shape %d", i, r[0]);
    TQEC_REQUIRE(r[1] >= 1 && r[1] <= 8, "sweep step %d: bad iteration count %d", i, r[1]);
    TQEC_REQUIRE(r[2] >= 0 && r[2] + 8 <= s->n_tvals + 8 && r[14] >= 0 && r[14] <= 4, "sweep step %d: bad offsets", i);
    for (int q = 0; q < r[14]; ++q)
      TQEC_REQUIRE(((uint32_t)r[16 + q] & 0xffffu) < (uint32_t)d->n_checks, "sweep step %d: syndrome bit out of range", i);
  }
  for (int j = 0; j < s->n_head_bits; ++j)
    TQEC_REQUIRE(s->head_bits[j] >= 0 && s->head_bits[j] < d->n_checks, "sweep: head bit out of range");

  SweepDev &D = p->sw;
  std::memset(&D, 0, sizeof(D));
  D.n_ss = s->n_ss; D.W = s->W; D.sg = s->sg; D.nh = s->n_head_bits; D.nsw = nsw; D.ncw = ncw;
  D.bp_words = s->bp_words > 0 ? s->bp_words : 1;
  D.sync_mode = 1;
  if (const char *e = std::getenv("TQEC_SWEEP_SYNC")) { const int v = std::atoi(e); if (v >= 0 && v <= 2) D.sync_mode = v; } D.n_tvals = s->n_tvals; D.n_obs = d->n_obs;
  TQEC_REQUIRE(d->n_obs <= 4, "sweep: more than 4 open observables");
  for (int i = 0; i < (1 << d->n_obs); ++i) D.out_index[i] = s->out_index[i];
  for (int j = 0; j < s->n_head_bits; ++j) D.head_bits[j] = s->head_bits[j];
  const size_t nhp = (size_t)1 << s->n_head_bits, ne = (size_t)1 << s->W;
  int rc;
  if ((rc = sw_upload(&p->d_sw[0], s->rec, (size_t)s->n_ss * SW_REC_INTS))) return rc;
  if ((rc = sw_upload(&p->d_sw[1], s->tb, (size_t)s->n_ss * SW_TB_INTS))) return rc;
  if ((rc = sw_upload(&p->d_sw[2], s->lanetab, (size_t)s->n_ss * 32))) return rc;
  if ((rc = sw_upload(&p->d_sw[3], s->tvals, (size_t)s->n_tvals))) return rc;
  // W >= 8: the swizzle (index bits 4..7 into bits 0..3) stays inside one shot's row, so the rows are stored in the
  // state's swizzled order and a row is fetched by one bulk copy (TMA); TQEC_SWEEP_NO_TMA=1 keeps the load / store loop
  const bool head_tma = s->W >= 8 && std::getenv("TQEC_SWEEP_NO_TMA") == nullptr;
  if (head_tma) {
    std::vector<double> sw(nhp * ne);
    for (size_t h = 0; h < nhp; ++h)
      for (size_t e = 0; e < ne; ++e) sw[h * ne + sw_phys((uint32_t)e)] = s->head_state[h * ne + e];
    if ((rc = sw_upload(&p->d_sw[4], sw.data(), nhp * ne))) return rc;
  } else if ((rc = sw_upload(&p->d_sw[4], s->head_state, nhp * ne))) return rc;
  D.head_tma = head_tma ? 1 : 0;
  // shots per deferred-traceback group: 32 keeps every lane busy in the traceback; a smaller group shrinks the team's
  // back-pointer ring (groups of 8: 87 MB at d = 9, L2 resident) at the price of idle lanes in the traceback
  D.grp = 32;
  if (const char *e = std::getenv("TQEC_SWEEP_GROUP")) { const int v = std::atoi(e); if ((v == 32 || v == 16 || v == 8 || v == 4 || v == 2) && v >= (1 << s->sg)) D.grp = v; }
  if ((rc = sw_upload(&p->d_sw[5], s->head_cfg, d->semiring == TQEC_SEMIRING_MAXPLUS ? nhp * ne * ncw : (size_t)1))) return rc;
  D.rec = (const int32_t *)p->d_sw[0]; D.tb = (const int32_t *)p->d_sw[1]; D.lanetab = (const uint32_t *)p->d_sw[2];
  D.tvals = (const double *)p->d_sw[3]; D.head_state = (const double *)p->d_sw[4]; D.head_cfg = (const uint64_t *)p->d_sw[5];

  // shared-memory layout: [front gap up to the first 8 KiB-aligned absolute address][states][remaining blobs]
  int reserved = 1024;
  cudaDeviceGetAttribute(&reserved, cudaDevAttrReservedSharedMemoryPerBlock, d->device);
  const size_t gap = (8192 - (size_t)reserved % 8192) % 8192;
  const size_t rec_b = ((size_t)s->n_ss * SW_REC_INTS * 4 + 15) & ~(size_t)15, lt_b = (size_t)s->n_ss * 128;
  const size_t tv_b = ((size_t)s->n_tvals * 8 + 15) & ~(size_t)15;
  const size_t tb_b = d->semiring == TQEC_SEMIRING_MAXPLUS && std::getenv("TQEC_SWEEP_TB_GLOBAL") == nullptr ? (size_t)s->n_ss * SW_TB_INTS * 4 : 0;
  const size_t words_b = ((((size_t)nsw << s->sg) * 8 + ((size_t)s->n_ss << s->sg) * 4 + 15) & ~(size_t)15) + 16;   // syndromes, early + late masks, mbarrier
  // register budget variant: 512 threads (128 registers), 640 (96) or 768 (80); TQEC_SWEEP_MAXT overrides the default
  int maxt = 512;
  if (const char *e = std::getenv("TQEC_SWEEP_MAXT")) { const int v = std::atoi(e); if (v == 768 || v == 640 || v == 512) maxt = v; }
  int ext = 0;
  for (int i = 0; i < s->n_ss; ++i) {
    const int id = s->rec[(size_t)i * SW_REC_INTS];
    if ((id >= TQEC_SWEEP_MENU_BASE && id < TQEC_SWEEP_MENU_MAXPLUS) || id >= TQEC_SWEEP_MENU_SP_BASE) ext = 1;
  }
  if (ext) maxt = 512;
  const void *kern = d->semiring == TQEC_SEMIRING_MAXPLUS ? sweep_kernel<TQEC_SEMIRING_MAXPLUS>(maxt, ext)
                                                          : sweep_kernel<TQEC_SEMIRING_SUMPROD>(maxt, ext);
  p->sw_maxt = maxt;
  p->sw_ext = ext;
  cudaFuncAttributes fa;
  TQEC_CUDA(cudaFuncGetAttributes(&fa, kern));
  int cap = fa.maxThreadsPerBlock / 32;
  if (fa.numRegs > 0) {
    const int by_regs = prop.regsPerBlock / (((fa.numRegs + 7) & ~7) * 32);
    if (by_regs < cap) cap = by_regs;
  }
  if (const char *e = std::getenv("TQEC_SWEEP_TEAMS")) { const int v = std::atoi(e); if (v >= 1 && v < cap) cap = v; }
  const size_t budget = (size_t)prop.sharedMemPerBlockOptin;
  int nw = 0;
  size_t offs[5] = {0, 0, 0, 0, 0}, total = 0;
  for (int cand = cap; cand >= 1 && nw == 0; --cand) {
    size_t front = 0, tail = gap + (size_t)8192 * cand;
    const size_t sizes[5] = {rec_b, lt_b, tv_b, words_b * cand, tb_b};
    size_t o[5];
    for (int i = 0; i < 5; ++i) {
      if (front + sizes[i] <= gap) { o[i] = front; front += sizes[i]; }
      else { o[i] = tail; tail += sizes[i]; }
    }
    if (tail <= budget) { nw = cand; total = tail; for (int i = 0; i < 5; ++i) offs[i] = o[i]; }
  }
  if (nw < 1) { sweep_destroy(p); std::memset(p->d_sw, 0, sizeof(p->d_sw)); return TQEC_OK; }   // does not fit: general kernels
  D.off_states = (int32_t)gap; D.off_rec = (int32_t)offs[0]; D.off_lanetab = (int32_t)offs[1]; D.off_tvals = (int32_t)offs[2];
  D.off_words = (int32_t)offs[3]; D.words_bytes = (int32_t)words_b;
  D.off_tb = tb_b ? (int32_t)offs[4] : -1;
  p->sw_teams = nw; p->sw_smem = (int)total;
  TQEC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)total));
  const size_t bp_bytes = (size_t)p->sm_count * nw * (D.grp >> s->sg) * D.bp_words * 32 * sizeof(uint32_t);
  cudaError_t e = cudaMalloc((void **)&p->d_sw_bp, bp_bytes);
  if (e != cudaSuccess) { set_error("cudaMalloc(%zu B sweep back-pointer scratch): %s", bp_bytes, cudaGetErrorString(e)); return TQEC_ERR_NOMEM; }
  p->has_sweep = 1;
  return TQEC_OK;
}

}  // namespace tqec
