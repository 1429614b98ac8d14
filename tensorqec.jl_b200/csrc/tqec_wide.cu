// Global-memory executor (k_wide_pass): sum-product frontier plans whose state does not fit on chip (14..31 bits).
// Host side of the lowering and the full description: tensorqec.jl_b200/wide.py.  Per batch of shots the state lives in
// two HBM arrays (entry sigma of shot b at (b << w_cap) | sigma, ping-pong).  One launch per PASS: for every value of
// the pass's spectator bits and every shot, a CTA
//   1. gathers the 2^t_in tile entries (bit deposit of the local index at the tile positions; the lowest index bits are
//      always tile bits, so loads come in contiguous runs of 128 B),
//   2. runs the pass's steps on the tile in shared memory (ping-pong; the generic gather of the frontier recurrence in
//      tile-local coordinates: full = deposit(tau) | closed syndrome bits, candidates = coset of the opened pattern),
//   3. scatters the 2^t_out entries of the result to the other array.
// HBM traffic = one read + one write of the state per pass (instead of per step): 16 * 2^w bytes per pass and shot.
#include <cstdlib>
#include <cstring>

#include "tqec_common.h"

namespace tqec {

#define WD_MAX_STEPS 256

struct WidePassArgs {
  const double *gin;
  double *gout;
  const uint64_t *synd;      // syndromes of the batch's first shot
  int64_t nb;                // shots in the batch
  const int32_t *pass;       // this pass's header
  const int32_t *step_hdr, *ints;
  const double *tables;
  int32_t nsw, w_cap, t_max;
  // dynamic per-shot rescaling (optional): largest state entry of every shot after the previous pass (bit pattern of a
  // non-negative double), the same for this pass (filled with atomicMax), and the shots' accumulated exponents
  const unsigned long long *max_in;
  unsigned long long *max_out;
  int32_t *exps;
};

// deposit the low bits of x at the set bits of mask (software pdep; mask has at most 31 bits)
__host__ __device__ __forceinline__ uint32_t wd_pdep(uint32_t x, uint32_t mask) {
  uint32_t out = 0;
  while (mask) {
    const uint32_t low = mask & (0u - mask);
    if (x & 1u) out |= low;
    x >>= 1;
    mask ^= low;
  }
  return out;
}

// one local step on the tile: Sout[tau] = sum_k Sin[(full & inmask) ^ ML[pat] ^ MK[k]] * T[pat * nk + k]
template <int NK>
__device__ __forceinline__ void wd_step(const int32_t *__restrict__ q, const int32_t *__restrict__ I, const double *__restrict__ T,
                                        const uint32_t *__restrict__ sc, uint32_t cb, const double *__restrict__ Sin,
                                        double *__restrict__ Sout, int tid, int NT) {
  const int w_in = q[TQEC_WL_WIN], n_open = q[TQEC_WL_NOPEN], w_out = q[TQEC_WL_WOUT];
  const int nk = NK > 0 ? NK : q[TQEC_WL_NK];
  const int32_t *ML = I + q[TQEC_WL_OFF_ML], *MK = I + q[TQEC_WL_OFF_MK];
  const double *Tt = T + q[TQEC_WL_OFF_T];
  const uint32_t inmask = (1u << w_in) - 1u;
  const int n = 1 << w_out;
  if (n_open == 0) {
    // nothing opened: one coset for every output, masks and factor values live in registers
    const uint32_t m0 = (uint32_t)ML[0] ^ (uint32_t)MK[0], m1 = NK == 2 ? ((uint32_t)ML[0] ^ (uint32_t)MK[1]) : 0u;
    const double t0 = Tt[0], t1 = NK == 2 ? Tt[1] : 0.0;
    if (NK == 1 || NK == 2) {
      for (int tau = tid; tau < n; tau += NT) {
        const uint32_t low = (sc[tau & 63] | sc[64 + (tau >> 6)] | cb) & inmask;
        double acc = Sin[low ^ m0] * t0;
        if (NK == 2) acc += Sin[low ^ m1] * t1;
        Sout[tau] = acc;
      }
      return;
    }
  }
  for (int tau = tid; tau < n; tau += NT) {
    const uint32_t full = sc[tau & 63] | sc[64 + (tau >> 6)] | cb;
    const uint32_t pat = full >> w_in;
    const uint32_t low = (full & inmask) ^ (uint32_t)ML[pat];
    const double *tb = Tt + pat * nk;
    double acc = Sin[low ^ (uint32_t)MK[0]] * tb[0];
    for (int k = 1; k < nk; ++k) acc += Sin[low ^ (uint32_t)MK[k]] * tb[k];
    Sout[tau] = acc;
  }
}

// A PURE step (nothing opened, nothing closed, two candidates: most steps of a detector error model -- one mechanism
// with prior [1 - p, p] flipping the checks in `m`) is a butterfly over the pairs (tau, tau ^ m): both entries are read
// once, both results written back IN PLACE (no ping-pong, no scatter table, half the shared-memory traffic of the
// generic gather).  Same operations in the same order as wd_step<2>: out = fma(S[tau ^ m], t1, S[tau] * t0).
__device__ __forceinline__ void wd_pure_step(const int32_t *__restrict__ q, const int32_t *__restrict__ I, const double *__restrict__ T,
                                             double *__restrict__ S, int tid, int NT) {
  const uint32_t m = (uint32_t)I[q[TQEC_WL_OFF_MK] + 1];
  const double t0 = T[q[TQEC_WL_OFF_T]], t1 = T[q[TQEC_WL_OFF_T] + 1];
  const int p = 31 - __clz(m);                                   // pivot: the pair member with this bit clear comes first
  const uint32_t lowmask = (1u << p) - 1u;
  const int n_pairs = 1 << (q[TQEC_WL_WOUT] - 1);
  for (int i = tid; i < n_pairs; i += NT) {
    const uint32_t tau = (((uint32_t)i & ~lowmask) << 1) | ((uint32_t)i & lowmask);
    const double a = S[tau], b = S[tau ^ m];
    S[tau] = fma(b, t1, a * t0);
    S[tau ^ m] = fma(a, t1, b * t0);
  }
}

// Up to three consecutive pure steps whose masks are linearly independent are FUSED: a thread loads the 2^G entries of
// one orbit of span{m_0..m_{G-1}} into registers, applies the G butterflies there (step j pairs the orbit members that
// differ in generator j; same operations in the same order as G separate steps) and stores the orbit back: one shared-
// memory round trip and one barrier for G steps.  `pivots`: the leading bits of the masks' echelon form; the orbit
// representative is the member with all pivot bits clear.
template <int G>
__device__ __forceinline__ void wd_pure_group(const int32_t *__restrict__ q, const int32_t *__restrict__ I, const double *__restrict__ T,
                                              double *__restrict__ S, uint32_t pivots, int tid, int NT) {
  uint32_t m[G], cm[1 << G];
  double t0[G], t1[G];
  int piv[G];
#pragma unroll
  for (int j = 0; j < G; ++j) {
    const int32_t *qj = q + j * TQEC_WIDE_STEP_INTS;
    m[j] = (uint32_t)I[qj[TQEC_WL_OFF_MK] + 1];
    t0[j] = T[qj[TQEC_WL_OFF_T]];
    t1[j] = T[qj[TQEC_WL_OFF_T] + 1];
    piv[j] = __ffs(pivots) - 1;                                  // ascending
    pivots &= pivots - 1;
  }
#pragma unroll
  for (int k = 0; k < (1 << G); ++k) {
    uint32_t x = 0;
#pragma unroll
    for (int j = 0; j < G; ++j)
      if ((k >> j) & 1) x ^= m[j];
    cm[k] = x;
  }
  const int n_orb = 1 << (q[TQEC_WL_WOUT] - G);
  for (int i = tid; i < n_orb; i += NT) {
    uint32_t tau = (uint32_t)i;
#pragma unroll
    for (int j = 0; j < G; ++j) {
      const uint32_t lowmask = (1u << piv[j]) - 1u;
      tau = ((tau & ~lowmask) << 1) | (tau & lowmask);
    }
    double v[1 << G];
#pragma unroll
    for (int k = 0; k < (1 << G); ++k) v[k] = S[tau ^ cm[k]];
#pragma unroll
    for (int j = 0; j < G; ++j)
#pragma unroll
      for (int k = 0; k < (1 << G); ++k)
        if (!((k >> j) & 1)) {
          const double a = v[k], b = v[k | (1 << j)];
          v[k] = fma(b, t1[j], a * t0[j]);
          v[k | (1 << j)] = fma(a, t1[j], b * t0[j]);
        }
#pragma unroll
    for (int k = 0; k < (1 << G); ++k) S[tau ^ cm[k]] = v[k];
  }
}

__device__ __forceinline__ bool wd_is_pure(const int32_t *q, const int32_t *I) {
  return q[TQEC_WL_NK] == 2 && q[TQEC_WL_NOPEN] == 0 && q[TQEC_WL_NCLOSE] == 0 && I[q[TQEC_WL_OFF_MK]] == 0 &&
         I[q[TQEC_WL_OFF_ML]] == 0 && I[q[TQEC_WL_OFF_MK] + 1] != 0;
}

template <int NT>
__global__ void __launch_bounds__(NT) k_wide_pass(const WidePassArgs A) {
  extern __shared__ __align__(16) unsigned char wd_smem[];
  const int tid = threadIdx.x;
  const int32_t *ph = A.pass;
  const int w_in = ph[TQEC_WP_WIN], w_out = ph[TQEC_WP_WOUT], t_in = ph[TQEC_WP_TIN], t_out = ph[TQEC_WP_TOUT];
  const int ns = ph[TQEC_WP_NSTEPS], s0 = ph[TQEC_WP_STEP0];
  const uint32_t tin = (uint32_t)ph[TQEC_WP_TINMASK], tout = (uint32_t)ph[TQEC_WP_TOUTMASK];
  const int n_ints = ph[TQEC_WP_N_INTS], n_tab = ph[TQEC_WP_N_TAB];
  double *S0 = reinterpret_cast<double *>(wd_smem);
  double *S1 = S0 + ((size_t)1 << A.t_max);
  double *sT = S1 + ((size_t)1 << A.t_max);
  int32_t *sI = reinterpret_cast<int32_t *>(sT + ((n_tab + 1) & ~1));
  int32_t *sQ = sI + ((n_ints + 3) & ~3);                        // step headers of the pass
  uint32_t *dep = reinterpret_cast<uint32_t *>(sQ + ns * TQEC_WIDE_STEP_INTS);   // [0..127] input, [128..255] output deposits
  uint32_t *sc = dep + 256;                                      // two scatter tables of 128 entries (double-buffered)
  uint32_t *cbs = sc + 256;                                      // closed-bit value of every step for the current shot
  uint32_t *grp = cbs + WD_MAX_STEPS;                            // per step: 0 = generic, (pivot mask << 2) | G = head of a fused pure group
  for (int i = tid; i < n_tab; i += NT) sT[i] = A.tables[ph[TQEC_WP_OFF_TAB] + i];
  for (int i = tid; i < n_ints; i += NT) sI[i] = A.ints[ph[TQEC_WP_OFF_INTS] + i];
  for (int i = tid; i < ns * TQEC_WIDE_STEP_INTS; i += NT) sQ[i] = A.step_hdr[(size_t)s0 * TQEC_WIDE_STEP_INTS + i];
  if (tid < 128) {
    const uint32_t x = tid < 64 ? (uint32_t)tid : ((uint32_t)(tid - 64) << 6);
    dep[tid] = wd_pdep(x, tin);
    dep[128 + tid] = wd_pdep(x, tout);
  }
  __syncthreads();
  if (tid == 0) {
    // group consecutive pure steps (up to three, masks linearly independent): echelon form by leading bits
    int s = 0;
    while (s < ns) {
      grp[s] = 0;
      if (!wd_is_pure(sQ + s * TQEC_WIDE_STEP_INTS, sI)) { ++s; continue; }
      uint32_t b[3], pm = 0;
      int g = 0;
      while (g < 3 && s + g < ns && wd_is_pure(sQ + (s + g) * TQEC_WIDE_STEP_INTS, sI)) {
        uint32_t x = (uint32_t)sI[sQ[(s + g) * TQEC_WIDE_STEP_INTS + TQEC_WL_OFF_MK] + 1];
        for (int pass2 = 0; pass2 < 2; ++pass2)
          for (int j = 0; j < g; ++j)
            if ((x >> (31 - __clz(b[j]))) & 1u) x ^= b[j];
        if (x == 0) break;                                       // dependent on the masks already in the group
        b[g] = x;
        pm |= 1u << (31 - __clz(x));
        ++g;
      }
      grp[s] = (pm << 2) | (uint32_t)g;
      for (int j = 1; j < g; ++j) grp[s + j] = 0;
      s += g;
    }
  }
  __syncthreads();
  const int n_spec = w_in - t_in;
  const uint32_t spec_in = (w_in >= 32 ? 0xffffffffu : ((1u << w_in) - 1u)) & ~tin;
  const uint32_t spec_out = (w_out >= 32 ? 0xffffffffu : ((1u << w_out) - 1u)) & ~tout;
  const int64_t n_tiles = A.nb << n_spec;
  const int n_in = 1 << t_in, n_out = 1 << t_out;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t b = tile >> n_spec;
    const uint32_t sp = (uint32_t)(tile & (((int64_t)1 << n_spec) - 1));
    const double *gi = A.gin + ((size_t)b << A.w_cap) + wd_pdep(sp, spec_in);
    double *go = A.gout + ((size_t)b << A.w_cap) + wd_pdep(sp, spec_out);
    // Dynamic rescaling: when the shot's largest entry has fallen below 2^-300, every tile of the shot multiplies what it
    // loads by the same power of two (exact) and the shot's exponent absorbs it -- all tiles read the same maximum, so
    // they take the same decision; tile 0 of the shot books it.
    double scale = 1.0;
    if (A.max_in) {
      const double mx = __longlong_as_double((long long)A.max_in[b]);
      if (mx > 0.0 && mx < 4.909093465297727e-91) {              // 2^-300
        const int k = -ilogb(mx);
        scale = ldexp(1.0, k);
        if (sp == 0 && tid == 0) A.exps[b] -= k;
      }
    }
    if (tid < ns) {
      const int32_t *q = sQ + tid * TQEC_WIDE_STEP_INTS;
      const int32_t *CL = sI + q[TQEC_WL_OFF_CLOSE];
      const uint64_t *syn = A.synd + (size_t)b * A.nsw;
      uint32_t v = 0;
      for (int c = 0; c < q[TQEC_WL_NCLOSE]; ++c) {
        const int sb = CL[2 * c + 1];
        v |= (uint32_t)((__ldg(syn + (sb >> 6)) >> (sb & 63)) & 1ull) << CL[2 * c];
      }
      cbs[tid] = v;
    }
    if (tid >= 128 && tid < 256) {                               // scatter table of the first step
      const int i = tid - 128;
      const uint32_t x = i < 64 ? (uint32_t)i : ((uint32_t)(i - 64) << 6);
      sc[i] = wd_pdep(x, (uint32_t)sQ[TQEC_WL_KEEPMASK]);
    }
    // (prefetching the next tile into registers was tried: 126 registers, two CTAs per SM instead of three, 20 % slower)
    for (int l = tid; l < n_in; l += NT) S0[l] = __ldcs(gi + (dep[l & 63] | dep[64 + (l >> 6)])) * scale;
    __syncthreads();
    double *Sin = S0, *Sout = S1;
    for (int s = 0; s < ns;) {
      const int32_t *q = sQ + s * TQEC_WIDE_STEP_INTS;
      const uint32_t *scs = sc + ((s & 1) << 7);
      const int g = (int)(grp[s] & 3u);
      const int nxt = s + (g ? g : 1);
      if (nxt < ns && tid < 128) {                               // scatter table of the next step to run (other buffer)
        const uint32_t x = tid < 64 ? (uint32_t)tid : ((uint32_t)(tid - 64) << 6);
        sc[((nxt & 1) << 7) + tid] = wd_pdep(x, (uint32_t)sQ[nxt * TQEC_WIDE_STEP_INTS + TQEC_WL_KEEPMASK]);
      }
      if (g) {
        const uint32_t pm = grp[s] >> 2;
        if (g == 3) wd_pure_group<3>(q, sI, sT, Sin, pm, tid, NT);
        else if (g == 2) wd_pure_group<2>(q, sI, sT, Sin, pm, tid, NT);
        else wd_pure_step(q, sI, sT, Sin, tid, NT);
        __syncthreads();
        s = nxt;
        continue;
      }
      const int nk = q[TQEC_WL_NK];
      if (nk == 1) wd_step<1>(q, sI, sT, scs, cbs[s], Sin, Sout, tid, NT);
      else if (nk == 2) wd_step<2>(q, sI, sT, scs, cbs[s], Sin, Sout, tid, NT);
      else wd_step<0>(q, sI, sT, scs, cbs[s], Sin, Sout, tid, NT);
      __syncthreads();
      double *tmp = Sin; Sin = Sout; Sout = tmp;
      s = nxt;
    }
    double tmax = 0.0;
    for (int l = tid; l < n_out; l += NT) {
      const double v = Sin[l];
      __stcs(go + (dep[128 + (l & 63)] | dep[192 + (l >> 6)]), v);
      tmax = fmax(tmax, v);
    }
    if (A.max_out) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) tmax = fmax(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
      if ((tid & 31) == 0 && tmax > 0.0) atomicMax(A.max_out + b, (unsigned long long)__double_as_longlong(tmax));
    }
    __syncthreads();
  }
}

// ---- butterfly passes (k_wide_bf) ------------------------------------------------------------------------------------------
// Passes made of rank-1 factors only (the mechanisms of a detector error model) run here instead of k_wide_pass; the
// encoding is produced by tqec_lower_wide.cpp:bf_encode_pass (read its header first).  Per tile:
//   * the 2^t_in input entries arrive by cp.async (16 bytes per request, issued one tile ahead into the other half of a
//     double buffer) at index = their tile-local index (input bit i sits at position i), swizzled by bf_phys; the rest of
//     the 2^12 entries is zero-filled;
//   * per GROUP a thread takes one coset of the group's TQEC_BF_G-dimensional span into 2^G registers (Gray-code walks:
//     one XOR per address), zeroes the members the group record names (a reused position reopens), applies the G unit
//     steps  v[k] += r v[k ^ (1 << j)]  (straight-line code, one FP64 FMA per entry and step; r = 0 for a padding
//     dimension) and the dependent steps (switch on the coordinate vector), and writes the coset back: one shared-memory
//     round trip and one barrier per group of typically five steps;
//   * closed checks are not compacted away: the running mask XP (closed position x syndrome bit, in swizzled byte
//     offsets) is XOR-ed into every address, so the live half of the tile is the one the addresses reach;
//   * the 2^t_out surviving entries are gathered through the position of every output bit and stored.
// All address components are XOR-linear byte offsets into the tile.
#define BF_NT 128
#define BF_NE (1 << TQEC_BF_G)
#define BF_MAX_GROUPS 32
#define BF_MAX_STEPS 96    /* dependent steps of a pass */
#define BF_MAX_CLOSES 64
#define BF_TILE_BYTES 32768

struct WideBfArgs {
  const double *gin;
  double *gout;
  const uint64_t *synd;
  int64_t nb;
  const int32_t *pass;       // the pass's header in pass_hdr (widths, tile masks)
  const int32_t *bf;         // the pass's butterfly block
  const double *vals;        // bf_vals
  int32_t nsw, w_cap;
  const unsigned long long *max_in;
  unsigned long long *max_out;
  int32_t *exps;
};

struct BfGroupS {
  double r[TQEC_BF_G];               // ratios of the unit steps
  uint32_t zmask;                    // coset members zeroed at the load
  uint16_t orb_lo[16], orb_hi[16];   // swizzled byte offset of the coset representative: lo[i & 15] ^ hi[i >> 4]
  uint16_t bP[8];                    // swizzled byte offsets of the basis vectors
  int16_t n_free, dep0, n_dep, close0, n_closes, pad;
};

// swizzle of the tile index: bits 5..7 and 8..10 are folded into bits 1..3 (bit 0 stays: entries travel in 16-byte pairs)
__host__ __device__ __forceinline__ uint32_t bf_phys(uint32_t x) { return TQEC_BF_SWZ_WIDE ? (x ^ ((x >> 4) & 14u) ^ ((x >> 7) & 14u)) : (x ^ ((x >> 4) & 14u)); }
__host__ __device__ constexpr int bf_ctz(int k) { return (k & 1) ? 0 : (k & 2) ? 1 : (k & 4) ? 2 : (k & 8) ? 3 : (k & 16) ? 4 : 5; }

template <int C>
__device__ __forceinline__ void bf_pair(double (&v)[BF_NE], const double r) {
  constexpr int LOW = C & -C;
#pragma unroll
  for (int k = 0; k < BF_NE; ++k)
    if (!(k & LOW)) {
      const double a = v[k], b = v[k ^ C];
      v[k] = fma(r, b, a);
      v[k ^ C] = fma(r, a, b);
    }
}

__device__ __forceinline__ void bf_cp_async16(uint32_t dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

__global__ void __launch_bounds__(BF_NT, 3) k_wide_bf(const WideBfArgs A) {
  extern __shared__ __align__(16) unsigned char bf_dyn[];      // two tiles of 2^12 doubles
  __shared__ uint32_t dep_in[128];
  __shared__ uint32_t sdep_in[288], sdep_out[288];         // deposit of the tile number's spectator part: 7 + 7 + 5 bits
  // the last group stores straight to the state: offsets and logical tile indices of its coset representatives and basis
  __shared__ uint32_t gout_lo[16], gout_hi[16], gb_out[8];
  __shared__ uint16_t elog_lo[16], elog_hi[16], b_log[8];
  __shared__ uint32_t s_cmask;
  __shared__ uint32_t s_xp[BF_MAX_GROUPS + 1];               // address mask of the closed checks before every group (per tile)
  __shared__ BfGroupS grp[BF_MAX_GROUPS];
  __shared__ double s_r[BF_MAX_STEPS];
  __shared__ int32_t s_code[BF_MAX_STEPS];
  __shared__ uint16_t s_closeP[BF_MAX_CLOSES], s_closeB[BF_MAX_CLOSES];
  __shared__ uint8_t s_closePos[BF_MAX_CLOSES];
  const int tid = threadIdx.x;
  const int32_t *ph = A.pass, *bf = A.bf;
  const int w_in = ph[TQEC_WP_WIN], w_out = ph[TQEC_WP_WOUT], t_in = ph[TQEC_WP_TIN];
  const uint32_t tin = (uint32_t)ph[TQEC_WP_TINMASK], toutm = (uint32_t)ph[TQEC_WP_TOUTMASK];
  const int n_groups = bf[0], n_dep = bf[1], n_closes = bf[2];
  const int32_t *pout = bf + 8, *grec = bf + 20, *drec = grec + n_groups * TQEC_BF_GROUP_INTS, *crec = drec + n_dep;
  {
    const uint32_t x = tid < 64 ? (uint32_t)tid : ((uint32_t)(tid - 64) << 6);
    dep_in[tid] = wd_pdep(x, tin);
  }
  // logical tile index -> offset in the output state: the bit at position pout[i] goes to the i-th set bit of toutm
  auto out_off = [&](uint32_t e) {
    uint32_t idx = 0;
    for (int i = 0; i < 12; ++i)
      if (pout[i] >= 0 && ((e >> pout[i]) & 1u)) idx |= 1u << i;
    return wd_pdep(idx, toutm);
  };
  if (tid < 16) {
    const int32_t *rl = grec + (n_groups - 1) * TQEC_BF_GROUP_INTS;
    uint32_t llo = 0, lhi = 0;
    for (int b = 0; b < 4; ++b) {
      if (((tid >> b) & 1) && b < rl[0]) llo |= 1u << (((uint32_t)rl[13] >> (4 * b)) & 15u);
      if (((tid >> b) & 1) && b + 4 < rl[0]) lhi |= 1u << (((uint32_t)rl[13] >> (4 * (b + 4))) & 15u);
    }
    gout_lo[tid] = out_off(llo); gout_hi[tid] = out_off(lhi);
    elog_lo[tid] = (uint16_t)llo; elog_hi[tid] = (uint16_t)lhi;
    if (tid < 8) {
      gb_out[tid] = tid < TQEC_BF_G ? out_off((uint32_t)rl[5 + tid]) : 0u;
      b_log[tid] = tid < TQEC_BF_G ? (uint16_t)rl[5 + tid] : (uint16_t)0;
    }
    if (tid == 0) {
      // a member of the last group's cosets survives when every bit outside the output positions has its dead value: the
      // syndrome bit at a position this group closes, zero elsewhere (padding dimensions, positions closed earlier)
      uint32_t outpos = 0;
      for (int i = 0; i < 12; ++i)
        if (pout[i] >= 0) outpos |= 1u << pout[i];
      s_cmask = 0xfffu & ~outpos;
    }
  }
  for (int g = tid; g < n_groups; g += BF_NT) {
    const int32_t *rec = grec + g * TQEC_BF_GROUP_INTS;
    BfGroupS &G = grp[g];
    G.n_free = (int16_t)rec[0]; G.dep0 = (int16_t)rec[1]; G.n_dep = (int16_t)rec[2]; G.close0 = (int16_t)rec[3]; G.n_closes = (int16_t)rec[4];
    G.zmask = (uint32_t)rec[10];
    for (int j = 0; j < TQEC_BF_G; ++j) G.r[j] = A.vals[rec[11] + j];
    for (int i = 0; i < rec[2]; ++i) s_r[rec[1] + i] = A.vals[rec[11] + TQEC_BF_G + i];
    const uint32_t ord = (uint32_t)rec[13];
    for (int i = 0; i < 16; ++i) {
      uint32_t lo = 0, hi = 0;
      for (int b = 0; b < 4; ++b) {
        if (((i >> b) & 1) && b < rec[0]) lo |= 1u << ((ord >> (4 * b)) & 15u);
        if (((i >> b) & 1) && b + 4 < rec[0]) hi |= 1u << ((ord >> (4 * (b + 4))) & 15u);
      }
      G.orb_lo[i] = (uint16_t)(bf_phys(lo) << 3);
      G.orb_hi[i] = (uint16_t)(bf_phys(hi) << 3);
    }
    for (int j = 0; j < 8; ++j) G.bP[j] = j < TQEC_BF_G ? (uint16_t)(bf_phys((uint32_t)rec[5 + j]) << 3) : (uint16_t)0;
  }
  for (int i = tid; i < n_dep; i += BF_NT) s_code[i] = drec[i];
  for (int i = tid; i < n_closes; i += BF_NT) {
    s_closeP[i] = (uint16_t)(bf_phys(1u << crec[2 * i]) << 3) | (uint16_t)0;
    s_closeB[i] = (uint16_t)crec[2 * i + 1];
    s_closePos[i] = (uint8_t)crec[2 * i];
  }
  const int n_spec = w_in - t_in;
  const uint32_t spec_in = (w_in >= 32 ? 0xffffffffu : ((1u << w_in) - 1u)) & ~tin;
  const uint32_t spec_out = (w_out >= 32 ? 0xffffffffu : ((1u << w_out) - 1u)) & ~toutm;
  for (int i = tid; i < 288; i += BF_NT) {
    const uint32_t x = (uint32_t)(i & 127) << (7 * (i >> 7));
    sdep_in[i] = wd_pdep(x, spec_in);
    sdep_out[i] = wd_pdep(x, spec_out);
  }
  const int64_t n_tiles = A.nb << n_spec;
  const int n_in = 1 << t_in;
  const uint32_t dyn_abs = (uint32_t)__cvta_generic_to_shared(bf_dyn);
  const bool pairs = (tin & 1u) != 0;                            // tile bit 0 = index bit 0: entries come in 16-byte pairs
  __syncthreads();

  // tile -> buffer `which`: asynchronous copies of the input entries, zeros everywhere else
  auto fetch = [&](int64_t tile, int which) {
    const int64_t b = tile >> n_spec;
    const uint32_t sp = (uint32_t)(tile & (((int64_t)1 << n_spec) - 1));
    const double *gi = A.gin + ((size_t)b << A.w_cap) + (sdep_in[sp & 127] | sdep_in[128 + ((sp >> 7) & 127)] | sdep_in[256 + (sp >> 14)]);
    unsigned char *Sb = bf_dyn + which * BF_TILE_BYTES;
    const uint32_t sabs = dyn_abs + (uint32_t)(which * BF_TILE_BYTES);
    if (pairs) {
#pragma unroll 4
      for (int q = tid; q < (n_in >> 1); q += BF_NT) {
        const int l = q << 1;
        bf_cp_async16(sabs + (bf_phys((uint32_t)l) << 3), gi + (dep_in[l & 63] | dep_in[64 + (l >> 6)]));
      }
    } else {
      for (int l = tid; l < n_in; l += BF_NT)
        *reinterpret_cast<double *>(Sb + (bf_phys((uint32_t)l) << 3)) = __ldcs(gi + (dep_in[l & 63] | dep_in[64 + (l >> 6)]));
    }
    for (int q = ((n_in + 1) >> 1) + tid; q < 2048; q += BF_NT)   // (a tile of one entry: its odd neighbour is zeroed below)
      *reinterpret_cast<double2 *>(Sb + (bf_phys((uint32_t)(q << 1)) << 3)) = make_double2(0.0, 0.0);
    if (n_in == 1 && tid == 0) *reinterpret_cast<double *>(Sb + 8) = 0.0;
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int which = 0;
  if ((int64_t)blockIdx.x < n_tiles) fetch(blockIdx.x, 0);
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, which ^= 1) {
    const int64_t b = tile >> n_spec;
    const uint32_t sp = (uint32_t)(tile & (((int64_t)1 << n_spec) - 1));
    double *go = A.gout + ((size_t)b << A.w_cap) + (sdep_out[sp & 127] | sdep_out[128 + ((sp >> 7) & 127)] | sdep_out[256 + (sp >> 14)]);
    const uint64_t *syn = A.synd + (size_t)b * A.nsw;
    unsigned char *Sb = bf_dyn + which * BF_TILE_BYTES;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                             // this tile has landed; everyone is done with the other buffer
    if (tid <= n_groups) {                                       // address masks of the shot's closed checks (read after the first group's barrier)
      const int upto = tid < n_groups ? grp[tid].close0 : n_closes;
      uint32_t x = 0;
      for (int c = 0; c < upto; ++c) {
        const int sb = s_closeB[c];
        if ((__ldg(syn + (sb >> 6)) >> (sb & 63)) & 1ull) x ^= (uint32_t)s_closeP[c];
      }
      s_xp[tid] = x;
    }
    if (tile + gridDim.x < n_tiles) fetch(tile + gridDim.x, which ^ 1);
    if (A.max_in) {                                              // dynamic rescaling, as in k_wide_pass
      const double mx = __longlong_as_double((long long)A.max_in[b]);
      if (mx > 0.0 && mx < 4.909093465297727e-91) {
        const int k = -ilogb(mx);
        const double scale = ldexp(1.0, k);
        if (sp == 0 && tid == 0) A.exps[b] -= k;
        for (int l = tid; l < n_in; l += BF_NT) *reinterpret_cast<double *>(Sb + (bf_phys((uint32_t)l) << 3)) *= scale;
        __syncthreads();
      }
    }
    uint32_t XP = 0;
    double tmax = 0.0;
    // syndrome bits of the checks the last group closes, at their positions
    uint32_t sc_last = 0;
    for (int c = grp[n_groups - 1].close0; c < n_closes; ++c) {
      const int sb = s_closeB[c];
      if ((__ldg(syn + (sb >> 6)) >> (sb & 63)) & 1ull) sc_last |= 1u << s_closePos[c];
    }
    for (int g = 0; g < n_groups; ++g) {
      const BfGroupS &G = grp[g];
      const int n_orb = 1 << G.n_free;
      uint32_t bP[TQEC_BF_G];
#pragma unroll
      for (int j = 0; j < TQEC_BF_G; ++j) bP[j] = G.bP[j];
      for (int i = tid; i < n_orb; i += BF_NT) {
        const uint32_t base = ((uint32_t)G.orb_lo[i & 15] ^ (uint32_t)G.orb_hi[i >> 4]) ^ XP;
        double v[BF_NE];
        // addresses: four independent Gray-code walks over the three low coordinates, one per value of the two high ones
        uint32_t offs[4] = {base, base ^ bP[3], base ^ bP[4], base ^ bP[3] ^ bP[4]};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int gk = k ^ (k >> 1);
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            if (k) offs[h] ^= bP[bf_ctz(k)];
            v[gk | (h << 3)] = *reinterpret_cast<const double *>(Sb + offs[h]);
          }
        }
        const uint32_t Z = G.zmask;
        if (Z) {                                                 // a reused position reopens: its dead half holds stale entries
#pragma unroll
          for (int k = 0; k < BF_NE; ++k)
            if ((Z >> k) & 1u) v[k] = 0.0;
        }
        bf_pair<1>(v, G.r[0]);
        bf_pair<2>(v, G.r[1]);
        bf_pair<4>(v, G.r[2]);
        bf_pair<8>(v, G.r[3]);
        bf_pair<16>(v, G.r[4]);
        for (int s = G.dep0; s < G.dep0 + G.n_dep; ++s) {
          const double r = s_r[s];
          switch (s_code[s]) {
#define BF_CASE(C) case C: bf_pair<C>(v, r); break;
            BF_CASE(1) BF_CASE(2) BF_CASE(3) BF_CASE(4) BF_CASE(5) BF_CASE(6) BF_CASE(7) BF_CASE(8) BF_CASE(9) BF_CASE(10)
            BF_CASE(11) BF_CASE(12) BF_CASE(13) BF_CASE(14) BF_CASE(15) BF_CASE(16) BF_CASE(17) BF_CASE(18) BF_CASE(19)
            BF_CASE(20) BF_CASE(21) BF_CASE(22) BF_CASE(23) BF_CASE(24) BF_CASE(25) BF_CASE(26) BF_CASE(27) BF_CASE(28)
            BF_CASE(29) BF_CASE(30) BF_CASE(31)
#undef BF_CASE
            default: break;
          }
        }
        if (g == n_groups - 1) {
          // last group: straight to the state in HBM.  The output offset of a coset member is XOR-linear in its coordinates
          // like its address; a member survives when its bits outside the output positions have their dead values
          uint32_t gbo[TQEC_BF_G], bl[TQEC_BF_G];
#pragma unroll
          for (int j = 0; j < TQEC_BF_G; ++j) { gbo[j] = gb_out[j]; bl[j] = b_log[j]; }
          const uint32_t cmask = s_cmask;
          const uint32_t ebase = ((uint32_t)elog_lo[i & 15] | (uint32_t)elog_hi[i >> 4]) ^ sc_last;   // syndrome pre-XOR-ed: survivors have (e & cmask) == 0
          const uint32_t obase = gout_lo[i & 15] | gout_hi[i >> 4];
          uint32_t e4[4] = {ebase, ebase ^ bl[3], ebase ^ bl[4], ebase ^ bl[3] ^ bl[4]};
          uint32_t o4[4] = {obase, obase ^ gbo[3], obase ^ gbo[4], obase ^ gbo[3] ^ gbo[4]};
          if (A.max_out) {                                     // dynamic rescaling: also track the shot's largest stored entry
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int gk = k ^ (k >> 1);
#pragma unroll
              for (int h = 0; h < 4; ++h) {
                if (k) { e4[h] ^= bl[bf_ctz(k)]; o4[h] ^= gbo[bf_ctz(k)]; }
                if ((e4[h] & cmask) == 0) {
                  __stcs(go + o4[h], v[gk | (h << 3)]);
                  tmax = fmax(tmax, v[gk | (h << 3)]);
                }
              }
            }
          } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int gk = k ^ (k >> 1);
#pragma unroll
              for (int h = 0; h < 4; ++h) {
                if (k) { e4[h] ^= bl[bf_ctz(k)]; o4[h] ^= gbo[bf_ctz(k)]; }
                if ((e4[h] & cmask) == 0) __stcs(go + o4[h], v[gk | (h << 3)]);
              }
            }
          }
        } else {
          // the walks ended at Gray code 4 (k = 7); walk back
#pragma unroll
          for (int k = 7; k >= 0; --k) {
            const int gk = k ^ (k >> 1);
#pragma unroll
            for (int h = 0; h < 4; ++h) {
              *reinterpret_cast<double *>(Sb + offs[h]) = v[gk | (h << 3)];
              if (k) offs[h] ^= bP[bf_ctz(k)];
            }
          }
        }
      }
      if (g < n_groups - 1) {
        __syncthreads();
        XP = s_xp[g + 1];
      }
    }
    if (A.max_out) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) tmax = fmax(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
      if ((tid & 31) == 0 && tmax > 0.0) atomicMax(A.max_out + b, (unsigned long long)__double_as_longlong(tmax));
    }
  }
}

__global__ void k_wide_init(double *g, int64_t nb, int w_cap, unsigned long long *mx, int32_t *exps) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nb) {
    g[(size_t)i << w_cap] = 1.0;
    if (mx) { mx[i] = (unsigned long long)__double_as_longlong(1.0); exps[i] = 0; }
  }
}

// marginals over the open observable slots (observable 0 fastest) and their first maximal entry (findmax)
__global__ void k_wide_out(const WideDev P, const double *__restrict__ g, int64_t nb, double *__restrict__ out,
                           int32_t *__restrict__ argmax_out, const int32_t *__restrict__ exps, int32_t *__restrict__ log2_out) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  const int NO = 1 << P.n_obs;
  double best = -1.0;
  int bi = 0;
  for (int idx = 0; idx < NO; ++idx) {
    uint32_t src = 0;
    for (int o = 0; o < P.n_obs; ++o) src |= ((uint32_t)(idx >> o) & 1u) << P.obs_pos[o];
    const double v = g[((size_t)b << P.w_cap) + src] * P.out_mant;
    // with a separate exponent output the mantissas are returned as they are; otherwise the exponent is applied here
    // (values below the FP64 range flush to zero, like the reference's; the argmax is taken on the mantissas)
    out[b * NO + idx] = (exps && !log2_out) ? ldexp(v, exps[b]) : v;
    if (v > best) { best = v; bi = idx; }
  }
  if (argmax_out) argmax_out[b] = bi;
  if (log2_out) log2_out[b] = exps ? exps[b] : 0;
}

static const int WD_THREADS = 256;

static size_t wide_smem_bytes(int t_max, int n_tab, int n_ints, int ns) {
  size_t b = ((size_t)16 << t_max);
  b += (size_t)((n_tab + 1) & ~1) * 8;
  b += (size_t)((n_ints + 3) & ~3) * 4;
  b += (size_t)ns * TQEC_WIDE_STEP_INTS * 4;
  b += (256 + 256 + 2 * WD_MAX_STEPS) * 4;
  return (b + 15) & ~(size_t)15;
}

template <typename T>
static int wd_upload(void **slot, const T *src, size_t n) {
  TQEC_CUDA(cudaMalloc(slot, (n ? n : 1) * sizeof(T)));
  if (n) TQEC_CUDA(cudaMemcpy(*slot, src, n * sizeof(T), cudaMemcpyHostToDevice));
  return TQEC_OK;
}

void wide_destroy(tqec_plan *p) {
  for (int i = 0; i < 6; ++i) if (p->d_wd[i]) cudaFree(p->d_wd[i]);
  if (p->wd_bf_off) { std::free(p->wd_bf_off); p->wd_bf_off = nullptr; }
  for (int i = 0; i < 2; ++i) if (p->d_wd_state[i]) cudaFree(p->d_wd_state[i]);
  if (p->d_wd_max) cudaFree(p->d_wd_max);
  if (p->d_wd_exp) cudaFree(p->d_wd_exp);
}

// Validate and upload the global-memory lowering of a plan descriptor.
int wide_create(tqec_plan *p, const tqec_plan_desc *d, const cudaDeviceProp &prop) {
  const tqec_wide_desc *w = d->wide;
  p->has_wide = 0;
  TQEC_REQUIRE(d->semiring == TQEC_SEMIRING_SUMPROD, "wide: the global-memory executor runs sum-product plans only");
  TQEC_REQUIRE(w->n_pass > 0 && w->n_steps > 0 && w->pass_hdr && w->step_hdr && w->ints && w->tables, "wide: missing table");
  TQEC_REQUIRE(w->w_cap >= 0 && w->w_cap <= 31 && w->t_max >= 1 && w->t_max <= 13, "wide: w_cap=%d t_max=%d out of range", w->w_cap, w->t_max);
  TQEC_REQUIRE(d->n_obs == 0 || w->obs_pos, "wide: obs_pos is NULL");
  size_t smem = 0;
  int wcur = 0, step = 0;
  for (int i = 0; i < w->n_pass; ++i) {
    const int32_t *h = w->pass_hdr + (size_t)i * TQEC_WIDE_PASS_INTS;
    const int w_in = h[TQEC_WP_WIN], w_out = h[TQEC_WP_WOUT], t_in = h[TQEC_WP_TIN], t_out = h[TQEC_WP_TOUT], ns = h[TQEC_WP_NSTEPS];
    TQEC_REQUIRE(w_in == wcur, "wide pass %d: w_in=%d does not continue the previous width %d", i, w_in, wcur);
    TQEC_REQUIRE(w_in <= w->w_cap && w_out <= w->w_cap && t_in <= w->t_max && t_out <= w->t_max && t_in <= w_in && t_out <= w_out &&
                     w_in - t_in == w_out - t_out,
                 "wide pass %d: inconsistent widths (w %d -> %d, tile %d -> %d)", i, w_in, w_out, t_in, t_out);
    TQEC_REQUIRE(__builtin_popcount((uint32_t)h[TQEC_WP_TINMASK]) == t_in && ((uint64_t)(uint32_t)h[TQEC_WP_TINMASK] >> w_in) == 0 &&
                     __builtin_popcount((uint32_t)h[TQEC_WP_TOUTMASK]) == t_out && ((uint64_t)(uint32_t)h[TQEC_WP_TOUTMASK] >> w_out) == 0,
                 "wide pass %d: bad tile masks", i);
    TQEC_REQUIRE(ns >= 1 && ns <= WD_MAX_STEPS && h[TQEC_WP_STEP0] == step && step + ns <= w->n_steps, "wide pass %d: bad step range", i);
    const int64_t oi = h[TQEC_WP_OFF_INTS], ni = h[TQEC_WP_N_INTS], ot = h[TQEC_WP_OFF_TAB], nt = h[TQEC_WP_N_TAB];
    TQEC_REQUIRE(oi >= 0 && ni >= 0 && oi + ni <= w->n_ints && ot >= 0 && nt >= 0 && ot + nt <= w->n_tables, "wide pass %d: pool block out of range", i);
    int t = t_in;
    for (int s = 0; s < ns; ++s) {
      const int32_t *q = w->step_hdr + (size_t)(step + s) * TQEC_WIDE_STEP_INTS;
      const int lw_in = q[TQEC_WL_WIN], n_open = q[TQEC_WL_NOPEN], n_close = q[TQEC_WL_NCLOSE], lw_out = q[TQEC_WL_WOUT], nk = q[TQEC_WL_NK];
      TQEC_REQUIRE(lw_in == t && n_open >= 0 && n_open <= 10 && n_close >= 0 && lw_out == lw_in + n_open - n_close && lw_out >= 0 &&
                       lw_out <= w->t_max && lw_in + n_open <= 31,
                   "wide step %d: inconsistent widths", step + s);
      TQEC_REQUIRE(nk >= 1 && nk <= 1024, "wide step %d: bad candidate count %d", step + s, nk);
      const int64_t np = (int64_t)1 << n_open;
      TQEC_REQUIRE(q[TQEC_WL_OFF_T] >= 0 && q[TQEC_WL_OFF_T] + np * nk <= nt, "wide step %d: table offset out of range", step + s);
      TQEC_REQUIRE(q[TQEC_WL_OFF_ML] >= 0 && q[TQEC_WL_OFF_ML] + np <= ni && q[TQEC_WL_OFF_MK] >= 0 && q[TQEC_WL_OFF_MK] + nk <= ni &&
                       q[TQEC_WL_OFF_CLOSE] >= 0 && q[TQEC_WL_OFF_CLOSE] + 2 * (int64_t)n_close <= ni,
                   "wide step %d: int table out of range", step + s);
      const int32_t *I = w->ints + oi;
      for (int pp = 0; pp < np; ++pp) TQEC_REQUIRE(((uint32_t)I[q[TQEC_WL_OFF_ML] + pp] >> lw_in) == 0, "wide step %d: mask leaves the tile", step + s);
      for (int k = 0; k < nk; ++k) TQEC_REQUIRE(((uint32_t)I[q[TQEC_WL_OFF_MK] + k] >> lw_in) == 0, "wide step %d: mask leaves the tile", step + s);
      uint32_t used = 0;
      for (int c = 0; c < n_close; ++c) {
        const int slot = I[q[TQEC_WL_OFF_CLOSE] + 2 * c], bit = I[q[TQEC_WL_OFF_CLOSE] + 2 * c + 1];
        TQEC_REQUIRE(slot >= 0 && slot < lw_in + n_open && !((used >> slot) & 1u), "wide step %d: bad closed slot", step + s);
        TQEC_REQUIRE(bit >= 0 && bit < d->n_checks, "wide step %d: syndrome bit out of range", step + s);
        used |= 1u << slot;
      }
      const uint32_t keep = (uint32_t)q[TQEC_WL_KEEPMASK];
      TQEC_REQUIRE((keep & used) == 0 && __builtin_popcount(keep) == lw_out && ((uint64_t)keep >> (lw_in + n_open)) == 0,
                   "wide step %d: bad keep mask", step + s);
      t = lw_out;
    }
    TQEC_REQUIRE(t == t_out, "wide pass %d: steps end at tile width %d, header says %d", i, t, t_out);
    const size_t need = wide_smem_bytes(w->t_max, (int)nt, (int)ni, ns);
    if (need > smem) smem = need;
    wcur = w_out;
    step += ns;
  }
  TQEC_REQUIRE(wcur == d->n_obs && step == w->n_steps, "wide: final state has %d bits but n_obs=%d", wcur, d->n_obs);
  for (int o = 0; o < d->n_obs; ++o) TQEC_REQUIRE(w->obs_pos[o] >= 0 && w->obs_pos[o] < d->n_obs, "wide: obs_pos[%d] out of range", o);
  if (smem > (size_t)prop.sharedMemPerBlockOptin) {
    set_error("wide: a pass needs %zu B of shared memory > %zu available", smem, (size_t)prop.sharedMemPerBlockOptin);
    return TQEC_ERR_UNSUPPORTED;
  }
  WideDev &D = p->wd;
  std::memset(&D, 0, sizeof(D));
  D.n_pass = w->n_pass; D.n_steps = w->n_steps; D.w_cap = w->w_cap; D.t_max = w->t_max; D.n_obs = d->n_obs; D.nsw = words_for(d->n_checks);
  for (int o = 0; o < d->n_obs; ++o) D.obs_pos[o] = w->obs_pos[o];
  int rc;
  if ((rc = wd_upload(&p->d_wd[0], w->pass_hdr, (size_t)w->n_pass * TQEC_WIDE_PASS_INTS))) return rc;
  if ((rc = wd_upload(&p->d_wd[1], w->step_hdr, (size_t)w->n_steps * TQEC_WIDE_STEP_INTS))) return rc;
  if ((rc = wd_upload(&p->d_wd[2], w->ints, (size_t)w->n_ints))) return rc;
  if ((rc = wd_upload(&p->d_wd[3], w->tables, (size_t)w->n_tables))) return rc;
  D.pass_hdr = (const int32_t *)p->d_wd[0]; D.step_hdr = (const int32_t *)p->d_wd[1]; D.ints = (const int32_t *)p->d_wd[2];
  D.tables = (const double *)p->d_wd[3];
  // butterfly encoding of the rank-1 passes (TQEC_WIDE_NO_BF=1 keeps every pass on k_wide_pass)
  D.out_mant = 1.0;
  p->wd_bf_off = nullptr; p->wd_bf_grid = 0;
  if (w->bf_off && w->bf_ints && w->bf_vals && std::getenv("TQEC_WIDE_NO_BF") == nullptr) {
    bool any = false;
    for (int i = 0; i < w->n_pass; ++i) {
      const int32_t o = w->bf_off[i];
      if (o < 0) continue;
      any = true;
      TQEC_REQUIRE((int64_t)o + 20 <= w->n_bf_ints, "wide: butterfly block of pass %d out of range", i);
      const int32_t *bfp = w->bf_ints + o;
      const int ng = bfp[0], ns = bfp[1], nc = bfp[2];
      TQEC_REQUIRE(ng >= 1 && ng <= BF_MAX_GROUPS && ns >= 0 && ns <= BF_MAX_STEPS && nc >= 0 && nc <= BF_MAX_CLOSES && bfp[4] <= 12 &&
                       bfp[7] == TQEC_BF_G &&
                       (int64_t)o + 20 + (int64_t)ng * TQEC_BF_GROUP_INTS + ns + 2 * (int64_t)nc <= w->n_bf_ints,
                   "wide: bad butterfly block for pass %d", i);
      const int32_t *h = w->pass_hdr + (size_t)i * TQEC_WIDE_PASS_INTS;
      TQEC_REQUIRE(h[TQEC_WP_TIN] <= 12 && h[TQEC_WP_TOUT] <= 12 && bfp[5] == h[TQEC_WP_TIN] && bfp[6] == h[TQEC_WP_TOUT] &&
                       h[TQEC_WP_WIN] - h[TQEC_WP_TIN] <= 19,
                   "wide: butterfly block of pass %d does not match its header", i);
      const int32_t *grec = bfp + 20, *crec = grec + ng * TQEC_BF_GROUP_INTS + ns;
      int s_sum = 0, c_sum = 0;
      for (int g = 0; g < ng; ++g) {
        const int32_t *r = grec + g * TQEC_BF_GROUP_INTS;
        TQEC_REQUIRE(r[0] >= 0 && r[0] <= 8 && r[1] == s_sum && r[2] >= 0 && r[3] == c_sum && r[4] >= 0 && r[11] >= 0 &&
                         (int64_t)r[11] + TQEC_BF_G + r[2] <= w->n_bf_vals,
                     "wide: bad butterfly group %d of pass %d", g, i);
        for (int j = 0; j < TQEC_BF_G; ++j) TQEC_REQUIRE(r[5 + j] > 0 && r[5 + j] < 4096, "wide: bad basis vector in pass %d", i);
        uint32_t seen = 0;
        for (int q = 0; q < r[0]; ++q) {
          const uint32_t pq = ((uint32_t)r[13] >> (4 * q)) & 15u;
          TQEC_REQUIRE(pq < 12 && !((seen >> pq) & 1u), "wide: bad coset enumeration order in pass %d", i);
          seen |= 1u << pq;
        }
        s_sum += r[2]; c_sum += r[4];
      }
      TQEC_REQUIRE(s_sum == ns && c_sum == nc, "wide: butterfly groups of pass %d do not cover its steps", i);
      {
        uint32_t seen = 0;
        for (int q = 0; q < 12; ++q) {
          const int pq = bfp[8 + q];
          TQEC_REQUIRE(q < bfp[6] ? (pq >= 0 && pq < 12 && !((seen >> pq) & 1u)) : pq == -1, "wide: bad output positions in pass %d", i);
          if (pq >= 0) seen |= 1u << pq;
        }
      }
      for (int q = 0; q < ns; ++q) TQEC_REQUIRE(crec[q - ns] > 0 && crec[q - ns] < (1 << TQEC_BF_G), "wide: bad dependent step in pass %d", i);
      for (int c = 0; c < nc; ++c)
        TQEC_REQUIRE(crec[2 * c] >= 0 && crec[2 * c] < 12 && crec[2 * c + 1] >= 0 && crec[2 * c + 1] < d->n_checks, "wide: bad close record in pass %d", i);
    }
    if (any) {
      if ((rc = wd_upload(&p->d_wd[4], w->bf_ints, (size_t)w->n_bf_ints))) return rc;
      if ((rc = wd_upload(&p->d_wd[5], w->bf_vals, (size_t)w->n_bf_vals))) return rc;
      D.bf_ints = (const int32_t *)p->d_wd[4]; D.bf_vals = (const double *)p->d_wd[5];
      p->wd_bf_off = (int32_t *)std::malloc(sizeof(int32_t) * (size_t)w->n_pass);
      std::memcpy(p->wd_bf_off, w->bf_off, sizeof(int32_t) * (size_t)w->n_pass);
      D.out_mant = w->bf_mant;
      p->log2_scale += w->bf_log2;
      int per_sm_bf = 0;
      TQEC_CUDA(cudaFuncSetAttribute(k_wide_bf, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * BF_TILE_BYTES));
      TQEC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_bf, k_wide_bf, BF_NT, 2 * BF_TILE_BYTES));
      if (per_sm_bf < 1) per_sm_bf = 1;
      if (const char *e = std::getenv("TQEC_WIDE_BF_CTAS")) { const int v = std::atoi(e); if (v >= 1 && v < per_sm_bf) per_sm_bf = v; }
      p->wd_bf_grid = per_sm_bf * p->sm_count;
    }
  }
  p->wd_smem = (int)smem;
  TQEC_CUDA(cudaFuncSetAttribute(k_wide_pass<WD_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  TQEC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_wide_pass<WD_THREADS>, WD_THREADS, smem));
  if (per_sm < 1) per_sm = 1;
  p->wd_grid = per_sm * p->sm_count;
  p->teams_per_sm = per_sm;
  p->team_threads = WD_THREADS;
  p->smem_bytes = (int)smem;
  p->grid_max = p->wd_grid;
  // candidate evaluations per shot
  double cand = 0.0;
  for (int i = 0; i < w->n_pass; ++i) {
    const int32_t *h = w->pass_hdr + (size_t)i * TQEC_WIDE_PASS_INTS;
    const int n_spec = h[TQEC_WP_WIN] - h[TQEC_WP_TIN];
    for (int s = 0; s < h[TQEC_WP_NSTEPS]; ++s) {
      const int32_t *q = w->step_hdr + (size_t)(h[TQEC_WP_STEP0] + s) * TQEC_WIDE_STEP_INTS;
      cand += std::ldexp(1.0, n_spec + q[TQEC_WL_WOUT]) * q[TQEC_WL_NK];
    }
  }
  p->candidates_per_shot = cand;
  p->wd_dynamic = (d->flags & TQEC_PLAN_DYNAMIC_RESCALE) ? 1 : 0;
  p->has_wide = 1;
  return TQEC_OK;
}

// Shots whose states fit the device at once: 80 % of the free memory (TQEC_WIDE_MEM_GB overrides), two arrays, at most
// 4096 shots.  The arrays are allocated at the first decode and grown when a larger batch arrives.
static int wide_reserve(tqec_plan *p, int64_t want) {
  if (p->wd_batch >= want || (p->wd_batch > 0 && p->wd_full)) return TQEC_OK;
  const size_t per_shot = (size_t)16 << p->wd.w_cap;
  for (int i = 0; i < 2; ++i) { if (p->d_wd_state[i]) cudaFree(p->d_wd_state[i]); p->d_wd_state[i] = nullptr; }
  if (p->d_wd_max) { cudaFree(p->d_wd_max); p->d_wd_max = nullptr; }
  if (p->d_wd_exp) { cudaFree(p->d_wd_exp); p->d_wd_exp = nullptr; }
  p->wd_batch = 0;
  size_t free_b = 0, total_b = 0;
  TQEC_CUDA(cudaMemGetInfo(&free_b, &total_b));
  double budget = 0.8 * (double)free_b;
  if (const char *e = std::getenv("TQEC_WIDE_MEM_GB")) { const double v = std::atof(e) * 1e9; if (v > 0 && v < budget) budget = v; }
  int64_t nb = (int64_t)(budget / (double)per_shot);
  if (nb > 4096) nb = 4096;
  p->wd_full = nb <= want;
  if (nb > want) nb = want;
  if (nb < 1) {
    set_error("wide: one shot needs %zu B of state, %zu B free on the device", per_shot, free_b);
    return TQEC_ERR_NOMEM;
  }
  for (int i = 0; i < 2; ++i) {
    cudaError_t e = cudaMalloc((void **)&p->d_wd_state[i], (size_t)nb * (per_shot / 2));
    if (e != cudaSuccess) {
      set_error("cudaMalloc(%zu B wide state): %s", (size_t)nb * (per_shot / 2), cudaGetErrorString(e));
      return TQEC_ERR_NOMEM;
    }
  }
  if (p->wd_dynamic) {
    TQEC_CUDA(cudaMalloc((void **)&p->d_wd_max, (size_t)nb * 2 * sizeof(unsigned long long)));
    TQEC_CUDA(cudaMalloc((void **)&p->d_wd_exp, (size_t)nb * sizeof(int32_t)));
  }
  p->wd_batch = nb;
  return TQEC_OK;
}

int launch_wide(tqec_plan *plan, const uint64_t *d_synd, int64_t B, double *d_out, int32_t *d_argmax, cudaStream_t stream,
                int32_t *d_log2) {
  tqec::NvtxRange nvtx_range("tqec_wide_decode");
  int rc = wide_reserve(plan, B);
  if (rc) return rc;
  const WideDev &D = plan->wd;
  const int NO = 1 << D.n_obs;
  const bool dyn = plan->wd_dynamic != 0;
  for (int64_t o = 0; o < B; o += plan->wd_batch) {
    const int64_t nb = B - o < plan->wd_batch ? B - o : plan->wd_batch;
    unsigned long long *mx[2] = {dyn ? plan->d_wd_max : nullptr, dyn ? plan->d_wd_max + plan->wd_batch : nullptr};
    k_wide_init<<<(unsigned)((nb + 255) / 256), 256, 0, stream>>>(plan->d_wd_state[0], nb, D.w_cap, mx[0], plan->d_wd_exp);
    int cur = 0;
    for (int i = 0; i < D.n_pass; ++i) {
      WidePassArgs A;
      A.gin = plan->d_wd_state[cur]; A.gout = plan->d_wd_state[cur ^ 1];
      A.synd = d_synd + (size_t)o * D.nsw; A.nb = nb;
      A.pass = D.pass_hdr + (size_t)i * TQEC_WIDE_PASS_INTS; A.step_hdr = D.step_hdr; A.ints = D.ints; A.tables = D.tables;
      A.nsw = D.nsw; A.w_cap = D.w_cap; A.t_max = D.t_max;
      A.max_in = mx[cur]; A.max_out = mx[cur ^ 1]; A.exps = plan->d_wd_exp;
      if (dyn) TQEC_CUDA(cudaMemsetAsync(mx[cur ^ 1], 0, (size_t)nb * sizeof(unsigned long long), stream));
      // tiles of the pass: the host copy of the header is not kept, so size the grid for the widest case and let the
      // kernel's tile loop run short
      if (plan->wd_bf_off && plan->wd_bf_off[i] >= 0) {
        WideBfArgs Bf;
        Bf.gin = A.gin; Bf.gout = A.gout; Bf.synd = A.synd; Bf.nb = nb; Bf.pass = A.pass;
        Bf.bf = D.bf_ints + plan->wd_bf_off[i]; Bf.vals = D.bf_vals; Bf.nsw = D.nsw; Bf.w_cap = D.w_cap;
        Bf.max_in = A.max_in; Bf.max_out = A.max_out; Bf.exps = A.exps;
        k_wide_bf<<<plan->wd_bf_grid, BF_NT, 2 * BF_TILE_BYTES, stream>>>(Bf);
      } else
      k_wide_pass<WD_THREADS><<<plan->wd_grid, WD_THREADS, plan->wd_smem, stream>>>(A);
      cur ^= 1;
      plan->launches += 1;
    }
    k_wide_out<<<(unsigned)((nb + 127) / 128), 128, 0, stream>>>(D, plan->d_wd_state[cur], nb, d_out + (size_t)o * NO,
                                                                 d_argmax ? d_argmax + o : nullptr, dyn ? plan->d_wd_exp : nullptr,
                                                                 d_log2 ? d_log2 + o : nullptr);
    TQEC_CUDA(cudaGetLastError());
    plan->launches += 2;
  }
  return TQEC_OK;
}

}  // namespace tqec
