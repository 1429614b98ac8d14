// Bit-packed GF(2) front / back end of the decoding hot path for sm_100a: Philox error sampling, syndrome
// extraction (popcount-parity mat-vec), logical-error flags + counters, TNMMAP coset representative, and the fused
// Monte-Carlo pipeline that chains them around the decoder.  All of it is integer work moving < 100 B per shot;
// one thread owns one shot and keeps its words in registers, matrix rows are staged in shared memory.
#include <cstring>

#include "tqec_common.h"

namespace tqec {

static thread_local std::string g_err;

void set_error(const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
}

int ensure_cap(void **ptr, size_t *cap, size_t bytes) {
  if (*cap >= bytes && *ptr) return TQEC_OK;
  if (*ptr) cudaFree(*ptr);
  *ptr = nullptr;
  *cap = 0;
  size_t want = bytes + bytes / 4 + 256;
  cudaError_t e = cudaMalloc(ptr, want);
  if (e != cudaSuccess) {
    set_error("cudaMalloc(%zu B) failed: %s", want, cudaGetErrorString(e));
    return TQEC_ERR_NOMEM;
  }
  *cap = want;
  return TQEC_OK;
}

// ---- Philox4x32-10 ------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox_round(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3, uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
  const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
  c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
}

__device__ __forceinline__ double philox_uniform(uint64_t seed, uint64_t shot, uint32_t site) {
  uint32_t c0 = (uint32_t)shot, c1 = (uint32_t)(shot >> 32), c2 = site, c3 = 0u;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c0, c1, c2, c3, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  const uint64_t bits = (((uint64_t)c0 << 32) | c1) >> 11;
  return (double)bits * (1.0 / 9007199254740992.0);
}

// Raw 53-bit draw of philox_uniform: u = bits * 2^-53 exactly, so `u < p` is the integer compare `bits < ceil(p * 2^53)`
// (scaling a double by a power of two is exact): same decisions, no int -> FP64 conversion and no FP64 compare per site.
__device__ __forceinline__ uint64_t philox_bits53(uint64_t seed, uint64_t shot, uint32_t site) {
  uint32_t c0 = (uint32_t)shot, c1 = (uint32_t)(shot >> 32), c2 = site, c3 = 0u;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c0, c1, c2, c3, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return (((uint64_t)c0 << 32) | c1) >> 11;
}
__device__ __forceinline__ uint64_t prob_threshold53(double p) {       // negative / NaN -> 0 (never), p >= 1 -> always
  return p > 0.0 ? __double2ull_ru(fmin(p, 1.0) * 9007199254740992.0) : 0ull;
}

// one thread per (shot, output word): sites are visited word by word so that every thread writes whole words
__global__ void k_sample_errors(int model, int n_sites, const double *__restrict__ p, uint64_t seed, int64_t shot_offset,
                                int64_t B, uint64_t *__restrict__ err, int words) {
  const int64_t gid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (gid >= B * words) return;
  const int64_t s = gid / words;
  const int w = (int)(gid - s * words);
  const uint64_t shot = (uint64_t)(shot_offset + s);
  uint64_t out = 0;
  const int nbits = model == TQEC_MODEL_DEPOL ? 2 * n_sites : n_sites;
  for (int b = 0; b < 64; ++b) {
    const int bit = w * 64 + b;
    if (bit >= nbits) break;
    if (model == TQEC_MODEL_FLIP) {                                  // u < p as an integer compare (see prob_threshold53)
      if (philox_bits53(seed, shot, (uint32_t)bit) < prob_threshold53(__ldg(p + bit))) out |= 1ull << b;
    } else {
      const int q = bit < n_sites ? bit : bit - n_sites;
      const double u = philox_uniform(seed, shot, (uint32_t)q);
      const double px = p[q], py = p[n_sites + q], pz = p[2 * n_sites + q];
      // Y first, then X, then Z (error_model.jl:101-115)
      const bool isY = u < py;
      const bool isX = !isY && u < px + py;
      const bool isZ = !isY && !isX && u < px + py + pz;
      const bool on = bit < n_sites ? (isX || isY) : (isZ || isY);
      if (on) out |= 1ull << b;
    }
  }
  err[gid] = out;
}

// Depolarizing model, one thread per SHOT: the draw of qubit q decides both its X bit (position q) and its Z bit (position
// n_sites + q), so it is made once (k_sample_errors makes it once per output bit, i.e. twice); thresholds Y | X+Y | X+Y+Z
// as 53-bit integers in shared memory; the thread assembles its shot's words (<= SAMPLE_MAXW) and writes them out.
#define SAMPLE_MAXW 16
__global__ void __launch_bounds__(128)
k_sample_depol(int n_sites, const double *__restrict__ p, uint64_t seed, int64_t shot_offset, int64_t B,
               uint64_t *__restrict__ err, int words) {
  extern __shared__ uint64_t sh_thr[];                                  // [3][n_sites]
  for (int q = threadIdx.x; q < n_sites; q += blockDim.x) {
    const double px = p[q], py = p[n_sites + q], pz = p[2 * n_sites + q];
    sh_thr[q] = prob_threshold53(py);                                   // Y first, then X, then Z (error_model.jl:101-115)
    sh_thr[n_sites + q] = prob_threshold53(px + py);
    sh_thr[2 * n_sites + q] = prob_threshold53(px + py + pz);
  }
  __syncthreads();
  const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (s >= B) return;
  const uint64_t shot = (uint64_t)(shot_offset + s);
  uint64_t out[SAMPLE_MAXW];
#pragma unroll
  for (int w = 0; w < SAMPLE_MAXW; ++w) out[w] = 0ull;
  for (int q0 = 0; q0 < n_sites; q0 += 64) {
    uint64_t xw = 0ull, zw = 0ull;
    const int nb = min(64, n_sites - q0);
#pragma unroll 4
    for (int b = 0; b < nb; ++b) {
      const int q = q0 + b;
      const uint64_t u = philox_bits53(seed, shot, (uint32_t)q);
      const bool isY = u < sh_thr[q];
      const bool isXY = u < sh_thr[n_sites + q];                        // X or Y (thresholds ascend unless a p is negative)
      const bool isX = !isY && isXY;
      const bool isZ = !isY && !isX && u < sh_thr[2 * n_sites + q];
      xw |= (uint64_t)(isX || isY) << b;
      zw |= (uint64_t)(isZ || isY) << b;
    }
    const int zp = n_sites + q0, zs = zp & 63;
    out[q0 >> 6] |= xw;
    out[zp >> 6] |= zw << zs;
    if (zs && (zp >> 6) + 1 < words) out[(zp >> 6) + 1] |= zw >> (64 - zs);
  }
  for (int w = 0; w < words; ++w) err[s * words + w] = out[w];
}

int launch_sample(int model, int n_sites, const double *d_p, uint64_t seed, int64_t shot_offset, int64_t B,
                  uint64_t *d_err, int words, cudaStream_t stream) {
  if (B <= 0) return TQEC_OK;
  if (model == TQEC_MODEL_DEPOL && words <= SAMPLE_MAXW && n_sites > 0 && (size_t)n_sites * 24 <= 40 * 1024) {
    const int threads = 128;
    k_sample_depol<<<(unsigned)((B + threads - 1) / threads), threads, (size_t)n_sites * 24, stream>>>(n_sites, d_p, seed, shot_offset,
                                                                                                      B, d_err, words);
    TQEC_CUDA(cudaGetLastError());
    return TQEC_OK;
  }
  const int64_t n = B * words;
  const int threads = 256;
  k_sample_errors<<<(unsigned)((n + threads - 1) / threads), threads, 0, stream>>>(model, n_sites, d_p, seed, shot_offset, B, d_err, words);
  TQEC_CUDA(cudaGetLastError());
  return TQEC_OK;
}

// ---- packed mat-vec: out[shot] = M in[shot] -------------------------------------------------------------------
template <int CW>
__global__ void k_gf2_apply(const uint64_t *__restrict__ rowsM, int rows, int cw_rt, int rw, const uint64_t *__restrict__ in,
                            int64_t B, uint64_t *__restrict__ out) {
  extern __shared__ uint64_t sh_rows[];
  const int cw = CW > 0 ? CW : cw_rt;
  for (int i = threadIdx.x; i < rows * cw; i += blockDim.x) sh_rows[i] = rowsM[i];
  __syncthreads();
  const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (s >= B) return;
  uint64_t x[CW > 0 ? CW : 1];
  if (CW > 0) {
#pragma unroll
    for (int w = 0; w < CW; ++w) x[w] = in[s * CW + w];
  }
  for (int ow = 0; ow < rw; ++ow) {
    uint64_t o = 0;
    const int r1 = min(rows, (ow + 1) * 64);
    for (int r = ow * 64; r < r1; ++r) {
      uint64_t acc = 0;
      if (CW > 0) {
#pragma unroll
        for (int w = 0; w < CW; ++w) acc ^= sh_rows[r * CW + w] & x[w];
      } else {
        for (int w = 0; w < cw; ++w) acc ^= sh_rows[r * cw + w] & in[s * cw + w];
      }
      o |= (uint64_t)(__popcll(acc) & 1) << (r & 63);
    }
    out[s * rw + ow] = o;
  }
}

int launch_gf2_apply(tqec_gf2 *m, const uint64_t *d_in, int64_t B, uint64_t *d_out, cudaStream_t stream) {
  if (B <= 0) return TQEC_OK;
  const int threads = 128;
  const unsigned grid = (unsigned)((B + threads - 1) / threads);
  const size_t smem = (size_t)m->rows * m->cw * 8;
  if (smem > 48 * 1024) {
    set_error("GF(2) matrix of %d x %d does not fit the 48 KiB row cache", m->rows, m->cols);
    return TQEC_ERR_UNSUPPORTED;
  }
  switch (m->cw) {
    case 1: k_gf2_apply<1><<<grid, threads, smem, stream>>>(m->d_rows, m->rows, m->cw, m->rw, d_in, B, d_out); break;
    case 2: k_gf2_apply<2><<<grid, threads, smem, stream>>>(m->d_rows, m->rows, m->cw, m->rw, d_in, B, d_out); break;
    case 3: k_gf2_apply<3><<<grid, threads, smem, stream>>>(m->d_rows, m->rows, m->cw, m->rw, d_in, B, d_out); break;
    case 4: k_gf2_apply<4><<<grid, threads, smem, stream>>>(m->d_rows, m->rows, m->cw, m->rw, d_in, B, d_out); break;
    default: k_gf2_apply<0><<<grid, threads, smem, stream>>>(m->d_rows, m->rows, m->cw, m->rw, d_in, B, d_out); break;
  }
  TQEC_CUDA(cudaGetLastError());
  m->launches += 1;
  return TQEC_OK;
}

// ---- logical flags + counters ---------------------------------------------------------------------------------
__global__ void k_logical_flags(const uint64_t *__restrict__ rowsM, const int32_t *__restrict__ row_class, int rows, int cw,
                                const uint64_t *__restrict__ e1, const uint64_t *__restrict__ e2, int64_t B,
                                uint8_t *__restrict__ flags, unsigned long long *__restrict__ counts) {
  extern __shared__ uint64_t sh_rows[];
  __shared__ unsigned int sh_cnt[3];
  for (int i = threadIdx.x; i < rows * cw; i += blockDim.x) sh_rows[i] = rowsM[i];
  if (threadIdx.x < 3) sh_cnt[threadIdx.x] = 0;
  __syncthreads();
  const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  unsigned f = 0;
  if (s < B) {
    for (int r = 0; r < rows; ++r) {
      uint64_t acc = 0;
      for (int w = 0; w < cw; ++w) {
        const uint64_t d = e1[s * cw + w] ^ (e2 ? e2[s * cw + w] : 0ull);
        acc ^= sh_rows[r * cw + w] & d;
      }
      if (__popcll(acc) & 1) f |= 1u << (row_class[r] & 1);
    }
    if (flags) flags[s] = (uint8_t)f;
  }
  if (counts) {
    const unsigned m0 = __ballot_sync(0xffffffffu, f & 1u), m1 = __ballot_sync(0xffffffffu, f & 2u),
                   ma = __ballot_sync(0xffffffffu, f != 0u);
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(&sh_cnt[0], __popc(m0));
      atomicAdd(&sh_cnt[1], __popc(m1));
      atomicAdd(&sh_cnt[2], __popc(ma));
    }
    __syncthreads();
    if (threadIdx.x < 3 && sh_cnt[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)sh_cnt[threadIdx.x]);
  }
}

int launch_flags(tqec_gf2 *L, const int32_t *d_row_class, const uint64_t *d_e1, const uint64_t *d_e2, int64_t B,
                 uint8_t *d_flags, unsigned long long *d_counts, cudaStream_t stream) {
  if (B <= 0) return TQEC_OK;
  const int threads = 128;
  const size_t smem = (size_t)L->rows * L->cw * 8;
  if (smem > 40 * 1024) {
    set_error("logical matrix of %d x %d does not fit the row cache", L->rows, L->cols);
    return TQEC_ERR_UNSUPPORTED;
  }
  k_logical_flags<<<(unsigned)((B + threads - 1) / threads), threads, smem, stream>>>(L->d_rows, d_row_class, L->rows, L->cw,
                                                                                     d_e1, d_e2, B, d_flags, d_counts);
  TQEC_CUDA(cudaGetLastError());
  L->launches += 1;
  return TQEC_OK;
}

// ---- TNMMAP error pattern: e = R s, then move it into the decoded logical sector --------------------------------
__global__ void k_coset_fix(const uint64_t *__restrict__ Lrows, const uint64_t *__restrict__ Frows, int n_obs, int n_fix,
                            int cw, const int32_t *__restrict__ sector, int64_t B, uint64_t *__restrict__ err,
                            uint8_t *__restrict__ ok_out) {
  const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (s >= B) return;
  // sector of the representative, relative to the requested one
  uint32_t delta = (uint32_t)sector[s];
  for (int i = 0; i < n_obs; ++i) {
    uint64_t acc = 0;
    for (int w = 0; w < cw; ++w) acc ^= Lrows[i * cw + w] & err[s * cw + w];
    delta ^= (uint32_t)(__popcll(acc) & 1) << i;
  }
  // The fix rows are undetectable patterns whose sector flips d_j = L f_j are in reduced echelon form (the lowest set
  // bit of d_j is its pivot and no other row has that bit): one pass moves the representative into the requested sector
  // whenever the difference lies in their span -- including differences that only a JOINT flip of several observables
  // can realise.
  for (int j = 0; j < n_fix && delta; ++j) {
    uint32_t dj = 0;
    for (int i = 0; i < n_obs; ++i) {
      uint64_t acc = 0;
      for (int w = 0; w < cw; ++w) acc ^= Lrows[i * cw + w] & Frows[j * cw + w];
      dj |= (uint32_t)(__popcll(acc) & 1) << i;
    }
    if (delta & dj & (0u - dj)) {
      for (int w = 0; w < cw; ++w) err[s * cw + w] ^= Frows[j * cw + w];
      delta ^= dj;
    }
  }
  if (ok_out) ok_out[s] = delta == 0;
}

}  // namespace tqec

using namespace tqec;

extern "C" const char *tqec_last_error(void) { return g_err.c_str(); }
extern "C" int tqec_version(void) { return 100; }
extern "C" int tqec_device_count(int32_t *out) {
  TQEC_REQUIRE(out, "tqec_device_count: out is NULL");
  int n = 0;
  TQEC_CUDA(cudaGetDeviceCount(&n));
  *out = n;
  return TQEC_OK;
}

extern "C" int tqec_gf2_create(int32_t rows, int32_t cols, const uint64_t *packed_rows, int32_t device, tqec_gf2 **out) {
  TQEC_REQUIRE(out, "tqec_gf2_create: out is NULL");
  *out = nullptr;
  TQEC_REQUIRE(rows >= 0 && cols >= 0 && (rows == 0 || packed_rows), "tqec_gf2_create: bad matrix %d x %d", rows, cols);
  int ndev = 0;
  TQEC_CUDA(cudaGetDeviceCount(&ndev));
  TQEC_REQUIRE(device >= 0 && device < ndev, "tqec_gf2_create: device %d not present (%d visible)", device, ndev);
  TQEC_CUDA(cudaSetDevice(device));
  tqec_gf2 *m = new tqec_gf2();
  std::memset(m, 0, sizeof(*m));
  m->device = device; m->rows = rows; m->cols = cols; m->rw = words_for(rows); m->cw = words_for(cols);
  const size_t bytes = (size_t)(rows ? rows : 1) * m->cw * 8;
  cudaError_t e = cudaMalloc((void **)&m->d_rows, bytes);
  if (e == cudaSuccess && rows) e = cudaMemcpy(m->d_rows, packed_rows, (size_t)rows * m->cw * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    set_error("tqec_gf2_create: %s", cudaGetErrorString(e));
    tqec_gf2_destroy(m);
    return TQEC_ERR_CUDA;
  }
  *out = m;
  return TQEC_OK;
}

extern "C" int tqec_gf2_destroy(tqec_gf2 *m) {
  if (!m) return TQEC_OK;
  cudaSetDevice(m->device);
  cudaFree(m->d_rows);
  for (int i = 0; i < 4; ++i) cudaFree(m->d_io[i]);
  if (m->stream) cudaStreamDestroy(m->stream);
  delete m;
  return TQEC_OK;
}

extern "C" int tqec_gf2_apply_dev(tqec_gf2 *m, const uint64_t *d_in, int64_t B, uint64_t *d_out, void *stream) {
  TQEC_REQUIRE(m && B >= 0 && (B == 0 || (d_in && d_out)), "tqec_gf2_apply: NULL argument");
  TQEC_CUDA(cudaSetDevice(m->device));
  return launch_gf2_apply(m, d_in, B, d_out, (cudaStream_t)stream);
}

extern "C" int tqec_gf2_apply(tqec_gf2 *m, const uint64_t *in, int64_t B, uint64_t *out) {
  TQEC_REQUIRE(m && B >= 0 && (B == 0 || (in && out)), "tqec_gf2_apply: NULL argument");
  if (B == 0) return TQEC_OK;
  TQEC_CUDA(cudaSetDevice(m->device));
  const size_t ib = (size_t)B * m->cw * 8, ob = (size_t)B * m->rw * 8;
  int rc;
  if ((rc = ensure_cap(&m->d_io[0], &m->io_cap[0], ib))) return rc;
  if ((rc = ensure_cap(&m->d_io[1], &m->io_cap[1], ob))) return rc;
  TQEC_CUDA(cudaMemcpyAsync(m->d_io[0], in, ib, cudaMemcpyHostToDevice, m->stream));
  if ((rc = launch_gf2_apply(m, (const uint64_t *)m->d_io[0], B, (uint64_t *)m->d_io[1], m->stream))) return rc;
  TQEC_CUDA(cudaMemcpyAsync(out, m->d_io[1], ob, cudaMemcpyDeviceToHost, m->stream));
  TQEC_CUDA(cudaStreamSynchronize(m->stream));
  return TQEC_OK;
}

extern "C" int tqec_logical_flags(tqec_gf2 *L, const int32_t *row_class, const uint64_t *e1, const uint64_t *e2,
                                  int64_t B, uint8_t *flags_out, int64_t counts[4]) {
  TQEC_REQUIRE(L && row_class && B >= 0 && (B == 0 || e1), "tqec_logical_flags: NULL argument");
  if (B == 0) return TQEC_OK;
  TQEC_CUDA(cudaSetDevice(L->device));
  const size_t eb = (size_t)B * L->cw * 8;
  int rc;
  if ((rc = ensure_cap(&L->d_io[0], &L->io_cap[0], eb))) return rc;
  if ((rc = ensure_cap(&L->d_io[1], &L->io_cap[1], eb))) return rc;
  if ((rc = ensure_cap(&L->d_io[2], &L->io_cap[2], (size_t)B))) return rc;
  if ((rc = ensure_cap(&L->d_io[3], &L->io_cap[3], 64 + (size_t)L->rows * 4))) return rc;
  unsigned long long *d_counts = (unsigned long long *)L->d_io[3];
  int32_t *d_cls = (int32_t *)((char *)L->d_io[3] + 64);
  TQEC_CUDA(cudaMemcpyAsync(L->d_io[0], e1, eb, cudaMemcpyHostToDevice, L->stream));
  if (e2) TQEC_CUDA(cudaMemcpyAsync(L->d_io[1], e2, eb, cudaMemcpyHostToDevice, L->stream));
  TQEC_CUDA(cudaMemsetAsync(d_counts, 0, 32, L->stream));
  TQEC_CUDA(cudaMemcpyAsync(d_cls, row_class, (size_t)L->rows * 4, cudaMemcpyHostToDevice, L->stream));
  if ((rc = launch_flags(L, d_cls, (const uint64_t *)L->d_io[0], e2 ? (const uint64_t *)L->d_io[1] : nullptr, B,
                         (uint8_t *)L->d_io[2], counts ? d_counts : nullptr, L->stream))) return rc;
  if (flags_out) TQEC_CUDA(cudaMemcpyAsync(flags_out, L->d_io[2], (size_t)B, cudaMemcpyDeviceToHost, L->stream));
  unsigned long long h[4] = {0, 0, 0, 0};
  if (counts) TQEC_CUDA(cudaMemcpyAsync(h, d_counts, 24, cudaMemcpyDeviceToHost, L->stream));
  TQEC_CUDA(cudaStreamSynchronize(L->stream));
  if (counts) {
    counts[0] += (int64_t)h[0]; counts[1] += (int64_t)h[1]; counts[2] += (int64_t)h[2]; counts[3] += B;
  }
  return TQEC_OK;
}

extern "C" int tqec_coset_rep(tqec_gf2 *R, tqec_gf2 *L, tqec_gf2 *FIX, const uint64_t *synd, const int32_t *sector,
                              int64_t B, uint64_t *err_out, uint8_t *ok_out) {
  TQEC_REQUIRE(R && B >= 0 && (B == 0 || (synd && err_out)), "tqec_coset_rep: NULL argument");
  TQEC_REQUIRE((L == nullptr) == (FIX == nullptr), "tqec_coset_rep: L and FIX go together");
  if (L) {
    TQEC_REQUIRE(L->cols == R->rows && FIX->cols == R->rows && FIX->rows <= L->rows && L->rows <= 16 && sector,
                 "tqec_coset_rep: L must be n_obs x n_vars (n_obs <= 16), FIX at most n_obs x n_vars, sector non-NULL");
    TQEC_REQUIRE(L->device == R->device && FIX->device == R->device, "tqec_coset_rep: matrices live on different devices");
  }
  if (B == 0) return TQEC_OK;
  TQEC_CUDA(cudaSetDevice(R->device));
  const size_t ib = (size_t)B * R->cw * 8, ob = (size_t)B * R->rw * 8;
  int rc;
  if ((rc = ensure_cap(&R->d_io[0], &R->io_cap[0], ib))) return rc;
  if ((rc = ensure_cap(&R->d_io[1], &R->io_cap[1], ob))) return rc;
  TQEC_CUDA(cudaMemcpyAsync(R->d_io[0], synd, ib, cudaMemcpyHostToDevice, R->stream));
  if ((rc = launch_gf2_apply(R, (const uint64_t *)R->d_io[0], B, (uint64_t *)R->d_io[1], R->stream))) return rc;
  if (L && L->rows > 0) {
    if ((rc = ensure_cap(&R->d_io[2], &R->io_cap[2], (size_t)B * 4))) return rc;
    TQEC_CUDA(cudaMemcpyAsync(R->d_io[2], sector, (size_t)B * 4, cudaMemcpyHostToDevice, R->stream));
    const int threads = 128;
    if (ok_out && (rc = ensure_cap(&R->d_io[3], &R->io_cap[3], (size_t)B))) return rc;
    k_coset_fix<<<(unsigned)((B + threads - 1) / threads), threads, 0, R->stream>>>(L->d_rows, FIX->d_rows, L->rows, FIX->rows, L->cw,
                                                                                   (const int32_t *)R->d_io[2], B, (uint64_t *)R->d_io[1],
                                                                                   ok_out ? (uint8_t *)R->d_io[3] : nullptr);
    TQEC_CUDA(cudaGetLastError());
    R->launches += 1;
    if (ok_out) TQEC_CUDA(cudaMemcpyAsync(ok_out, R->d_io[3], (size_t)B, cudaMemcpyDeviceToHost, R->stream));
  } else if (ok_out) {
    std::memset(ok_out, 1, (size_t)B);
  }
  TQEC_CUDA(cudaMemcpyAsync(err_out, R->d_io[1], ob, cudaMemcpyDeviceToHost, R->stream));
  TQEC_CUDA(cudaStreamSynchronize(R->stream));
  return TQEC_OK;
}

static int upload_probs(int model, int n_sites, const double *p0, const double *p1, const double *p2, double **d_p,
                        cudaStream_t stream) {
  const int np = model == TQEC_MODEL_DEPOL ? 3 : 1;
  TQEC_CUDA(cudaMalloc((void **)d_p, (size_t)np * (n_sites ? n_sites : 1) * 8));
  TQEC_CUDA(cudaMemcpyAsync(*d_p, p0, (size_t)n_sites * 8, cudaMemcpyHostToDevice, stream));
  if (np == 3) {
    TQEC_CUDA(cudaMemcpyAsync(*d_p + n_sites, p1, (size_t)n_sites * 8, cudaMemcpyHostToDevice, stream));
    TQEC_CUDA(cudaMemcpyAsync(*d_p + 2 * n_sites, p2, (size_t)n_sites * 8, cudaMemcpyHostToDevice, stream));
  }
  return TQEC_OK;
}

extern "C" int tqec_sample_errors(int32_t model, int32_t n_sites, const double *p0, const double *p1, const double *p2,
                                  uint64_t seed, int64_t shot_offset, int64_t B, uint64_t *err_out, int32_t device) {
  TQEC_REQUIRE(model == TQEC_MODEL_FLIP || model == TQEC_MODEL_DEPOL, "tqec_sample_errors: unknown model %d", model);
  TQEC_REQUIRE(n_sites >= 0 && p0 && (model == TQEC_MODEL_FLIP || (p1 && p2)), "tqec_sample_errors: NULL probability vector");
  TQEC_REQUIRE(B >= 0 && (B == 0 || err_out), "tqec_sample_errors: NULL output");
  if (B == 0) return TQEC_OK;
  int ndev = 0;
  TQEC_CUDA(cudaGetDeviceCount(&ndev));
  TQEC_REQUIRE(device >= 0 && device < ndev, "tqec_sample_errors: device %d not present (%d visible)", device, ndev);
  TQEC_CUDA(cudaSetDevice(device));
  const int words = words_for(model == TQEC_MODEL_DEPOL ? 2 * n_sites : n_sites);
  double *d_p = nullptr;
  uint64_t *d_err = nullptr;
  int rc = upload_probs(model, n_sites, p0, p1, p2, &d_p, 0);
  if (!rc && cudaMalloc((void **)&d_err, (size_t)B * words * 8) != cudaSuccess) { set_error("tqec_sample_errors: out of device memory"); rc = TQEC_ERR_NOMEM; }
  if (!rc) rc = launch_sample(model, n_sites, d_p, seed, shot_offset, B, d_err, words, 0);
  if (!rc && cudaMemcpy(err_out, d_err, (size_t)B * words * 8, cudaMemcpyDeviceToHost) != cudaSuccess) { set_error("tqec_sample_errors: copy back failed: %s", cudaGetErrorString(cudaGetLastError())); rc = TQEC_ERR_CUDA; }
  cudaFree(d_p);
  cudaFree(d_err);
  return rc;
}

// ---- fused pipeline -------------------------------------------------------------------------------------------
extern "C" int tqec_mc_run(const tqec_mc_desc *mc, uint64_t seed, int64_t shot_offset, int64_t n_shots, int64_t counts[4],
                           float *elapsed_ms) {
  tqec::NvtxRange nvtx_range("tqec_mc_run");
  TQEC_REQUIRE(mc && mc->plan && mc->H && mc->L && mc->row_class && counts, "tqec_mc_run: NULL argument");
  tqec_plan *P = mc->plan;
  TQEC_REQUIRE(P->semiring == TQEC_SEMIRING_MAXPLUS, "tqec_mc_run: needs a max-plus (TNMAP) plan");
  const int nbits = mc->model == TQEC_MODEL_DEPOL ? 2 * mc->n_sites : mc->n_sites;
  TQEC_REQUIRE(mc->model == TQEC_MODEL_FLIP || mc->model == TQEC_MODEL_DEPOL, "tqec_mc_run: unknown model %d", mc->model);
  TQEC_REQUIRE(nbits == P->dev.n_vars && mc->H->cols == nbits && mc->H->rows == P->dev.n_checks && mc->L->cols == nbits,
               "tqec_mc_run: shapes disagree (model bits %d, plan vars %d, H %dx%d, L %dx%d)", nbits, P->dev.n_vars,
               mc->H->rows, mc->H->cols, mc->L->rows, mc->L->cols);
  TQEC_REQUIRE(mc->H->device == P->device && mc->L->device == P->device, "tqec_mc_run: handles live on different devices");
  TQEC_REQUIRE(n_shots >= 0, "tqec_mc_run: negative shot count");
  TQEC_CUDA(cudaSetDevice(P->device));
  int64_t chunk = mc->chunk > 0 ? mc->chunk : (int64_t)1 << 20;
  if (mc->chunk <= 0 && P->has_sweep) {
    // whole rounds of k_sweep per chunk (148 SMs x teams x 32 shots): no half-empty last round in every chunk
    const int64_t per_round = (int64_t)P->sm_count * P->sw_teams * P->sw.grp;
    if (per_round > 0 && chunk > per_round) chunk = (chunk / per_round) * per_round;
  }
  if (chunk > n_shots) chunk = n_shots > 0 ? n_shots : 1;
  const int ew = P->dev.ncw, sw = P->dev.nsw;
  cudaStream_t st = P->stream;
  uint64_t *d_err = nullptr, *d_syn = nullptr, *d_cor = nullptr;
  double *d_p = nullptr;
  unsigned long long *d_counts = nullptr;
  int32_t *d_cls = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  int rc = TQEC_OK;
#define MC_TRY(call)                                                                          \
  do {                                                                                        \
    cudaError_t _e = (call);                                                                  \
    if (_e != cudaSuccess && rc == TQEC_OK) {                                                 \
      set_error("%s failed: %s", #call, cudaGetErrorString(_e));                              \
      rc = TQEC_ERR_CUDA;                                                                     \
    }                                                                                         \
  } while (0)
  // per-chunk scratch is owned by the plan and reused across calls (allocation used to cost more than a chunk's decode)
  {
    const size_t need = (size_t)chunk * (2 * ew + sw) * 8 + 64;
    if (P->mc_cap < need) {
      if (P->d_mc) cudaFree(P->d_mc);
      P->d_mc = nullptr; P->mc_cap = 0;
      MC_TRY(cudaMalloc(&P->d_mc, need));
      if (rc == TQEC_OK) P->mc_cap = need;
    }
    if (rc == TQEC_OK) {
      d_counts = (unsigned long long *)P->d_mc;
      d_err = (uint64_t *)((char *)P->d_mc + 64);
      d_syn = d_err + (size_t)chunk * ew;
      d_cor = d_syn + (size_t)chunk * sw;
    }
  }
  MC_TRY(cudaMalloc((void **)&d_cls, (size_t)(mc->L->rows ? mc->L->rows : 1) * 4));
  MC_TRY(cudaEventCreate(&e0));
  MC_TRY(cudaEventCreate(&e1));
  if (rc == TQEC_OK) rc = upload_probs(mc->model, mc->n_sites, mc->p0, mc->p1, mc->p2, &d_p, st);
  if (rc == TQEC_OK) {
    MC_TRY(cudaMemsetAsync(d_counts, 0, 32, st));
    MC_TRY(cudaMemcpyAsync(d_cls, mc->row_class, (size_t)mc->L->rows * 4, cudaMemcpyHostToDevice, st));
    MC_TRY(cudaEventRecord(e0, st));
  }
  for (int64_t done = 0; rc == TQEC_OK && done < n_shots; done += chunk) {
    const int64_t b = n_shots - done < chunk ? n_shots - done : chunk;
    rc = launch_sample(mc->model, mc->n_sites, d_p, seed, shot_offset + done, b, d_err, ew, st);
    if (!rc) rc = launch_gf2_apply(mc->H, d_err, b, d_syn, st);
    if (!rc) rc = launch_decode(P, d_syn, b, d_cor, nullptr, nullptr, st);
    if (!rc) rc = launch_flags(mc->L, d_cls, d_err, d_cor, b, nullptr, d_counts, st);
  }
  unsigned long long h[4] = {0, 0, 0, 0};
  if (rc == TQEC_OK && mc->comm) {
    // the one collective: device counters {x, z, any, shots} summed over the ranks, on the pipeline's stream
    const unsigned long long ns = (unsigned long long)n_shots;
    MC_TRY(cudaMemcpyAsync(d_counts + 3, &ns, 8, cudaMemcpyHostToDevice, st));
    if (rc == TQEC_OK) rc = comm_allreduce_dev(mc->comm, d_counts, st);
  }
  if (rc == TQEC_OK) {
    MC_TRY(cudaEventRecord(e1, st));
    MC_TRY(cudaMemcpyAsync(h, d_counts, 32, cudaMemcpyDeviceToHost, st));
    MC_TRY(cudaStreamSynchronize(st));
    if (rc == TQEC_OK && elapsed_ms) MC_TRY(cudaEventElapsedTime(elapsed_ms, e0, e1));
  }
#undef MC_TRY
  if (rc == TQEC_OK) {
    counts[0] += (int64_t)h[0]; counts[1] += (int64_t)h[1]; counts[2] += (int64_t)h[2];
    counts[3] += mc->comm ? (int64_t)h[3] : n_shots;
  }
  cudaFree(d_cls); cudaFree(d_p);
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  return rc;
}

// ---- FP64-pipe peak (roofline denominator of the max-plus / sum-product kernels) ------------------------------------
// MEASURED_PEAKS.json carries HBM and bf16 tensor peaks only; the decode kernels are bound by the FP64 CUDA-core pipe
// (DADD + DSETP for max-plus, DFMA for sum-product), so the denominator is measured here: register-resident,
// 8 independent dependency chains per thread, every SM filled.
template <int MODE>
__global__ void k_fp64_peak(double *out, int iters, double seed) {
  double a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 1e-9 + i;
  const double b = seed * 1e-6 + 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) a[i] = a[i] + b;                 // DADD
      else if (MODE == 1) a[i] = fma(a[i], b, c);     // DFMA
      else a[i] = (a[i] + b > a[(i + 1) & 7]) ? a[i] + b : a[(i + 1) & 7];  // DADD + DSETP + select (max-plus candidate)
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  if (s == 12345.678) out[0] = s;
}

extern "C" int tqec_fp64_peak(int32_t device, double *dadd_tops, double *dfma_tflops, double *maxplus_tops) {
  TQEC_REQUIRE(dadd_tops && dfma_tflops && maxplus_tops, "tqec_fp64_peak: NULL output");
  int ndev = 0;
  TQEC_CUDA(cudaGetDeviceCount(&ndev));
  TQEC_REQUIRE(device >= 0 && device < ndev, "tqec_fp64_peak: device %d not present (%d visible)", device, ndev);
  TQEC_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  TQEC_CUDA(cudaGetDeviceProperties(&prop, device));
  double *d_out = nullptr;
  TQEC_CUDA(cudaMalloc((void **)&d_out, 64));
  cudaEvent_t e0, e1;
  TQEC_CUDA(cudaEventCreate(&e0));
  TQEC_CUDA(cudaEventCreate(&e1));
  const int threads = 512, blocks = prop.multiProcessorCount * 4, iters = 1 << 14;
  double res[3] = {0, 0, 0};
  for (int mode = 0; mode < 3; ++mode) {
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      TQEC_CUDA(cudaEventRecord(e0, 0));
      if (mode == 0) k_fp64_peak<0><<<blocks, threads>>>(d_out, iters, 1.0);
      else if (mode == 1) k_fp64_peak<1><<<blocks, threads>>>(d_out, iters, 1.0);
      else k_fp64_peak<2><<<blocks, threads>>>(d_out, iters, 1.0);
      TQEC_CUDA(cudaEventRecord(e1, 0));
      TQEC_CUDA(cudaEventSynchronize(e1));
      float ms = 0;
      TQEC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      if (rep > 0 && ms < best) best = ms;
    }
    const double ops = (double)blocks * threads * iters * 8.0;
    res[mode] = ops / (best * 1e-3) / 1e12;
  }
  *dadd_tops = res[0];
  *dfma_tflops = 2.0 * res[1];
  *maxplus_tops = 2.0 * res[2];      // one add + one compare per candidate
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_out);
  return TQEC_OK;
}

// ---- FP64 tensor-core (DMMA) rate: the other possible roof of the sum-product steps ---------------------------------------
// mma.sync.aligned.m8n8k4.row.col.f64 (SASS DMMA.8x8x4): 512 flops per warp instruction, four independent accumulator
// chains per warp, register resident.  Used by DESIGN.md to settle whether any step of the frontier schedule should
// run as a dense product on the tensor cores (north_star: "FP64 DMMA ... only on steps that are dense GEMMs of useful
// size"); there is no tcgen05 FP64 path on sm_100a.
__global__ void k_dmma_peak(double *out, int iters, double seed) {
  double a = seed + threadIdx.x * 1e-9, b = 1.0000001;
  double c[4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) { c[i][0] = seed * i; c[i][1] = seed + i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;
}

extern "C" int tqec_dmma_peak(int32_t device, double *dmma_tflops) {
  TQEC_REQUIRE(dmma_tflops, "tqec_dmma_peak: NULL output");
  int ndev = 0;
  TQEC_CUDA(cudaGetDeviceCount(&ndev));
  TQEC_REQUIRE(device >= 0 && device < ndev, "tqec_dmma_peak: device %d not present (%d visible)", device, ndev);
  TQEC_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  TQEC_CUDA(cudaGetDeviceProperties(&prop, device));
  double *d_out = nullptr;
  TQEC_CUDA(cudaMalloc((void **)&d_out, 64));
  cudaEvent_t e0, e1;
  TQEC_CUDA(cudaEventCreate(&e0));
  TQEC_CUDA(cudaEventCreate(&e1));
  const int threads = 512, blocks = prop.multiProcessorCount * 4, iters = 1 << 13;
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    TQEC_CUDA(cudaEventRecord(e0, 0));
    k_dmma_peak<<<blocks, threads>>>(d_out, iters, 1.0);
    TQEC_CUDA(cudaEventRecord(e1, 0));
    TQEC_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    TQEC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  const double flops = (double)blocks * (threads / 32) * iters * 4.0 * 512.0;
  *dmma_tflops = flops / (best * 1e-3) / 1e12;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_out);
  return TQEC_OK;
}

// ---- lookup-table decoder (the reference's TableDecoder, src/decoding/truthtable.jl) -------------------------------------
// The table maps syndromes to error patterns; keys are sorted (most significant word last compares first) and a shot is
// decoded by binary search: HBM / L2 latency bound, log2(entries) dependent loads per shot.
struct tqec_table {
  int device, nsw, ncw;
  int64_t n;
  uint64_t *d_keys, *d_vals;
  void *d_io[3];
  size_t io_cap[3];
  cudaStream_t stream;
  int64_t launches;
};

namespace tqec {
__device__ __forceinline__ int key_cmp(const uint64_t *a, const uint64_t *b, int nsw) {
  for (int w = nsw - 1; w >= 0; --w) {
    if (a[w] < b[w]) return -1;
    if (a[w] > b[w]) return 1;
  }
  return 0;
}
__global__ void k_table_lookup(const uint64_t *__restrict__ keys, const uint64_t *__restrict__ vals, int64_t n, int nsw, int ncw,
                               const uint64_t *__restrict__ synd, int64_t B, uint64_t *__restrict__ corr, uint8_t *__restrict__ found) {
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < B; s += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t *q = synd + s * nsw;
    int64_t lo = 0, hi = n - 1, at = -1;
    while (lo <= hi) {
      const int64_t mid = (lo + hi) >> 1;
      const int c = key_cmp(keys + mid * nsw, q, nsw);
      if (c == 0) { at = mid; break; }
      if (c < 0) lo = mid + 1; else hi = mid - 1;
    }
    for (int w = 0; w < ncw; ++w) corr[s * ncw + w] = at >= 0 ? vals[at * ncw + w] : 0ull;
    if (found) found[s] = at >= 0;
  }
}
}  // namespace tqec

extern "C" int tqec_table_create(int64_t n_entries, int32_t n_checks, int32_t n_vars, const uint64_t *keys_sorted,
                                 const uint64_t *values, int32_t device, tqec_table **out) {
  TQEC_REQUIRE(out && n_entries >= 0 && n_checks >= 0 && n_vars >= 0 && (n_entries == 0 || (keys_sorted && values)),
               "tqec_table_create: bad arguments");
  *out = nullptr;
  const int nsw = words_for(n_checks), ncw = words_for(n_vars);
  for (int64_t i = 1; i < n_entries; ++i) {
    int c = 0;
    for (int w = nsw - 1; w >= 0 && c == 0; --w)
      c = keys_sorted[(i - 1) * nsw + w] < keys_sorted[i * nsw + w] ? -1 : (keys_sorted[(i - 1) * nsw + w] > keys_sorted[i * nsw + w] ? 1 : 0);
    TQEC_REQUIRE(c < 0, "tqec_table_create: keys must be strictly increasing (entry %lld)", (long long)i);
  }
  int ndev = 0;
  TQEC_CUDA(cudaGetDeviceCount(&ndev));
  TQEC_REQUIRE(device >= 0 && device < ndev, "tqec_table_create: device %d not present (%d visible)", device, ndev);
  TQEC_CUDA(cudaSetDevice(device));
  tqec_table *t = new tqec_table();
  std::memset(t, 0, sizeof(*t));
  t->device = device; t->nsw = nsw; t->ncw = ncw; t->n = n_entries;
  cudaError_t e = cudaMalloc((void **)&t->d_keys, (size_t)(n_entries ? n_entries : 1) * nsw * 8);
  if (e == cudaSuccess) e = cudaMalloc((void **)&t->d_vals, (size_t)(n_entries ? n_entries : 1) * ncw * 8);
  if (e == cudaSuccess && n_entries) e = cudaMemcpy(t->d_keys, keys_sorted, (size_t)n_entries * nsw * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && n_entries) e = cudaMemcpy(t->d_vals, values, (size_t)n_entries * ncw * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { set_error("tqec_table_create: %s", cudaGetErrorString(e)); cudaFree(t->d_keys); cudaFree(t->d_vals); delete t; return TQEC_ERR_CUDA; }
  *out = t;
  return TQEC_OK;
}

extern "C" int tqec_table_destroy(tqec_table *t) {
  if (!t) return TQEC_OK;
  cudaSetDevice(t->device);
  cudaFree(t->d_keys); cudaFree(t->d_vals);
  for (int i = 0; i < 3; ++i) cudaFree(t->d_io[i]);
  if (t->stream) cudaStreamDestroy(t->stream);
  delete t;
  return TQEC_OK;
}

extern "C" int tqec_table_decode(tqec_table *t, const uint64_t *synd, int64_t B, uint64_t *corr_out, uint8_t *found_out) {
  TQEC_REQUIRE(t && B >= 0 && (B == 0 || (synd && corr_out)), "tqec_table_decode: NULL argument");
  if (B == 0) return TQEC_OK;
  TQEC_CUDA(cudaSetDevice(t->device));
  int rc;
  if ((rc = ensure_cap(&t->d_io[0], &t->io_cap[0], (size_t)B * t->nsw * 8))) return rc;
  if ((rc = ensure_cap(&t->d_io[1], &t->io_cap[1], (size_t)B * t->ncw * 8))) return rc;
  if ((rc = ensure_cap(&t->d_io[2], &t->io_cap[2], (size_t)B))) return rc;
  TQEC_CUDA(cudaMemcpyAsync(t->d_io[0], synd, (size_t)B * t->nsw * 8, cudaMemcpyHostToDevice, t->stream));
  const int64_t want = (B + 255) / 256;
  k_table_lookup<<<(unsigned)(want < 4096 ? want : 4096), 256, 0, t->stream>>>(t->d_keys, t->d_vals, t->n, t->nsw, t->ncw,
                                                                              (const uint64_t *)t->d_io[0], B, (uint64_t *)t->d_io[1],
                                                                              (uint8_t *)t->d_io[2]);
  TQEC_CUDA(cudaGetLastError());
  t->launches += 1;
  TQEC_CUDA(cudaMemcpyAsync(corr_out, t->d_io[1], (size_t)B * t->ncw * 8, cudaMemcpyDeviceToHost, t->stream));
  if (found_out) TQEC_CUDA(cudaMemcpyAsync(found_out, t->d_io[2], (size_t)B, cudaMemcpyDeviceToHost, t->stream));
  TQEC_CUDA(cudaStreamSynchronize(t->stream));
  return TQEC_OK;
}
