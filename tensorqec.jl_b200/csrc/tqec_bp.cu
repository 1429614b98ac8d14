// Belief propagation + ordered-statistics decoding, batched: one thread per shot (a baseline decoder
// sharing the GF(2) front / back end of the tensor-network decoders).
// Reference: src/decoding/bposd.jl -- `belief_propagation` (:55-78: tanh-rule sum-product on log-likelihood ratios,
// flooding schedule = all checks, then all bits; bit-to-check messages clamped to [-10, 10]; stop as soon as the hard
// decision reproduces the syndrome), `osd` (:80-97: order-0 OSD -- bits in increasing reliability order, the first
// linearly independent columns of H form an invertible system that is solved for the syndrome).
// Layout: messages live in global memory as [edge][shot] (a warp's 32 shots read 32 consecutive doubles per edge); the
// Tanner graph (CSR both ways) is read-only and L1-resident.  The OSD stage keeps one bit-row per check in local memory.
#include <cstring>

#include "tqec_common.h"

#define BP_MAX_ROW_WORDS 8      // bits (variables) <= 512
#define BP_MAX_CHECKS 256

struct tqec_bp {
  int device, nq, ns, n_edges, max_iter, osd;
  int32_t *d_s_ptr, *d_s_adj;        // check -> edges: s_ptr[ns + 1], s_adj[e] = bit of edge e (edges are numbered check-major)
  int32_t *d_q_ptr, *d_q_edge;       // bit -> edges: q_ptr[nq + 1], q_edge[k] = edge id
  double *d_mu;                      // prior log-likelihood ratios log((1 - p) / p)
  double *d_msg;                     // 2 * n_edges * capacity doubles: bit-to-check, then check-to-bit messages
  size_t msg_cap;
  void *d_io[3];
  size_t io_cap[3];
  cudaStream_t stream;
};

namespace tqec {

__global__ void k_bp_osd(const tqec_bp P, const uint64_t *__restrict__ synd, int64_t B, uint64_t *__restrict__ corr,
                         uint8_t *__restrict__ flags, double *__restrict__ msg) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int nsw = (P.ns + 63) / 64 > 0 ? (P.ns + 63) / 64 : 1, ncw = (P.nq + 63) / 64 > 0 ? (P.nq + 63) / 64 : 1;
  double *mq2s = msg + b, *ms2q = msg + (size_t)P.n_edges * B + b;      // element e at [e * B]
  const uint64_t *syn = synd + b * nsw;
  for (int q = 0; q < P.nq; ++q)
    for (int k = P.d_q_ptr[q]; k < P.d_q_ptr[q + 1]; ++k) mq2s[(size_t)P.d_q_edge[k] * B] = P.d_mu[q];
  uint64_t hard[BP_MAX_ROW_WORDS];
  bool ok = false;
  for (int it = 0; it < P.max_iter && !ok; ++it) {
    for (int s = 0; s < P.ns; ++s) {
      double pro = 1.0;
      for (int e = P.d_s_ptr[s]; e < P.d_s_ptr[s + 1]; ++e) pro *= tanh(mq2s[(size_t)e * B]);
      const double sgn = ((syn[s >> 6] >> (s & 63)) & 1ull) ? -1.0 : 1.0;
      for (int e = P.d_s_ptr[s]; e < P.d_s_ptr[s + 1]; ++e) ms2q[(size_t)e * B] = sgn * atanh(pro / tanh(mq2s[(size_t)e * B]));
    }
    for (int w = 0; w < ncw; ++w) hard[w] = 0;
    for (int q = 0; q < P.nq; ++q) {
      double qv = 0.0;
      for (int k = P.d_q_ptr[q]; k < P.d_q_ptr[q + 1]; ++k) qv += ms2q[(size_t)P.d_q_edge[k] * B];
      qv += P.d_mu[q];
      for (int k = P.d_q_ptr[q]; k < P.d_q_ptr[q + 1]; ++k) {
        const size_t e = (size_t)P.d_q_edge[k] * B;
        mq2s[e] = fmax(fmin(qv - ms2q[e], 10.0), -10.0);
      }
      if (qv < 0.0) hard[q >> 6] |= 1ull << (q & 63);
    }
    ok = true;
    for (int s = 0; s < P.ns && ok; ++s) {
      int par = 0;
      for (int e = P.d_s_ptr[s]; e < P.d_s_ptr[s + 1]; ++e) par ^= (int)((hard[P.d_s_adj[e] >> 6] >> (P.d_s_adj[e] & 63)) & 1ull);
      ok = par == (int)((syn[s >> 6] >> (s & 63)) & 1ull);
    }
  }
  uint8_t flag = ok ? 1 : 0;                                      // bit 0: BP converged; bit 1: OSD produced the pattern
  if (!ok) {
    for (int w = 0; w < ncw; ++w) hard[w] = 0;
    if (P.osd) {
      // reliability order: increasing q_vec (sortperm, bposd.jl:73, 77).  q_vec of the last iteration is rebuilt from the
      // final messages; ties keep the lower index (stable), as sortperm does.
      uint64_t rows[BP_MAX_CHECKS][BP_MAX_ROW_WORDS];             // one bit-row per check (local memory)
      uint8_t rhs[BP_MAX_CHECKS], used[BP_MAX_CHECKS];
      for (int s = 0; s < P.ns; ++s) {
        for (int w = 0; w < ncw; ++w) rows[s][w] = 0;
        for (int e = P.d_s_ptr[s]; e < P.d_s_ptr[s + 1]; ++e) rows[s][P.d_s_adj[e] >> 6] |= 1ull << (P.d_s_adj[e] & 63);
        rhs[s] = (uint8_t)((syn[s >> 6] >> (s & 63)) & 1ull);
        used[s] = 0;
      }
      // selection sort over the bits by (q_vec, index), one bit per step; each selected bit is a candidate pivot column
      uint64_t taken[BP_MAX_ROW_WORDS];
      for (int w = 0; w < ncw; ++w) taken[w] = 0;
      int16_t piv_col[BP_MAX_CHECKS];
      int n_piv = 0;
      for (int step = 0; step < P.nq && n_piv < P.ns; ++step) {
        int best = -1;
        double bv = 0.0;
        for (int q = 0; q < P.nq; ++q) {
          if ((taken[q >> 6] >> (q & 63)) & 1ull) continue;
          double qv = P.d_mu[q];
          for (int k = P.d_q_ptr[q]; k < P.d_q_ptr[q + 1]; ++k) qv += ms2q[(size_t)P.d_q_edge[k] * B];
          if (best < 0 || qv < bv) { best = q; bv = qv; }
        }
        taken[best >> 6] |= 1ull << (best & 63);
        int r = -1;
        for (int s = 0; s < P.ns; ++s)
          if (!used[s] && ((rows[s][best >> 6] >> (best & 63)) & 1ull)) { r = s; break; }
        if (r < 0) continue;                                      // dependent on the columns already chosen
        used[r] = 1;
        piv_col[r] = (int16_t)best;
        ++n_piv;
        for (int s = 0; s < P.ns; ++s)
          if (s != r && ((rows[s][best >> 6] >> (best & 63)) & 1ull)) {
            for (int w = 0; w < ncw; ++w) rows[s][w] ^= rows[r][w];
            rhs[s] ^= rhs[r];
          }
      }
      for (int s = 0; s < P.ns; ++s)
        if (used[s] && rhs[s]) hard[piv_col[s] >> 6] |= 1ull << (piv_col[s] & 63);
      flag |= 2;
    }
  }
  for (int w = 0; w < ncw; ++w) corr[b * ncw + w] = hard[w];
  if (flags) flags[b] = flag;
}

}  // namespace tqec

using namespace tqec;

extern "C" int tqec_bp_create(int32_t nq, int32_t ns, const int32_t *s_ptr, const int32_t *s_adj, const double *p,
                              int32_t max_iter, int32_t osd, int32_t device, tqec_bp **out) {
  TQEC_REQUIRE(out && s_ptr && s_adj && p && nq >= 1 && ns >= 1 && max_iter >= 1, "tqec_bp_create: bad arguments");
  *out = nullptr;
  TQEC_REQUIRE(nq <= 64 * BP_MAX_ROW_WORDS && ns <= BP_MAX_CHECKS, "tqec_bp_create: at most %d bits and %d checks", 64 * BP_MAX_ROW_WORDS, BP_MAX_CHECKS);
  const int ne = s_ptr[ns];
  std::vector<std::vector<int32_t>> q_edges(nq);
  for (int s = 0; s < ns; ++s) {
    TQEC_REQUIRE(s_ptr[s] <= s_ptr[s + 1], "tqec_bp_create: s_ptr must be non-decreasing");
    for (int e = s_ptr[s]; e < s_ptr[s + 1]; ++e) {
      TQEC_REQUIRE(s_adj[e] >= 0 && s_adj[e] < nq, "tqec_bp_create: bit index out of range");
      q_edges[s_adj[e]].push_back(e);
    }
  }
  std::vector<int32_t> q_ptr(nq + 1, 0), q_edge;
  std::vector<double> mu(nq);
  for (int q = 0; q < nq; ++q) {
    TQEC_REQUIRE(p[q] > 0.0 && p[q] < 1.0, "tqec_bp_create: flip probabilities must lie in (0, 1)");
    mu[q] = std::log((1.0 - p[q]) / p[q]);
    q_edge.insert(q_edge.end(), q_edges[q].begin(), q_edges[q].end());
    q_ptr[q + 1] = (int32_t)q_edge.size();
  }
  int ndev = 0;
  TQEC_CUDA(cudaGetDeviceCount(&ndev));
  TQEC_REQUIRE(device >= 0 && device < ndev, "tqec_bp_create: device %d not present (%d visible)", device, ndev);
  TQEC_CUDA(cudaSetDevice(device));
  tqec_bp *t = new tqec_bp();
  std::memset(t, 0, sizeof(*t));
  t->device = device; t->nq = nq; t->ns = ns; t->n_edges = ne; t->max_iter = max_iter; t->osd = osd ? 1 : 0;
  cudaError_t e = cudaMalloc((void **)&t->d_s_ptr, (ns + 1) * 4);
  if (e == cudaSuccess) e = cudaMalloc((void **)&t->d_s_adj, (ne ? ne : 1) * 4);
  if (e == cudaSuccess) e = cudaMalloc((void **)&t->d_q_ptr, (nq + 1) * 4);
  if (e == cudaSuccess) e = cudaMalloc((void **)&t->d_q_edge, (ne ? ne : 1) * 4);
  if (e == cudaSuccess) e = cudaMalloc((void **)&t->d_mu, nq * 8);
  if (e == cudaSuccess) e = cudaMemcpy(t->d_s_ptr, s_ptr, (ns + 1) * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(t->d_s_adj, s_adj, ne * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(t->d_q_ptr, q_ptr.data(), (nq + 1) * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(t->d_q_edge, q_edge.data(), ne * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(t->d_mu, mu.data(), nq * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { set_error("tqec_bp_create: %s", cudaGetErrorString(e)); tqec_bp_destroy(t); return TQEC_ERR_CUDA; }
  *out = t;
  return TQEC_OK;
}

extern "C" int tqec_bp_destroy(tqec_bp *t) {
  if (!t) return TQEC_OK;
  cudaSetDevice(t->device);
  cudaFree(t->d_s_ptr); cudaFree(t->d_s_adj); cudaFree(t->d_q_ptr); cudaFree(t->d_q_edge); cudaFree(t->d_mu); cudaFree(t->d_msg);
  for (int i = 0; i < 3; ++i) cudaFree(t->d_io[i]);
  if (t->stream) cudaStreamDestroy(t->stream);
  delete t;
  return TQEC_OK;
}

extern "C" int tqec_bp_decode(tqec_bp *t, const uint64_t *synd, int64_t B, uint64_t *corr_out, uint8_t *flags_out) {
  TQEC_REQUIRE(t && B >= 0 && (B == 0 || (synd && corr_out)), "tqec_bp_decode: NULL argument");
  if (B == 0) return TQEC_OK;
  TQEC_CUDA(cudaSetDevice(t->device));
  const int nsw = words_for(t->ns), ncw = words_for(t->nq);
  const int64_t CH = (int64_t)1 << 18;                            // shots per launch: bounds the message scratch
  const int64_t nb = B < CH ? B : CH;
  int rc;
  if ((rc = ensure_cap(&t->d_io[0], &t->io_cap[0], (size_t)nb * nsw * 8))) return rc;
  if ((rc = ensure_cap(&t->d_io[1], &t->io_cap[1], (size_t)nb * ncw * 8))) return rc;
  if ((rc = ensure_cap(&t->d_io[2], &t->io_cap[2], (size_t)nb))) return rc;
  if ((rc = ensure_cap((void **)&t->d_msg, &t->msg_cap, (size_t)2 * t->n_edges * nb * 8))) return rc;
  // the OSD rows live in per-thread local memory: make room for them
  size_t need = (size_t)BP_MAX_CHECKS * BP_MAX_ROW_WORDS * 8 + 4096, have = 0;
  cudaDeviceGetLimit(&have, cudaLimitStackSize);
  if (have < need) TQEC_CUDA(cudaDeviceSetLimit(cudaLimitStackSize, need));
  for (int64_t o = 0; o < B; o += nb) {
    const int64_t n = B - o < nb ? B - o : nb;
    TQEC_CUDA(cudaMemcpyAsync(t->d_io[0], synd + o * nsw, (size_t)n * nsw * 8, cudaMemcpyHostToDevice, t->stream));
    k_bp_osd<<<(unsigned)((n + 63) / 64), 64, 0, t->stream>>>(*t, (const uint64_t *)t->d_io[0], n, (uint64_t *)t->d_io[1],
                                                               (uint8_t *)t->d_io[2], t->d_msg);
    TQEC_CUDA(cudaGetLastError());
    TQEC_CUDA(cudaMemcpyAsync(corr_out + o * ncw, t->d_io[1], (size_t)n * ncw * 8, cudaMemcpyDeviceToHost, t->stream));
    if (flags_out) TQEC_CUDA(cudaMemcpyAsync(flags_out + o, t->d_io[2], (size_t)n, cudaMemcpyDeviceToHost, t->stream));
    TQEC_CUDA(cudaStreamSynchronize(t->stream));
  }
  return TQEC_OK;
}
