// Host-side lowering, part 2: the in-place patch sweep executed by k_sweep (see tqec_lower.h).
// Mirrors tensorqec.jl_b200/sweep.py decision for decision (that file carries the full description).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <stdexcept>
#include <system_error>
#include <thread>

#include "tqec_lower.h"
#include "tqec_sweep_menu.h"

namespace tqec {
namespace lower {

static const int NB = 10;
static const int MAX_PATCH = 4;
static const int REC_INTS = 32;
static const int TB_INTS = 64;

static bool has(const std::vector<int> &v, int x) { return std::find(v.begin(), v.end(), x) != v.end(); }
static uint32_t phys(uint32_t x) { return x ^ ((x >> 4) & 15u); }

// descriptor of a super-step shape: (M, layers -> [pinned bits, free masks(, extra flip masks)]), ordered like the
// nested tuples of sweep.py (element-wise, a proper prefix sorts first)
struct Desc {
  int M = 0;
  std::vector<std::vector<std::vector<int>>> layers;
  bool operator<(const Desc &o) const { return M != o.M ? M < o.M : layers < o.layers; }
  bool operator==(const Desc &o) const { return M == o.M && layers == o.layers; }
};

static std::vector<Desc> build_menu() {
  std::vector<Desc> menu;
#define SW_MENU_ROW(ID, M_, NL, NP0, P00, P01, NF0, F00, F01, K00, K01, NP1, P10, P11, NF1, F10, F11, K10, K11)              \
  {                                                                                                                       \
    Desc d;                                                                                                               \
    d.M = M_;                                                                                                             \
    const int np[2] = {NP0, NP1}, nf[2] = {NF0, NF1};                                                                      \
    const int P[2][2] = {{P00, P01}, {P10, P11}}, F[2][2] = {{F00, F01}, {F10, F11}}, K[2][2] = {{K00, K01}, {K10, K11}};  \
    for (int l = 0; l < NL; ++l) {                                                                                        \
      std::vector<int> pb, fm, pm;                                                                                        \
      bool anyk = false;                                                                                                  \
      for (int q = 0; q < np[l]; ++q) { pb.push_back(P[l][q]); pm.push_back(K[l][q]); anyk = anyk || K[l][q] != 0; }       \
      for (int q = 0; q < nf[l]; ++q) fm.push_back(F[l][q]);                                                               \
      if (anyk) d.layers.push_back({pb, fm, pm}); else d.layers.push_back({pb, fm});                                      \
    }                                                                                                                     \
    menu.push_back(d);                                                                                                    \
  }
  TQEC_SWEEP_MENU(SW_MENU_ROW)
#undef SW_MENU_ROW
  return menu;
}

struct Role { int fi; std::vector<int> touched, opened, closing; };
struct PinnedC { int j, o, c; std::vector<int> extra; };
struct FreeC { int j; std::vector<int> tv; };
struct Classified { bool ok = false; std::vector<PinnedC> pinned; std::vector<FreeC> free; };

struct Raw {
  int step = 0, fi = 0;
  std::vector<std::pair<int, int>> pinned;                 // (variable index j, chain)
  std::vector<std::vector<int>> pk;                        // per pinned variable: other chains it flips
  std::vector<std::pair<int, std::vector<int>>> free;      // (variable index j, chains)
  std::vector<std::pair<int, int>> closed;                 // (syndrome bit, chain)
  std::vector<int> chains;
  std::vector<int> fresh, freed;                           // chains that come alive (fresh pins) / die in this step
};

struct Layer {
  int step = 0, factor = 0;
  std::vector<int> vars;
  std::vector<std::pair<int, int>> pinned, free, closed;   // (j, patch bit) / (j, flip mask) / (syndrome bit, patch bit)
  std::vector<double> T;
  std::vector<int> pk;
};

struct SuperStep {
  std::vector<Layer> layers;
  std::vector<int> chains, pos, lanepos, looppos;
  int menu = -1;
  bool conflict = false, has_late = false;
  int late_sb = 0, late_bit = 0;
  int wbase = 0, bpp = 0, n_words = 0;
  std::vector<std::vector<int>> fresh;                     // per layer: chains that come alive
};

static bool check_has(const Check &c, int v) { return std::find(c.vars.begin(), c.vars.end(), v) != c.vars.end(); }

static Classified classify(const Factor &f, const Role &R, const std::vector<Check> &checks) {
  Classified out;
  std::set<int> opened(R.opened.begin(), R.opened.end()), closing(R.closing.begin(), R.closing.end());
  for (int c : opened)
    if (closing.count(c)) return out;
  std::set<int> donors;
  for (size_t j = 0; j < f.vars.size(); ++j) {
    const int v = f.vars[j];
    std::vector<int> tv, ov;
    for (int c : R.touched)
      if (check_has(checks[c], v)) tv.push_back(c);
    for (int c : tv)
      if (opened.count(c)) ov.push_back(c);
    if (ov.empty()) {
      out.free.push_back({(int)j, tv});
      continue;
    }
    if (ov.size() != 1) return out;
    int cnt = 0;
    for (int w : f.vars)
      if (check_has(checks[ov[0]], w)) ++cnt;
    if (cnt != 1) return out;
    std::vector<int> rest;
    for (int c : tv)
      if (c != ov[0]) rest.push_back(c);
    int donor = -1;
    for (int c : rest)
      if (closing.count(c) && !donors.count(c)) { donor = c; break; }
    PinnedC p;
    p.j = (int)j; p.o = ov[0]; p.c = donor;
    if (donor < 0) {
      // FRESH pin: the opened check takes a dead slot; its own slot joins the flip mask of the pinned variable, so that
      // output bit 1 reads the live (bit 0) half of the slot
      p.extra = rest;
      out.pinned.push_back(p);
      continue;
    }
    donors.insert(donor);
    for (int c : rest)
      if (c != donor) p.extra.push_back(c);
    out.pinned.push_back(p);
  }
  if (out.pinned.size() > 2 || out.free.size() > 2) return out;
  out.ok = true;
  return out;
}

static Desc descriptor(const std::vector<const Raw *> &group, const std::vector<int> &chain_order) {
  std::map<int, int> bit;
  for (size_t b = 0; b < chain_order.size(); ++b) bit[chain_order[b]] = (int)b;
  Desc d;
  d.M = (int)chain_order.size();
  for (const Raw *g : group) {
    std::vector<int> pb, fm, pm;
    bool anyk = false;
    for (auto &p : g->pinned) pb.push_back(bit[p.second]);
    for (auto &fr : g->free) {
      int m = 0;
      for (int ch : fr.second) m += 1 << bit[ch];
      fm.push_back(m);
    }
    for (auto &chs : g->pk) {
      int m = 0;
      for (int ch : chs) m += 1 << bit[ch];
      pm.push_back(m);
      anyk = anyk || m != 0;
    }
    if (anyk) d.layers.push_back({pb, fm, pm}); else d.layers.push_back({pb, fm});
  }
  return d;
}

// Split [0, n) over a few host threads when the range is large (head tables of 2^20+ entries); every index is computed
// independently of the others, so the result does not depend on the split.
template <typename F>
static void parallel_ranges(size_t n, F &&fn) {
  unsigned hw = std::thread::hardware_concurrency();
  size_t parts = n >> 16;
  if (parts > 16) parts = 16;
  if (hw && parts > hw) parts = hw;
  if (parts <= 1) { fn((size_t)0, n); return; }
  const size_t step = (n + parts - 1) / parts;
  std::vector<std::thread> th;
  for (size_t k = 1; k < parts; ++k) {
    const size_t lo = k * step, hi = std::min(n, lo + step);
    if (lo >= hi) continue;
    try {
      th.emplace_back([&fn, lo, hi] { fn(lo, hi); });
    } catch (const std::system_error &) {                        // no thread to be had: this range runs here
      fn(lo, hi);
    }
  }
  fn((size_t)0, std::min(n, step));
  for (auto &t : th) t.join();
}

// Tabulate the first h steps for every value of the syndrome bits they close (sweep.py:_head_eval).  Internal layout:
// one index bit per axis in creation order; closing a check only changes the role of its axis (state axis -> batch
// axis of that syndrome bit), no data moves.  Output in POSITION order: [head pattern][state index], index bit
// chain_pos[k] = parity of check live_order[k].
static void head_eval(const Schedule &sch, int h, const std::vector<Role> &roles, const std::vector<int> &head_bits,
                      const std::vector<int> &live_order, const std::vector<int> &chain_pos, int Wt, std::vector<double> &hs,
                      std::vector<uint64_t> &hc) {
  const bool maxplus = sch.semiring == TQEC_SEMIRING_MAXPLUS;
  const int ncw = std::max(1, (sch.n_vars + 63) / 64);
  std::vector<int> axis_check;   // check id of index bit p
  std::vector<char> axis_batch;  // 1 once the check has been closed
  std::vector<int> batch_order;  // index bits in closing order
  std::vector<double> St(1, maxplus ? 0.0 : 1.0);
  std::vector<uint64_t> cfg;
  if (maxplus) cfg.assign(ncw, 0);
  const double zero = maxplus ? -INFINITY : 0.0;
  std::vector<double> best;
  std::vector<uint64_t> bcfg;
  for (int t = 0; t < h; ++t) {
    const Role &R = roles[t];
    const Factor &f = sch.factors[R.fi];
    const std::vector<double> &T = sch.steps[t].table;
    for (int c : R.opened) {
      const size_t n = St.size();
      St.resize(2 * n, zero);
      if (maxplus) {
        cfg.resize(2 * n * ncw);
        std::memcpy(cfg.data() + n * ncw, cfg.data(), n * ncw * sizeof(uint64_t));
      }
      axis_check.push_back(c);
      axis_batch.push_back(0);
    }
    auto bit_of_check = [&](int c) {
      for (size_t p = 0; p < axis_check.size(); ++p)
        if (axis_check[p] == c && !axis_batch[p]) return (int)p;
      throw std::runtime_error("head_eval: check is not open");
    };
    const size_t n = St.size();
    if (best.size() < n) best.resize(n);                          // the two buffers of a step are kept and swapped: no fresh
    if (maxplus && bcfg.size() < n * ncw) bcfg.resize(n * ncw);   // 200 MB allocation (and its page faults) per step
    const int NA = 1 << f.vars.size();
    std::vector<size_t> flips(NA, 0);
    std::vector<uint64_t> amasks((size_t)NA * ncw, 0);
    for (int a = 0; a < NA; ++a) {
      for (int c : R.touched) {
        int p = 0;
        for (size_t j = 0; j < f.vars.size(); ++j)
          if (check_has(sch.checks[c], f.vars[j])) p ^= (a >> j) & 1;
        if (p) flips[a] |= (size_t)1 << bit_of_check(c);
      }
      if (maxplus)
        for (size_t j = 0; j < f.vars.size(); ++j)
          if ((a >> j) & 1) amasks[(size_t)a * ncw + (f.vars[j] >> 6)] |= (uint64_t)1 << (f.vars[j] & 63);
    }
    // per entry: candidates in ascending assignment order (the order of the additions and of the tie rule is per entry,
    // so the entries can be split over threads)
    parallel_ranges(n, [&](size_t i0, size_t i1) {
      for (int a = 0; a < NA; ++a) {
        const size_t flip = flips[a];
        const uint64_t *amask = amasks.data() + (size_t)a * ncw;
        const double ta = T[a];
        for (size_t i = i0; i < i1; ++i) {
          const size_t src = i ^ flip;
          const double cand = maxplus ? St[src] + ta : St[src] * ta;
          if (a == 0) {
            best[i] = cand;
            if (maxplus)
              for (int w = 0; w < ncw; ++w) bcfg[i * ncw + w] = cfg[src * ncw + w] | amask[w];
          } else if (maxplus) {
            if (cand > best[i]) {            // strict: the smallest assignment wins exact ties
              best[i] = cand;
              for (int w = 0; w < ncw; ++w) bcfg[i * ncw + w] = cfg[src * ncw + w] | amask[w];
            }
          } else {
            best[i] = best[i] + cand;
          }
        }
      }
    });
    best.resize(n);
    St.swap(best);
    if (maxplus) { bcfg.resize(n * ncw); cfg.swap(bcfg); }
    for (int c : R.closing) {
      const int p = bit_of_check(c);
      axis_batch[p] = 1;
      batch_order.push_back(p);
    }
  }
  const int nh = (int)head_bits.size(), W = (int)live_order.size();
  if ((int)batch_order.size() != nh) throw std::runtime_error("head_eval: closed bits do not match the head bits");
  for (int j = 0; j < nh; ++j)
    if (sch.checks[axis_check[batch_order[j]]].index != head_bits[j]) throw std::runtime_error("head_eval: head bit order");
  std::vector<int> state_bit(W);
  for (int k = 0; k < W; ++k) {
    int found = -1;
    for (size_t p = 0; p < axis_check.size(); ++p)
      if (axis_check[p] == live_order[k] && !axis_batch[p]) found = (int)p;
    if (found < 0) throw std::runtime_error("head_eval: live check missing");
    state_bit[k] = found;
  }
  if ((size_t)1 << (nh + W) != St.size()) throw std::runtime_error("head_eval: size mismatch");
  // Wt >= W chains in all: the ones added by fresh pins are dead after the head, their set halves hold the semiring's zero
  const size_t total = (size_t)1 << (nh + Wt);
  hs.assign(total, 0.0);
  if (maxplus) hc.assign(total * ncw, 0); else hc.assign(((size_t)1 << nh) * ncw, 0);
  uint32_t extra_mask = 0;
  for (int k = W; k < Wt; ++k) extra_mask |= 1u << chain_pos[k];
  parallel_ranges(total, [&](size_t e0, size_t e1) {
    for (size_t e = e0; e < e1; ++e) {
      const size_t hp = e >> Wt, idx = e & (((size_t)1 << Wt) - 1);
      if (idx & extra_mask) { hs[e] = zero; continue; }
      size_t src = 0;
      for (int j = 0; j < nh; ++j)
        if ((hp >> j) & 1) src |= (size_t)1 << batch_order[j];
      for (int k = 0; k < W; ++k)
        if ((idx >> chain_pos[k]) & 1) src |= (size_t)1 << state_bit[k];
      hs[e] = St[src];
      if (maxplus)
        for (int w = 0; w < ncw; ++w) hc[e * ncw + w] = cfg[src * ncw + w];
    }
  });
}

static std::pair<std::vector<int>, int> assign_positions(const std::vector<std::vector<int>> &groups, int W, uint64_t seed = 0) {
  auto cost = [&](const std::vector<int> &p) {
    int c = 0;
    for (auto &chains : groups) {
      bool ps[32] = {false};
      for (int ch : chains) ps[p[ch]] = true;
      bool hit = false;
      for (int q = 0; q < 4; ++q) hit = hit || (ps[q] && ps[q + 4]);
      if (hit) ++c;
    }
    return c;
  };
  uint64_t state = seed + 1;
  auto shuffled = [&]() {
    std::vector<int> q(W);
    for (int i = 0; i < W; ++i) q[i] = i;
    for (int i = W - 1; i > 0; --i) {
      state = state * 6364136223846793005ull + 1442695040888963407ull;
      const int j = (int)((state >> 33) % (uint64_t)(i + 1));
      std::swap(q[i], q[j]);
    }
    return q;
  };
  std::vector<int> best_p(W);
  for (int i = 0; i < W; ++i) best_p[i] = i;
  int best_c = -1;
  for (int restart = 0; restart < 40; ++restart) {
    std::vector<int> p;
    if (restart) p = shuffled();
    else { p.resize(W); for (int i = 0; i < W; ++i) p[i] = i; }
    int c = cost(p);
    bool improved = true;
    while (improved && c) {
      improved = false;
      for (int i = 0; i < W; ++i)
        for (int j = i + 1; j < W; ++j) {
          std::swap(p[i], p[j]);
          const int c2 = cost(p);
          if (c2 < c) { c = c2; improved = true; }
          else std::swap(p[i], p[j]);
        }
    }
    if (best_c < 0 || c < best_c) { best_p = p; best_c = c; }
    if (c == 0) break;
  }
  return {best_p, best_c};
}

static void encode_sweep(SweepPlan &p, const std::vector<SuperStep> &ssteps) {
  const int n = (int)ssteps.size();
  p.rec.assign((size_t)n * REC_INTS, 0);
  p.tb.assign((size_t)n * TB_INTS, 0);
  p.lanetab.assign((size_t)n * 32, 0);
  p.tvals.clear();
  const uint32_t submask = (((uint32_t)1 << p.sg) - 1u) << p.W;
  for (int i = 0; i < n; ++i) {
    const SuperStep &ss = ssteps[i];
    const int M = (int)ss.pos.size();
    int32_t *r = p.rec.data() + (size_t)i * REC_INTS;
    r[0] = ss.menu;
    r[1] = 1 << ss.looppos.size();
    while (p.tvals.size() % 2) p.tvals.push_back(0.0);
    r[2] = (int32_t)p.tvals.size();
    for (auto &l : ss.layers) p.tvals.insert(p.tvals.end(), l.T.begin(), l.T.end());
    int flipmask = 0;
    if (ss.has_late) {
      const Layer &l = ss.layers[1];
      const int NF = (int)l.free.size();
      for (size_t q = 0; q < l.pinned.size(); ++q)
        if (l.pinned[q].second == ss.late_bit) flipmask |= 1 << q;
      for (size_t ix = 0; ix < l.T.size(); ++ix)
        p.tvals.push_back(l.T[((((ix >> NF) ^ (size_t)flipmask) << NF) | (ix & (((size_t)1 << NF) - 1)))]);
    }
    r[3] = ss.wbase;
    int ain[4] = {0, 0, 0, 0};
    for (int b = 0; b < M; ++b) ain[b] = (int)(phys(1u << ss.pos[b]) << 3);
    r[4] = ain[0] | (ain[1] << 16);
    r[5] = ain[2] | (ain[3] << 16);
    r[6] = i * 32;
    r[7] = i << p.sg;
    int la[8], ls[8];
    for (int it = 0; it < 8; ++it) {
      uint32_t x = 0;
      for (size_t q = 0; q < ss.looppos.size(); ++q) x |= (uint32_t)((it >> q) & 1) << ss.looppos[q];
      const bool in = it < (1 << ss.looppos.size());
      la[it] = in ? (int)(phys(x) << 3) : 0;
      ls[it] = in ? (int)((x & submask) >> p.W) : 0;
    }
    for (int q = 0; q < 4; ++q) r[8 + q] = la[2 * q] | (la[2 * q + 1] << 16);
    r[12] = ls[0] | (ls[1] << 8) | (ls[2] << 16) | (ls[3] << 24);
    r[13] = ls[4] | (ls[5] << 8) | (ls[6] << 16) | (ls[7] << 24);
    std::vector<std::pair<int, int>> closed;   // (syndrome bit, position)
    for (auto &l : ss.layers)
      for (auto &c : l.closed) closed.push_back({c.first, ss.pos[c.second]});
    if (closed.size() > 4) throw std::runtime_error("sweep: more than four closed checks in a super-step");
    r[14] = (int32_t)closed.size();
    for (size_t q = 0; q < closed.size(); ++q) r[16 + q] = closed[q].first | (int32_t)(phys(1u << closed[q].second) << 19);
    r[20] = -1;
    if (ss.has_late) r[20] = ss.late_sb | (int32_t)(phys(1u << ss.pos[ss.late_bit]) << 19) | (flipmask ? (1 << 30) : 0);
    for (int lane = 0; lane < 32; ++lane) {
      uint32_t x = 0;
      for (size_t q = 0; q < ss.lanepos.size(); ++q) x |= (uint32_t)((lane >> q) & 1) << ss.lanepos[q];
      p.lanetab[(size_t)i * 32 + lane] = (phys(x) << 3) | (((x & submask) >> p.W) << 16);
    }
    int32_t *t = p.tb.data() + (size_t)i * TB_INTS;
    t[0] = M; t[1] = (int32_t)ss.layers.size(); t[2] = (int32_t)ss.looppos.size(); t[3] = ss.bpp; t[4] = ss.wbase;
    t[5] = ss.bpp ? 32 / ss.bpp : 0;
    t[6] = (int32_t)closed.size();
    t[7] = ss.has_late ? (ss.late_sb | (ss.late_bit << 16)) : -1;
    for (int b = 0; b < 4; ++b) t[8 + b] = b < M ? ss.pos[b] : -1;
    for (int q = 0; q < 5; ++q) {
      t[12 + q] = ss.lanepos[q];
      t[17 + q] = q < (int)ss.looppos.size() ? ss.looppos[q] : -1;
    }
    for (size_t q = 0; q < closed.size(); ++q) { t[22 + 2 * q] = closed[q].first; t[23 + 2 * q] = closed[q].second; }
    int bpoff = 0;
    for (size_t li = 0; li < ss.layers.size(); ++li) {
      const Layer &l = ss.layers[li];
      const int o = 30 + 14 * (int)li;
      t[o] = (int32_t)l.pinned.size(); t[o + 1] = (int32_t)l.free.size(); t[o + 2] = bpoff;
      for (int q = 0; q < 2; ++q) {
        if (q < (int)l.pinned.size()) { t[o + 3 + 2 * q] = l.pinned[q].second; t[o + 4 + 2 * q] = l.vars[l.pinned[q].first]; }
        else { t[o + 3 + 2 * q] = -1; t[o + 4 + 2 * q] = -1; }
        if (q < (int)l.free.size()) { t[o + 7 + 2 * q] = l.free[q].second; t[o + 8 + 2 * q] = l.vars[l.free[q].first]; }
        else { t[o + 7 + 2 * q] = 0; t[o + 8 + 2 * q] = -1; }
      }
      t[o + 11] = (li == 1 && ss.has_late) ? flipmask : 0;
      for (size_t q = 0; q < l.pinned.size(); ++q) t[o + 12 + q] = q < l.pk.size() ? l.pk[q] : 0;
      bpoff += (1 << M) * (int)l.free.size();
    }
  }
  if (p.tvals.empty()) p.tvals.push_back(0.0);
}

bool lower_sweep(const Schedule &sch, int max_head_bits, SweepPlan &plan) {
  for (auto &st : sch.steps)
    if (st.quad) throw std::runtime_error("lower_sweep expects the unfused schedule");
  if (sch.steps.size() != sch.factors.size()) throw std::runtime_error("lower_sweep expects the unfused schedule");
  const std::vector<Factor> &factors = sch.factors;
  const std::vector<Check> &checks = sch.checks;
  const bool maxplus = sch.semiring == TQEC_SEMIRING_MAXPLUS;
  // roles
  std::map<int, int> owner;
  for (size_t i = 0; i < factors.size(); ++i)
    for (int v : factors[i].vars) owner[v] = (int)i;
  std::vector<std::vector<int>> c_factors, f_checks(factors.size());
  for (size_t ci = 0; ci < checks.size(); ++ci) {
    std::set<int> fs;
    for (int v : checks[ci].vars) fs.insert(owner[v]);
    if (fs.empty()) return false;                                // orphan checks: general kernels only
    c_factors.push_back(std::vector<int>(fs.begin(), fs.end()));
    for (int fi : fs) f_checks[fi].push_back((int)ci);
  }
  std::vector<int> remaining;
  for (auto &x : c_factors) remaining.push_back((int)x.size());
  std::vector<Role> roles;
  {
    std::set<int> seen;
    for (int fi : sch.order) {
      Role R;
      R.fi = fi;
      R.touched = f_checks[fi];
      for (int c : R.touched)
        if (!seen.count(c)) R.opened.push_back(c);
      seen.insert(R.opened.begin(), R.opened.end());
      for (int c : R.touched) {
        remaining[c] -= 1;
        if (remaining[c] == 0 && checks[c].kind == 0) R.closing.push_back(c);
      }
      roles.push_back(R);
    }
  }
  const int n = (int)roles.size();
  std::vector<Classified> cls;
  for (auto &R : roles) cls.push_back(classify(factors[R.fi], R, checks));
  int h = n;
  while (h > 0 && cls[h - 1].ok) --h;
  if (h == 0) h = 1;
  std::vector<int> head_bits;
  for (int t = 0; t < h; ++t)
    for (int c : roles[t].closing) head_bits.push_back(checks[c].index);
  if ((int)head_bits.size() > max_head_bits || h >= n) return false;
  while (h + 1 < n && (int)(head_bits.size() + roles[h].closing.size()) <= max_head_bits) {
    for (int c : roles[h].closing) head_bits.push_back(checks[c].index);
    ++h;
  }
  std::vector<int> live;
  for (int t = 0; t < h; ++t) {
    live.insert(live.end(), roles[t].opened.begin(), roles[t].opened.end());
    std::vector<int> nl;
    for (int c : live)
      if (!has(roles[t].closing, c)) nl.push_back(c);
    live.swap(nl);
  }
  int W = (int)live.size();
  if (W > NB || W < 1) return false;
  const int W0 = W;                                            // chains alive after the head; fresh pins may add more
  std::map<int, int> chain_of;
  for (int k = 0; k < W; ++k) chain_of[live[k]] = k;
  const std::vector<int> live_order = live;

  std::vector<Raw> raw;
  std::map<int, int> dead_since;                               // chain -> step after which it is free again
  for (int t = h; t < n; ++t) {
    const Role &R = roles[t];
    const Classified &C = cls[t];
    Raw g;
    g.step = t; g.fi = R.fi;
    for (auto &p : C.pinned)
      if (p.c < 0) {
        // a chain that died in an earlier step (such two steps never share a super-step: see match), else a new one
        int ch = -1;
        for (auto &kv : dead_since)
          if (kv.second + 1 <= t) { ch = kv.first; break; }       // (std::map iterates in ascending chain order)
        if (ch >= 0) dead_since.erase(ch);
        else {
          ch = W++;
          if (W > NB) return false;
        }
        chain_of[p.o] = ch;
        g.fresh.push_back(ch);
      }
    for (auto &p : C.pinned) g.pinned.push_back({p.j, p.c >= 0 ? chain_of.at(p.c) : chain_of.at(p.o)});
    for (auto &p : C.pinned) {
      std::vector<int> k;
      for (int c : p.extra) k.push_back(chain_of.at(c));
      if (p.c < 0) k.push_back(chain_of.at(p.o));
      g.pk.push_back(k);
    }
    for (auto &fr : C.free) {
      std::vector<int> k;
      for (int c : fr.tv) k.push_back(chain_of.at(c));
      g.free.push_back({fr.j, k});
    }
    for (int c : R.closing) g.closed.push_back({checks[c].index, chain_of.at(c)});
    std::set<int> chs;
    for (auto &p : g.pinned) chs.insert(p.second);
    for (auto &k : g.pk) chs.insert(k.begin(), k.end());
    for (auto &fr : g.free) chs.insert(fr.second.begin(), fr.second.end());
    for (auto &c : g.closed) chs.insert(c.second);
    g.chains.assign(chs.begin(), chs.end());
    std::set<int> donors_used;
    for (auto &p : C.pinned)
      if (p.c >= 0) donors_used.insert(chain_of.at(p.c));
    for (auto &p : C.pinned)
      if (p.c >= 0) chain_of[p.o] = chain_of.at(p.c);
    for (int c : R.closing)
      if (!donors_used.count(chain_of.at(c))) {
        dead_since[chain_of.at(c)] = t;
        g.freed.push_back(chain_of.at(c));
      }
    const size_t nch = g.chains.size();
    raw.push_back(g);
    if (nch > (size_t)MAX_PATCH || nch == 0) return false;
  }
  const int sg = NB - W;
  if (sg > 5 || sg < 0) return false;
  std::vector<int> final_live;
  for (auto &kv : chain_of)
    if (checks[kv.first].kind == 1) final_live.push_back(kv.first);

  static const std::vector<Desc> MENU = build_menu();
  const int menu_limit = maxplus ? TQEC_SWEEP_MENU_MAXPLUS : (int)MENU.size();

  struct Match { bool ok = false; int menu = -1; std::vector<int> perm; };
  auto match = [&](int k, int cnt) {
    Match m;
    std::vector<const Raw *> group;
    for (int i = 0; i < cnt; ++i) group.push_back(&raw[k + i]);
    std::set<int> cs;
    for (auto *g : group) cs.insert(g->chains.begin(), g->chains.end());
    std::vector<int> chains(cs.begin(), cs.end());
    if ((int)chains.size() > MAX_PATCH) return m;
    if (cnt == 2)
      for (int a : group[0]->freed)
        for (int b : group[1]->fresh)
          if (a == b) return m;                                // a slot cannot die and reopen inside one super-step
    if (cnt == 2) {
      int both = 0;
      for (auto &p : group[0]->pinned)
        for (auto &c : group[1]->closed)
          if (p.second == c.second) ++both;
      // (chains of distinct pinned variables are distinct and a chain closes once per step, so this counts the set)
      if (both > 1) return m;
    }
    int nfree = 0;
    for (auto *g : group) nfree += (int)g->free.size();
    if (maxplus && (nfree << chains.size()) > 32) return m;
    // canonical form: the smallest descriptor over all orderings of the patch chains (first minimal ordering wins)
    bool have = false;
    Desc bestd;
    std::vector<int> bestp, perm = chains;
    do {
      Desc d = descriptor(group, perm);
      if (!have || d < bestd) { bestd = d; bestp = perm; have = true; }
    } while (std::next_permutation(perm.begin(), perm.end()));
    for (int i = 0; i < menu_limit; ++i)
      if (MENU[i] == bestd) { m.ok = true; m.menu = i; m.perm = bestp; return m; }
    return m;
  };

  const int nr = (int)raw.size();
  std::vector<Match> single(nr), pairm(nr);
  for (int k = 0; k < nr; ++k) {
    single[k] = match(k, 1);
    if (k + 1 < nr) pairm[k] = match(k, 2);
  }
  const int INF = 1000000000;
  std::vector<int> best(nr + 2, 0), take(nr, 0);
  best[nr + 1] = INF;
  for (int k = nr - 1; k >= 0; --k) {
    best[k] = INF;
    if (single[k].ok && 1 + best[k + 1] < best[k]) { best[k] = 1 + best[k + 1]; take[k] = 1; }
    if (pairm[k].ok && 1 + best[k + 2] <= best[k]) { best[k] = 1 + best[k + 2]; take[k] = 2; }
  }
  if (nr == 0 || best[0] >= INF) return false;

  std::vector<SuperStep> ssteps;
  for (int k = 0; k < nr;) {
    const int cnt = take[k];
    const Match &mt = cnt == 2 ? pairm[k] : single[k];
    std::map<int, int> bit;
    for (size_t b = 0; b < mt.perm.size(); ++b) bit[mt.perm[b]] = (int)b;
    SuperStep ss;
    ss.chains = mt.perm;
    ss.menu = mt.menu;
    for (int i = 0; i < cnt; ++i) {
      const Raw &g = raw[k + i];
      const Step &st = sch.steps[g.step];
      const Factor &f = factors[g.fi];
      Layer L;
      L.step = g.step; L.factor = g.fi; L.vars = f.vars;
      for (auto &p : g.pinned) L.pinned.push_back({p.first, bit[p.second]});
      for (auto &fr : g.free) {
        int m = 0;
        for (int ch : fr.second) m += 1 << bit[ch];
        L.free.push_back({fr.first, m});
      }
      for (auto &c : g.closed) L.closed.push_back({c.first, bit[c.second]});
      for (auto &chs : g.pk) {
        int m = 0;
        for (int ch : chs) m += 1 << bit[ch];
        L.pk.push_back(m);
      }
      const int NP = (int)L.pinned.size(), NF = (int)L.free.size();
      L.T.assign((size_t)1 << (NP + NF), 0.0);
      for (int pidx = 0; pidx < (1 << NP); ++pidx)
        for (int kk = 0; kk < (1 << NF); ++kk) {
          int a = 0;
          for (int q = 0; q < NP; ++q) a |= ((pidx >> q) & 1) << L.pinned[q].first;
          for (int q = 0; q < NF; ++q) a |= ((kk >> q) & 1) << L.free[q].first;
          L.T[((size_t)pidx << NF) | kk] = st.table[a];
        }
      ss.layers.push_back(L);
      ss.fresh.push_back(g.fresh);
    }
    if (cnt == 2) {
      int b = -1;
      for (auto &p : ss.layers[0].pinned)
        for (auto &c : ss.layers[1].closed)
          if (p.second == c.second) b = p.second;
      if (b >= 0) {
        int sb = 0;
        for (auto &c : ss.layers[1].closed)
          if (c.second == b) { sb = c.first; break; }
        ss.has_late = true; ss.late_sb = sb; ss.late_bit = b;
        std::vector<std::pair<int, int>> keep;
        for (auto &c : ss.layers[1].closed)
          if (c.second != b) keep.push_back(c);
        ss.layers[1].closed = keep;
      }
    }
    const int M = (int)mt.perm.size();
    ss.bpp = 0;
    if (maxplus)
      for (auto &l : ss.layers) ss.bpp += (1 << M) * (int)l.free.size();
    ssteps.push_back(ss);
    k += cnt;
  }

  std::vector<std::vector<int>> groups;
  for (auto &ss : ssteps) groups.push_back(ss.chains);
  auto ap = assign_positions(groups, W);
  const std::vector<int> &chain_pos = ap.first;
  std::set<int> alive;                                         // chains beyond W0 come alive when a fresh pin opens them
  for (int k = 0; k < W0; ++k) alive.insert(k);
  int wbase = 0;
  for (auto &ss : ssteps) {
    ss.pos.clear();
    for (int ch : ss.chains) ss.pos.push_back(chain_pos[ch]);
    std::vector<int> nonpatch;
    for (int p = 0; p < NB; ++p)
      if (!has(ss.pos, p)) nonpatch.push_back(p);
    std::set<int> active;
    for (int ch : alive) active.insert(chain_pos[ch]);
    for (int p = W; p < NB; ++p) active.insert(p);
    std::vector<int> lanes;
    for (int r = 0; r < 4; ++r) {
      std::vector<int> cands;
      for (int p : nonpatch)
        if (p < 8 && p % 4 == r && !has(lanes, p)) cands.push_back(p);
      std::stable_sort(cands.begin(), cands.end(), [&](int a, int b) {
        return std::make_pair(!active.count(a), a) < std::make_pair(!active.count(b), b);
      });
      if (!cands.empty()) lanes.push_back(cands[0]);
    }
    ss.conflict = lanes.size() < 4;
    std::vector<int> rest;
    for (int p : nonpatch)
      if (!has(lanes, p)) rest.push_back(p);
    std::stable_sort(rest.begin(), rest.end(), [&](int a, int b) {
      return std::make_tuple(!active.count(a), a >= W, a) < std::make_tuple(!active.count(b), b >= W, b);
    });
    while (lanes.size() < 5) { lanes.push_back(rest.front()); rest.erase(rest.begin()); }
    ss.lanepos = lanes;
    ss.looppos.clear();
    for (int p : rest)
      if (active.count(p)) ss.looppos.push_back(p);
    const int n_iter = 1 << ss.looppos.size();
    ss.n_words = 0;
    if (ss.bpp) {
      const int ipw = 32 / ss.bpp;
      ss.n_words = (n_iter + ipw - 1) / ipw;
    }
    ss.wbase = wbase;
    wbase += ss.n_words;
    for (size_t li = 0; li < ss.layers.size(); ++li) {
      const Layer &l = ss.layers[li];
      alive.insert(ss.fresh[li].begin(), ss.fresh[li].end());
      std::set<int> reused;
      for (auto &p : l.pinned) reused.insert(p.second);
      for (auto &c : l.closed)
        if (!reused.count(c.second)) alive.erase(ss.chains[c.second]);
    }
    if (ss.has_late) {
      bool re = false;
      for (auto &p : ss.layers[1].pinned) re = re || p.second == ss.late_bit;
      if (!re) alive.erase(ss.chains[ss.late_bit]);
    }
  }
  std::vector<int> out_index = {0};
  if (!maxplus) {
    std::vector<int> obs_pos(sch.n_obs, 0), a, b;
    for (int c : final_live) {
      obs_pos[checks[c].index] = chain_pos[chain_of.at(c)];
      a.push_back(chain_pos[chain_of.at(c)]);
    }
    for (int ch : alive) b.push_back(chain_pos[ch]);
    std::sort(a.begin(), a.end());
    std::sort(b.begin(), b.end());
    if (a != b) return false;
    out_index.assign((size_t)1 << sch.n_obs, 0);
    for (int i = 0; i < (1 << sch.n_obs); ++i) {
      int x = 0;
      for (int o = 0; o < sch.n_obs; ++o) x += ((i >> o) & 1) << obs_pos[o];
      out_index[i] = x;
    }
  } else if (!alive.empty()) {
    return false;
  }

  plan = SweepPlan();
  plan.semiring = sch.semiring; plan.n_vars = sch.n_vars; plan.n_checks = sch.n_checks; plan.n_obs = sch.n_obs;
  plan.W = W; plan.sg = sg; plan.head_steps = h; plan.n_ss = (int)ssteps.size(); plan.bp_words = wbase; plan.conflicts = ap.second;
  plan.head_bits = head_bits;
  plan.out_index = out_index;
  const auto tq_t0 = std::chrono::steady_clock::now();
  head_eval(sch, h, roles, head_bits, live_order, chain_pos, W, plan.head_state, plan.head_cfg);
  if (std::getenv("TQEC_LOWER_TIMING"))
    fprintf(stderr, "lower_sweep: head_eval %.3f s (%d head bits, %d steps)\n",
            std::chrono::duration<double>(std::chrono::steady_clock::now() - tq_t0).count(), (int)head_bits.size(), h);
  encode_sweep(plan, ssteps);
  return true;
}

}  // namespace lower
}  // namespace tqec
