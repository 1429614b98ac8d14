// Host-side lowering, part 3: passes of the global-memory executor k_wide_pass (see tqec_lower.h).
// Mirrors tensorqec.jl_b200/wide.py decision for decision (that file carries the full description).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <set>
#include <stdexcept>

#include "tqec_lower.h"

namespace tqec {
namespace lower {

static const int MAX_PASS_STEPS = 240;
static const int MAX_WIDE_WIDTH = 31;

static bool hasv(const std::vector<int> &v, int x) { return std::find(v.begin(), v.end(), x) != v.end(); }

struct WRole { int fi; std::vector<int> touched, opened, closing; };

struct LocalStep {
  int w_in = 0, n_open = 0, w_out = 0, nk = 0;
  std::vector<std::pair<int, int>> closed;
  uint32_t keepmask = 0;
  std::vector<int64_t> ML, MK;
  std::vector<double> T;
};

static LocalStep local_step(std::vector<int> &live, const Factor &f, const WRole &R, const std::vector<Check> &checks,
                            const std::vector<double> &table) {
  const int r = (int)f.vars.size(), w_in = (int)live.size();
  std::vector<int> full = live;
  full.insert(full.end(), R.opened.begin(), R.opened.end());
  std::map<int, int> pos;
  for (size_t k = 0; k < full.size(); ++k) pos[full[k]] = (int)k;
  std::vector<int64_t> m;
  for (int v : f.vars) {
    int64_t mv = 0;
    for (int c : R.touched)
      if (std::find(checks[c].vars.begin(), checks[c].vars.end(), v) != checks[c].vars.end()) mv |= (int64_t)1 << pos[c];
    m.push_back(mv);
  }
  const int NA = 1 << r, n_open = (int)R.opened.size();
  std::vector<int64_t> M(NA, 0), pat(NA), a0((size_t)1 << n_open, -1), ker;
  for (int a = 0; a < NA; ++a) {
    for (int j = 0; j < r; ++j)
      if ((a >> j) & 1) M[a] ^= m[j];
    pat[a] = M[a] >> w_in;
  }
  for (int a = NA - 1; a >= 0; --a) a0[(size_t)pat[a]] = a;
  for (int a = 0; a < NA; ++a)
    if (pat[a] == 0) ker.push_back(a);
  LocalStep ls;
  ls.w_in = w_in; ls.n_open = n_open; ls.nk = (int)ker.size();
  const int64_t inmask = ((int64_t)1 << w_in) - 1;
  std::vector<int> keep;
  for (int c : full)
    if (!hasv(R.closing, c)) keep.push_back(c);
  for (int c : keep) ls.keepmask |= 1u << pos[c];
  for (int c : R.closing) ls.closed.push_back({pos[c], checks[c].index});
  std::sort(ls.closed.begin(), ls.closed.end());
  for (int p = 0; p < (1 << n_open); ++p) ls.ML.push_back(a0[p] >= 0 ? (M[(size_t)a0[p]] & inmask) : 0);
  for (int64_t k : ker) ls.MK.push_back(M[(size_t)k] & inmask);
  ls.T.assign(((size_t)1 << n_open) * ls.nk, 0.0);
  for (int p = 0; p < (1 << n_open); ++p)
    if (a0[p] >= 0)
      for (int k = 0; k < ls.nk; ++k) ls.T[(size_t)p * ls.nk + k] = table[(size_t)(a0[p] ^ ker[k])];
  ls.w_out = (int)keep.size();
  live = keep;
  return ls;
}

// ---- butterfly passes ------------------------------------------------------------------------------------------------------
// A pass whose steps are all rank-1 factors [t0, t1] (one mechanism of a detector error model) gets a second encoding for
// k_wide_bf (tqec_wide.cu).  Every tile check has a fixed POSITION (bit of the shared-memory index) for the whole pass:
// input tile bit i sits at position i, an opened check takes a dead position, a closed check frees its position --
// nothing is compacted, the kernel folds the closed checks' syndrome bits into its addresses, and dead entries are kept
// at zero (zero-filled at the tile load; a position that is reused is zeroed when it reopens).  With the factors
// normalised to [1, r = t1 / t0] (the product of the t0 goes into the plan's output scale) every step, opening or not,
// is the butterfly  S[x] += r S[x ^ m],  S[x ^ m] += r S[x]  over the pairs of its position mask m -- commuting linear
// maps, so the steps of a pass may be applied in any order.  Consecutive steps whose masks span at most TQEC_BF_G
// dimensions form a GROUP: a thread holds one coset of the span (2^G entries) in registers.  Coordinates refer to the
// basis made of the masks of the group's first independent steps (plus unit vectors for padding), so those steps pair
// the coset members that differ in ONE coordinate -- straight-line code, ratio 0 for a padding dimension -- and only the
// remaining (dependent) steps go through a run-time switch on their coordinate vector.  A reduced echelon copy of the
// basis (pivot position p_j in vector j only) serves to test dependence and to enumerate coset representatives.
struct BfGroup {
  std::vector<uint32_t> red, expr, bvec;      // reduced vectors, their coordinates, the coordinate basis
  std::vector<int> pivot;
  std::vector<double> runit;                  // ratio of the step that introduced basis vector j (0: none)
  std::vector<std::pair<int, double>> dep;    // dependent steps: (coordinates, ratio)
  std::vector<std::pair<int, int>> closes;    // (position, syndrome bit)
  std::vector<int> zero_pos;                  // positions that reopen in this group with stale entries in their dead half
  uint32_t live_start = 0, touched = 0, freed = 0;
};

static int parity32(uint32_t x) { return __builtin_popcount(x) & 1; }

// reduce `v` against the group's echelon basis -> (remainder, coordinates of the part that was removed)
static std::pair<uint32_t, uint32_t> bf_reduce(const BfGroup &g, uint32_t v) {
  uint32_t cu = 0;
  for (size_t j = 0; j < g.red.size(); ++j)
    if ((v >> g.pivot[j]) & 1u) { v ^= g.red[j]; cu ^= g.expr[j]; }
  return {v, cu};
}

// new coordinate basis vector `bv` whose remainder after reduction is `v` (non-zero), pivot `piv`
static void bf_add_basis(BfGroup &g, uint32_t bv, uint32_t v, uint32_t cu, int piv, double r) {
  const uint32_t ev = (1u << g.bvec.size()) ^ cu;
  for (size_t j = 0; j < g.red.size(); ++j)
    if ((g.red[j] >> piv) & 1u) { g.red[j] ^= v; g.expr[j] ^= ev; }
  g.red.push_back(v); g.expr.push_back(ev); g.pivot.push_back(piv); g.bvec.push_back(bv); g.runit.push_back(r);
}

// bank image of a position under the kernel's swizzle phys(x) = x ^ ((x >> 4) & 14) ^ ((x >> 7) & 14): bank pair = phys & 15
static int bf_bank_image(int q) { return q < 4 ? (1 << q) : (q >= 5 && q <= 7 ? (1 << (q - 4)) : (TQEC_BF_SWZ_WIDE && q >= 8 && q <= 10 ? (1 << (q - 7)) : 0)); }

#define BF_FAIL(N) do { if (std::getenv("TQEC_BF_DEBUG")) std::fprintf(stderr, "bf_encode_pass: steps %d..%d rejected (reason %d)\n", t0, t1, N); return false; } while (0)
static bool bf_encode_pass(const std::vector<WRole> &roles, int t0, int t1, const std::vector<Factor> &factors,
                           const std::vector<Check> &checks, const std::vector<std::vector<double>> &tabs,
                           const std::vector<int> &L_in, const std::vector<int> &L_out, int n_pos, int gmax, int n_spec,
                           std::vector<int32_t> &ints, std::vector<double> &vals, double &mant, int &exp2) {
  if (n_pos > 12 || (int)L_in.size() > n_pos || (int)L_out.size() > n_pos || t1 - t0 > 240) BF_FAIL(1);
  if (n_spec > 19) BF_FAIL(13);                                  // k_wide_bf tabulates the spectator deposit for 7 + 7 + 5 bits
  std::map<int, int> pos;
  uint32_t live = 0, ever = 0;
  for (size_t i = 0; i < L_in.size(); ++i) { pos[L_in[i]] = (int)i; live |= 1u << i; }
  ever = live;
  std::vector<BfGroup> groups;
  BfGroup cur;
  cur.live_start = live;
  auto flush = [&]() {
    groups.push_back(cur);
    cur = BfGroup();
    cur.live_start = live;
  };
  double m_acc = 1.0;
  int e_acc = 0;
  for (int t = t0; t < t1; ++t) {
    const WRole &R = roles[t];
    const Factor &f = factors[R.fi];
    if (f.vars.size() != 1) BF_FAIL(2);
    const double a0 = tabs[t][0], a1 = tabs[t][1];
    const double r = a1 / a0;
    if (!(a0 != 0.0) || !std::isfinite(r) || !std::isfinite(a0)) BF_FAIL(3);
    // positions for the opened checks: dead, and not touched or freed inside the current group
    std::vector<int> pnews;
    const int n_opened = (int)R.opened.size();
    if (n_opened > gmax) BF_FAIL(4);
    if (n_opened > 0) {
      for (int attempt = 0; attempt < 2 && (int)pnews.size() < n_opened; ++attempt) {
        pnews.clear();
        const bool room = (int)cur.bvec.size() + n_opened <= gmax;   // every opened check takes a basis vector of its own
        for (int p = 0; room && p < n_pos && (int)pnews.size() < n_opened; ++p)
          if (!((live >> p) & 1u) && !(((cur.touched | cur.freed) >> p) & 1u)) pnews.push_back(p);
        if ((int)pnews.size() < n_opened && attempt == 0) flush();
      }
      if ((int)pnews.size() < n_opened) BF_FAIL(5);
      for (int i = 0; i < n_opened; ++i) pos[R.opened[i]] = pnews[i];
    }
    uint32_t m = 0;
    for (int c : R.touched) m |= 1u << pos.at(c);
    if (m == 0) BF_FAIL(6);
    // checks opened beyond the first: a unit basis vector each, so that the coset holds the entries where only some of
    // the new checks are set (they are zero and stay zero)
    for (int i = 1; i < n_opened; ++i) {
      auto rc = bf_reduce(cur, 1u << pnews[i]);
      bf_add_basis(cur, 1u << pnews[i], rc.first, rc.second, pnews[i], 0.0);
    }
    auto rc = bf_reduce(cur, m);
    if (rc.first != 0) {
      if ((int)cur.bvec.size() == gmax) {
        if (n_opened > 0) BF_FAIL(7);                            // cannot happen: room was checked above
        flush();
        rc = {m, 0u};
      }
      const int piv = n_opened > 0 ? pnews[0] : 31 - __builtin_clz(rc.first);
      bf_add_basis(cur, m, rc.first, rc.second, piv, r);
    } else {
      if (n_opened > 0 || rc.second == 0) BF_FAIL(8);            // cannot happen: an opened position is in no basis vector
      cur.dep.push_back({(int)rc.second, r});
    }
    cur.touched |= m;
    for (int p : pnews) {
      if ((ever >> p) & 1u) cur.zero_pos.push_back(p);
      live |= 1u << p;
      ever |= 1u << p;
    }
    for (int cc : R.closing) {
      const int p = pos.at(cc);
      cur.closes.push_back({p, checks[cc].index});
      cur.freed |= 1u << p;
      live &= ~(1u << p);
    }
    int e = 0;
    m_acc = std::frexp(m_acc * a0, &e);
    e_acc += e;
  }
  groups.push_back(cur);
  if (groups.size() > 32) BF_FAIL(9);
  // pad every basis to gmax vectors with unit vectors (ratio 0) and choose the order in which cosets are enumerated
  std::vector<uint32_t> orders, zmasks;
  std::vector<int> n_frees;
  for (auto &g : groups) {
    while ((int)g.bvec.size() < gmax) {
      uint32_t piv = 0;
      for (int j : g.pivot) piv |= 1u << j;
      uint32_t free_ = g.live_start & ~piv;
      // no live position left: a dead position nothing in the group touches (its half of the coset is zero and stays zero)
      if (!free_) free_ = ((1u << n_pos) - 1u) & ~g.live_start & ~g.touched & ~piv;
      if (!free_) BF_FAIL(10);
      const int q = 31 - __builtin_clz(free_);
      auto rq = bf_reduce(g, 1u << q);
      bf_add_basis(g, 1u << q, rq.first, rq.second, q, 0.0);
    }
    uint32_t piv = 0;
    for (int j : g.pivot) piv |= 1u << j;
    const uint32_t free_ = g.live_start & ~piv;
    std::vector<int> fr, ord;
    for (int q = 0; q < n_pos; ++q)
      if ((free_ >> q) & 1u) fr.push_back(q);
    if (fr.size() > 8) BF_FAIL(11);
    // the 16 lanes of a half-warp follow the first four coset-index bits: pick positions with independent bank images
    uint32_t span[4] = {0, 0, 0, 0};
    std::vector<char> used(fr.size(), 0);
    for (size_t i = 0; i < fr.size() && ord.size() < 4; ++i) {
      int x = bf_bank_image(fr[i]);
      for (int b = 3; b >= 0 && x; --b)
        if ((x >> b) & 1) { if (span[b]) x ^= (int)span[b]; else { span[b] = (uint32_t)x; ord.push_back(fr[i]); used[i] = 1; x = 0; } }
    }
    for (size_t i = 0; i < fr.size(); ++i)
      if (!used[i]) ord.push_back(fr[i]);
    uint32_t packed = 0;
    for (size_t i = 0; i < ord.size(); ++i) packed |= (uint32_t)ord[i] << (4 * i);
    orders.push_back(packed);
    n_frees.push_back((int)ord.size());
    // coset members to zero at the load: their bit at a reopening position differs from the dead value
    uint32_t Z = 0;
    for (int p : g.zero_pos) {
      uint32_t pm = 0;
      for (int j = 0; j < gmax; ++j)
        if ((g.bvec[j] >> p) & 1u) pm |= 1u << j;
      for (int k = 0; k < (1 << gmax); ++k)
        if (parity32((uint32_t)k & pm)) Z |= 1u << k;
    }
    zmasks.push_back(Z);
  }
  int n_dep = 0, n_closes = 0;
  for (auto &g : groups) { n_dep += (int)g.dep.size(); n_closes += (int)g.closes.size(); }
  if (n_dep > 96 || n_closes > 64) BF_FAIL(12);
  ints.push_back((int32_t)groups.size()); ints.push_back(n_dep); ints.push_back(n_closes); ints.push_back((int32_t)vals.size());
  ints.push_back(n_pos); ints.push_back((int32_t)L_in.size()); ints.push_back((int32_t)L_out.size()); ints.push_back(gmax);
  for (int i = 0; i < 12; ++i) ints.push_back(i < (int)L_out.size() ? pos.at(L_out[i]) : -1);
  int d0 = 0, c0 = 0;
  for (size_t gi = 0; gi < groups.size(); ++gi) {
    const BfGroup &g = groups[gi];
    int32_t rec[TQEC_BF_GROUP_INTS] = {0};
    rec[0] = n_frees[gi]; rec[1] = d0; rec[2] = (int32_t)g.dep.size(); rec[3] = c0; rec[4] = (int32_t)g.closes.size();
    for (int j = 0; j < gmax; ++j) rec[5 + j] = (int32_t)g.bvec[j];
    rec[10] = (int32_t)zmasks[gi];
    rec[11] = (int32_t)vals.size();                              // gmax unit ratios, then the dependent steps' ratios
    rec[13] = (int32_t)orders[gi];
    ints.insert(ints.end(), rec, rec + TQEC_BF_GROUP_INTS);
    for (int j = 0; j < gmax; ++j) vals.push_back(g.runit[j]);
    for (auto &dp : g.dep) vals.push_back(dp.second);
    d0 += (int)g.dep.size(); c0 += (int)g.closes.size();
  }
  for (auto &g : groups)
    for (auto &dp : g.dep) ints.push_back((int32_t)dp.first);
  for (auto &g : groups)
    for (auto &cl : g.closes) { ints.push_back(cl.first); ints.push_back(cl.second); }
  int e = 0;
  mant = std::frexp(mant * m_acc, &e);
  exp2 += e + e_acc;
  return true;
}
#undef BF_FAIL

WidePlan lower_wide(const std::vector<Factor> &factors_in, const std::vector<Check> &checks_in, int semiring, int n_vars, int n_checks,
                    int n_obs, const std::vector<int> *order_in, int t_max, int low_bits, double max_drop_bits) {
  if (semiring != TQEC_SEMIRING_SUMPROD) throw std::runtime_error("the global-memory executor runs sum-product plans only");
  std::vector<Factor> factors = merge_overlapping(factors_in, n_vars, checks_in, true);
  std::vector<Check> checks;
  for (auto &c : checks_in) {
    Check d;
    d.kind = c.kind; d.index = c.index;
    for (int v : c.vars)
      if (!hasv(d.vars, v)) d.vars.push_back(v);
    checks.push_back(d);
  }
  std::vector<int> order = order_in ? *order_in : choose_order(factors, checks);
  {
    std::vector<int> so(order);
    std::sort(so.begin(), so.end());
    bool ok = so.size() == factors.size();
    for (size_t i = 0; ok && i < so.size(); ++i) ok = so[i] == (int)i;
    if (!ok) throw std::runtime_error("order must be a permutation of the (merged) factors");
  }
  // roles
  std::map<int, int> owner;
  for (size_t i = 0; i < factors.size(); ++i)
    for (int v : factors[i].vars) owner[v] = (int)i;
  std::vector<std::vector<int>> c_factors, f_checks(factors.size());
  for (size_t ci = 0; ci < checks.size(); ++ci) {
    std::set<int> fs;
    for (int v : checks[ci].vars) fs.insert(owner.at(v));
    if (fs.empty()) throw std::runtime_error("wide lowering: a check without variables (orphan) is not supported");
    c_factors.push_back(std::vector<int>(fs.begin(), fs.end()));
    for (int fi : fs) f_checks[fi].push_back((int)ci);
  }
  std::vector<int> remaining;
  for (auto &x : c_factors) remaining.push_back((int)x.size());
  std::vector<WRole> roles;
  {
    std::set<int> seen;
    for (int fi : order) {
      WRole R;
      R.fi = fi; R.touched = f_checks[fi];
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
  std::vector<std::vector<double>> tabs;
  int log2_scale = 0;
  double log2_run = 0.0;
  for (auto &R : roles) {
    std::vector<double> tab = factors[R.fi].table;
    double mx = 0.0;
    for (double x : tab) mx = std::max(mx, std::fabs(x));
    if (mx > 0.0) {
      log2_run += std::log2(mx);
      const int e = (int)std::nearbyint(log2_run);
      log2_run -= e;
      for (double &x : tab) x = std::ldexp(x, -e);
      log2_scale += e;
    }
    tabs.push_back(tab);
  }
  std::vector<double> drops;   // log2(largest / smallest non-zero entry) of every step's table
  for (auto &tab : tabs) {
    double mx = 0.0, mn = 0.0;
    for (double x0 : tab) {
      const double x = std::fabs(x0);
      if (x > 0.0) { mx = std::max(mx, x); mn = mn == 0.0 ? x : std::min(mn, x); }
    }
    drops.push_back(mx > 0.0 ? std::log2(mx / mn) : 0.0);
  }

  struct SimOut { bool ok = false; std::set<int> tile; std::vector<int> g; int peak = 0; };
  auto simulate = [&](int t0, int t1, const std::vector<int> &glive, int lb) {
    SimOut so;
    std::set<int> touched_all;
    std::vector<int> g = glive;
    for (int t = t0; t < t1; ++t) {
      const WRole &R = roles[t];
      touched_all.insert(R.touched.begin(), R.touched.end());
      std::vector<int> ng;
      for (int c : g)
        if (!hasv(R.closing, c)) ng.push_back(c);
      for (int c : R.opened)
        if (!hasv(R.closing, c)) ng.push_back(c);
      g.swap(ng);
    }
    so.tile = touched_all;
    for (int k = 0; k < lb && k < (int)glive.size(); ++k) so.tile.insert(glive[k]);
    for (int k = 0; k < lb && k < (int)g.size(); ++k) so.tile.insert(g[k]);
    int w = 0;
    for (int c : glive)
      if (so.tile.count(c)) ++w;
    int peak = w;
    for (int t = t0; t < t1; ++t) {
      const WRole &R = roles[t];
      peak = std::max(peak, w + (int)R.opened.size());           // opened checks coexist with the ones this step closes
      w = w + (int)R.opened.size() - (int)R.closing.size();
      if (w + (int)R.closing.size() > MAX_WIDE_WIDTH) return so;
    }
    if (peak > t_max) return so;
    if (max_drop_bits > 0 && t1 - t0 > 1) {
      double acc = 0.0;
      for (int t = t0; t < t1; ++t) acc += drops[t];
      if (acc > max_drop_bits) return so;
    }
    so.ok = true; so.g = g; so.peak = peak;
    return so;
  };

  // plans made of rank-1 factors only (detector error models) keep a pass within what one butterfly block holds
  bool all_rank1 = true;
  for (auto &f : factors) all_rank1 = all_rank1 && f.vars.size() == 1;
  const int max_pass_steps = all_rank1 ? 96 : MAX_PASS_STEPS;
  WidePlan P;
  P.semiring = semiring; P.n_vars = n_vars; P.n_checks = n_checks; P.n_obs = n_obs; P.t_max = t_max;
  std::vector<int> glive;
  int w_cap = 0, t = 0, step_count = 0;
  double cost = 0.0, traffic = 0.0;
  while (t < n) {
    bool have = false;
    int best_t1 = 0;
    SimOut best;
    for (int lb = low_bits; lb >= 0 && !have; --lb) {
      for (int t1 = t + 1; t1 <= std::min(n, t + max_pass_steps); ++t1) {
        SimOut r = simulate(t, t1, glive, lb);
        if (!r.ok) break;
        best = r; best_t1 = t1; have = true;
      }
    }
    if (!have) throw std::runtime_error("wide lowering: step " + std::to_string(t) + " alone needs more than " + std::to_string(t_max) + " tile bits");
    // spare tile bits go to the lowest untouched spectators: larger tiles, longer contiguous runs in HBM
    for (size_t k = 0, room = (size_t)(t_max - best.peak); k < glive.size() && room > 0; ++k)
      if (!best.tile.count(glive[k])) { best.tile.insert(glive[k]); --room; }
    const std::vector<int> &gout = best.g;
    if ((int)glive.size() > MAX_WIDE_WIDTH || (int)gout.size() > MAX_WIDE_WIDTH)
      throw std::runtime_error("frontier needs " + std::to_string(std::max(glive.size(), gout.size())) + " bits > 31");
    uint32_t tin_mask = 0, tout_mask = 0;
    std::vector<int> L;
    for (size_t k = 0; k < glive.size(); ++k)
      if (best.tile.count(glive[k])) { tin_mask |= 1u << k; L.push_back(glive[k]); }
    for (size_t k = 0; k < gout.size(); ++k)
      if (best.tile.count(gout[k])) tout_mask |= 1u << k;
    const int t_in = (int)L.size(), n_spec = (int)glive.size() - t_in;
    const std::vector<int> L_in = L;
    const size_t i0 = P.ints.size(), f0 = P.tables.size();
    const int s0 = step_count;
    for (int tt = t; tt < best_t1; ++tt) {
      LocalStep ls = local_step(L, factors[roles[tt].fi], roles[tt], checks, tabs[tt]);
      int32_t q[TQEC_WIDE_STEP_INTS] = {0};
      q[TQEC_WL_WIN] = ls.w_in; q[TQEC_WL_NOPEN] = ls.n_open; q[TQEC_WL_NCLOSE] = (int)ls.closed.size(); q[TQEC_WL_WOUT] = ls.w_out;
      q[TQEC_WL_NK] = ls.nk;
      q[TQEC_WL_OFF_T] = (int32_t)(P.tables.size() - f0);
      P.tables.insert(P.tables.end(), ls.T.begin(), ls.T.end());
      q[TQEC_WL_OFF_ML] = (int32_t)(P.ints.size() - i0);
      for (int64_t x : ls.ML) P.ints.push_back((int32_t)x);
      q[TQEC_WL_OFF_MK] = (int32_t)(P.ints.size() - i0);
      for (int64_t x : ls.MK) P.ints.push_back((int32_t)x);
      q[TQEC_WL_OFF_CLOSE] = (int32_t)(P.ints.size() - i0);
      for (auto &c : ls.closed) { P.ints.push_back(c.first); P.ints.push_back(c.second); }
      q[TQEC_WL_KEEPMASK] = (int32_t)ls.keepmask;
      P.step_hdr.insert(P.step_hdr.end(), q, q + TQEC_WIDE_STEP_INTS);
      cost += std::ldexp(1.0, n_spec + ls.w_out) * ls.nk;
      P.w_peak = std::max(P.w_peak, n_spec + ls.w_out);
      ++step_count;
    }
    int32_t hdr[TQEC_WIDE_PASS_INTS] = {0};
    hdr[TQEC_WP_WIN] = (int)glive.size(); hdr[TQEC_WP_WOUT] = (int)gout.size(); hdr[TQEC_WP_TIN] = t_in; hdr[TQEC_WP_TOUT] = (int)L.size();
    hdr[TQEC_WP_NSTEPS] = best_t1 - t; hdr[TQEC_WP_STEP0] = s0; hdr[TQEC_WP_TINMASK] = (int32_t)tin_mask; hdr[TQEC_WP_TOUTMASK] = (int32_t)tout_mask;
    hdr[TQEC_WP_OFF_INTS] = (int32_t)i0; hdr[TQEC_WP_N_INTS] = (int32_t)(P.ints.size() - i0);
    hdr[TQEC_WP_OFF_TAB] = (int32_t)f0; hdr[TQEC_WP_N_TAB] = (int32_t)(P.tables.size() - f0);
    P.pass_hdr.insert(P.pass_hdr.end(), hdr, hdr + TQEC_WIDE_PASS_INTS);
    {
      std::vector<int> L_out;
      for (int c : gout)
        if (best.tile.count(c)) L_out.push_back(c);
      const size_t keep_i = P.bf_ints.size(), keep_v = P.bf_vals.size();
      double mant = P.bf_mant;
      int exp2 = P.bf_log2;
      const int n_pos = std::min(12, t_max);
      if (bf_encode_pass(roles, t, best_t1, factors, checks, tabs, L_in, L_out, n_pos, TQEC_BF_G, n_spec, P.bf_ints, P.bf_vals, mant, exp2)) {
        P.bf_off.push_back((int32_t)keep_i);
        P.bf_mant = mant; P.bf_log2 = exp2;
      } else {
        P.bf_ints.resize(keep_i); P.bf_vals.resize(keep_v);
        P.bf_off.push_back(-1);
      }
    }
    traffic += 8.0 * (std::ldexp(1.0, (int)glive.size()) + std::ldexp(1.0, (int)gout.size()));
    w_cap = std::max(w_cap, (int)std::max(glive.size(), gout.size()));
    glive = gout;
    t = best_t1;
  }
  P.obs_pos.assign(n_obs, -1);
  for (size_t k = 0; k < glive.size(); ++k) {
    if (checks[glive[k]].kind != 1) throw std::runtime_error("a clamped check survived the sweep");
    if (checks[glive[k]].index < 0 || checks[glive[k]].index >= n_obs) throw std::runtime_error("observable index out of range");
    P.obs_pos[checks[glive[k]].index] = (int)k;
  }
  for (int p : P.obs_pos)
    if (p < 0) throw std::runtime_error("every observable row must be declared exactly once");
  if ((int)glive.size() != n_obs) throw std::runtime_error("every observable row must be declared exactly once");
  P.w_cap = w_cap; P.n_pass = (int)(P.pass_hdr.size() / TQEC_WIDE_PASS_INTS); P.n_steps = step_count; P.log2_scale = log2_scale;
  P.order = order; P.cost = cost; P.bytes_per_shot = traffic;
  if (P.ints.empty()) P.ints.push_back(0);
  if (P.tables.empty()) P.tables.push_back(0.0);
  if (P.bf_ints.empty()) P.bf_ints.push_back(0);
  if (P.bf_vals.empty()) P.bf_vals.push_back(0.0);
  return P;
}

}  // namespace lower
}  // namespace tqec
