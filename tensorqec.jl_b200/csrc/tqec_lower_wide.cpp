// Host-side lowering, part 3: passes of the global-memory executor k_wide_pass (see tqec_lower.h).
// Mirrors tensorqec.jl_b200/wide.py decision for decision (that file carries the full description).
#include <algorithm>
#include <cmath>
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
      w = w + (int)R.opened.size() - (int)R.closing.size();
      peak = std::max(peak, w);
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
      for (int t1 = t + 1; t1 <= std::min(n, t + MAX_PASS_STEPS); ++t1) {
        SimOut r = simulate(t, t1, glive, lb);
        if (!r.ok) break;
        best = r; best_t1 = t1; have = true;
      }
    }
    if (!have) throw std::runtime_error("wide lowering: step " + std::to_string(t) + " alone needs more than " + std::to_string(t_max) + " tile bits");
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
  return P;
}

}  // namespace lower
}  // namespace tqec
