// Host-side lowering, part 1: factor merging, absorption order, frontier schedule (see tqec_lower.h).
// Mirrors tensorqec.jl_b200/schedule.py decision for decision, so that both produce the same tables.
#include "tqec_lower.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <stdexcept>

namespace tqec {
namespace lower {

static const int MAX_FACTOR_RANK = 10;
static const int MAX_SMEM_WIDTH = 13;
static const int MAX_WIDE_WIDTH = 31;

[[noreturn]] static void fail(const std::string &msg) { throw std::runtime_error(msg); }

static bool contains(const std::vector<int> &v, int x) { return std::find(v.begin(), v.end(), x) != v.end(); }

// ---- merge_overlapping (schedule.py) -----------------------------------------------------------------------------------
std::vector<Factor> merge_overlapping(const std::vector<Factor> &factors, int n_vars, const std::vector<Check> &checks,
                                      bool allow_negative) {
  (void)n_vars;
  const int nf = (int)factors.size();
  std::vector<int> parent(nf);
  for (int i = 0; i < nf; ++i) parent[i] = i;
  auto find = [&](int i) {
    while (parent[i] != i) {
      parent[i] = parent[parent[i]];
      i = parent[i];
    }
    return i;
  };
  std::map<int, int> owner;
  for (int i = 0; i < nf; ++i) {
    const Factor &f = factors[i];
    std::set<int> uniq(f.vars.begin(), f.vars.end());
    if (uniq.size() != f.vars.size()) fail("factor " + std::to_string(i) + " repeats a variable");
    for (int v : f.vars) {
      auto it = owner.find(v);
      if (it != owner.end()) parent[find(i)] = find(it->second);
      else owner[v] = i;
    }
  }
  std::map<int, std::vector<int>> groups;
  for (int i = 0; i < nf; ++i) groups[find(i)].push_back(i);
  std::vector<std::pair<int, int>> roots;  // (min member, root)
  for (auto &g : groups) roots.push_back({*std::min_element(g.second.begin(), g.second.end()), g.first});
  std::sort(roots.begin(), roots.end());
  std::vector<Factor> out;
  for (auto &rr : roots) {
    const std::vector<int> &members = groups[rr.second];
    if (members.size() == 1) {
      out.push_back(factors[members[0]]);
      continue;
    }
    std::vector<int> vs;
    for (int i : members)
      for (int v : factors[i].vars)
        if (!contains(vs, v)) vs.push_back(v);
    if ((int)vs.size() > MAX_FACTOR_RANK) fail("overlapping prior factors merge into rank " + std::to_string(vs.size()) + " > 10");
    Factor m;
    m.vars = vs;
    m.table.assign((size_t)1 << vs.size(), 1.0);
    for (int i : members) {
      const Factor &f = factors[i];
      std::vector<int> where;
      for (int v : f.vars) where.push_back((int)(std::find(vs.begin(), vs.end(), v) - vs.begin()));
      for (size_t a = 0; a < m.table.size(); ++a) {
        size_t idx = 0;
        for (size_t j = 0; j < where.size(); ++j) idx |= ((a >> where[j]) & 1) << j;
        m.table[a] = m.table[a] * f.table[idx];
      }
    }
    out.push_back(m);
  }
  std::set<int> covered;
  for (auto &f : out)
    for (int v : f.vars) covered.insert(v);
  std::set<int> cv;
  for (auto &c : checks)
    for (int v : c.vars) cv.insert(v);
  for (int v : cv)
    if (!covered.count(v)) {
      Factor u;
      u.vars = {v};
      u.table = {1.0, 1.0};
      out.push_back(u);
    }
  for (auto &f : out) {
    if ((int)f.vars.size() > MAX_FACTOR_RANK) fail("prior factor of rank " + std::to_string(f.vars.size()) + " > 10 is not supported");
    if (f.table.size() != ((size_t)1 << f.vars.size())) fail("factor table must have 2^rank entries");
    for (double x : f.table)
      if (!std::isfinite(x) || (x < 0.0 && !allow_negative)) fail("prior factor entries must be finite and non-negative");
  }
  return out;
}

std::vector<int> map_order(const std::vector<Factor> &original, const std::vector<Factor> &merged, const std::vector<int> &order) {
  std::vector<int> sorted_o(order);
  std::sort(sorted_o.begin(), sorted_o.end());
  for (size_t i = 0; i < sorted_o.size(); ++i)
    if (sorted_o[i] != (int)i) fail("order must be a permutation of the prior tensors");
  if (order.size() != original.size()) fail("order must be a permutation of the prior tensors");
  std::map<int, int> home;
  for (size_t mi = 0; mi < merged.size(); ++mi)
    for (int v : merged[mi].vars) home[v] = (int)mi;
  std::vector<int> out;
  for (int i : order)
    for (int v : original[i].vars) {
      const int mi = home[v];
      if (!contains(out, mi)) out.push_back(mi);
    }
  for (int mi = 0; mi < (int)merged.size(); ++mi)
    if (!contains(out, mi)) out.push_back(mi);
  return out;
}

// ---- ordering -----------------------------------------------------------------------------------------------------------
struct Sim {
  const std::vector<Factor> &factors;
  const std::vector<Check> &checks;
  std::vector<std::vector<int>> f_checks, c_factors;
  std::vector<std::vector<char>> c_has;      // c_has[c][v_local]: membership test by map
  std::map<int, int> var_owner;
  Sim(const std::vector<Factor> &f, const std::vector<Check> &c) : factors(f), checks(c) {
    f_checks.resize(f.size());
    for (size_t i = 0; i < f.size(); ++i)
      for (int v : f[i].vars) var_owner[v] = (int)i;
    for (size_t ci = 0; ci < c.size(); ++ci) {
      std::set<int> fs;
      for (int v : c[ci].vars) {
        auto it = var_owner.find(v);
        if (it == var_owner.end()) fail("check variable without a factor");
        fs.insert(it->second);
      }
      c_factors.push_back(std::vector<int>(fs.begin(), fs.end()));
      for (int fi : fs) f_checks[fi].push_back((int)ci);
    }
  }
  bool in_check(int c, int v) const { return contains(checks[c].vars, v); }
};

static int gf2_rank(std::vector<uint64_t> rows) {
  rows.erase(std::remove(rows.begin(), rows.end(), (uint64_t)0), rows.end());
  int rank = 0;
  while (!rows.empty()) {
    const uint64_t p = *std::max_element(rows.begin(), rows.end());
    int hb = 63;
    while (!((p >> hb) & 1)) --hb;
    std::vector<uint64_t> nxt;
    for (uint64_t r : rows) {
      if (r == p) continue;
      const uint64_t q = ((r >> hb) & 1) ? (r ^ p) : r;
      if (q) nxt.push_back(q);
    }
    rows.swap(nxt);
    ++rank;
  }
  return rank;
}

// (w_out, log2 work) of absorbing factor fi; live[c] = 1 if check c is open
static std::pair<int, int> step_cost(const Sim &sim, int fi, const std::vector<int> &remaining, const std::vector<char> &live, int n_live) {
  const Factor &f = sim.factors[fi];
  std::vector<int> opened;
  int n_close = 0;
  for (int c : sim.f_checks[fi]) {
    if (!live[c]) opened.push_back(c);
    if (remaining[c] == 1 && sim.checks[c].kind == 0) ++n_close;
  }
  const int w_out = n_live + (int)opened.size() - n_close;
  std::vector<uint64_t> rows;
  for (int v : f.vars) {
    uint64_t m = 0;
    for (size_t k = 0; k < opened.size(); ++k)
      if (sim.in_check(opened[k], v)) m |= (uint64_t)1 << k;
    rows.push_back(m);
  }
  return {w_out, w_out + (int)f.vars.size() - gf2_rank(rows)};
}

static std::pair<int, double> evaluate(const std::vector<int> &order, const Sim &sim) {
  std::vector<int> remaining;
  for (auto &fs : sim.c_factors) remaining.push_back((int)fs.size());
  std::vector<char> live(sim.checks.size(), 0);
  int n_live = 0, wmax = 0;
  double cost = 0.0;
  for (int fi : order) {
    auto wc = step_cost(sim, fi, remaining, live, n_live);
    for (int c : sim.f_checks[fi])
      if (!live[c]) { live[c] = 1; ++n_live; }
    for (int c : sim.f_checks[fi]) {
      remaining[c] -= 1;
      if (remaining[c] == 0 && sim.checks[c].kind == 0 && live[c]) { live[c] = 0; --n_live; }
    }
    wmax = std::max(wmax, wc.first);
    cost += std::ldexp(1.0, wc.second);
  }
  return {wmax, cost};
}

std::pair<int, double> evaluate_order(const std::vector<Factor> &factors, const std::vector<Check> &checks, const std::vector<int> &order) {
  Sim sim(factors, checks);
  return evaluate(order, sim);
}

static double score(int wmax, double cost, int n_steps) {
  return cost + 290.0 * n_steps / (double)(1 << std::max(0, 10 - wmax));
}

static std::vector<int> greedy_order(const Sim &sim, int start) {
  const int nF = (int)sim.factors.size();
  std::vector<int> remaining;
  for (auto &fs : sim.c_factors) remaining.push_back((int)fs.size());
  std::vector<char> live(sim.checks.size(), 0), done(nF, 0);
  int n_live = 0;
  std::vector<int> order;
  std::set<int> cand = {start};
  while ((int)order.size() < nF) {
    int best_fi = -1, bw = 0, bl = 0;
    auto consider = [&](int fi) {
      auto wc = step_cost(sim, fi, remaining, live, n_live);
      if (best_fi < 0 || std::make_tuple(wc.first, wc.second, fi) < std::make_tuple(bw, bl, best_fi)) {
        best_fi = fi; bw = wc.first; bl = wc.second;
      }
    };
    if (!cand.empty()) for (int fi : cand) consider(fi);
    else for (int fi = 0; fi < nF; ++fi) if (!done[fi]) consider(fi);
    const int fi = best_fi;
    order.push_back(fi);
    done[fi] = 1;
    cand.erase(fi);
    for (int c : sim.f_checks[fi])
      if (!live[c]) { live[c] = 1; ++n_live; }
    for (int c : sim.f_checks[fi]) {
      remaining[c] -= 1;
      if (remaining[c] == 0 && sim.checks[c].kind == 0 && live[c]) { live[c] = 0; --n_live; }
    }
    for (int c : sim.f_checks[fi])
      for (int fj : sim.c_factors[c])
        if (!done[fj]) cand.insert(fj);
  }
  return order;
}

// Symmetric eigen-decomposition by cyclic Jacobi rotations (n <= a few hundred): eigenvalues ascending in `w`, the
// eigenvector of w[k] in column k of V (row-major n x n).
static void jacobi_eigh(std::vector<double> A, int n, std::vector<double> &w, std::vector<double> &V) {
  V.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) V[(size_t)i * n + i] = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0;
    for (int i = 0; i < n; ++i)
      for (int j = i + 1; j < n; ++j) off += A[(size_t)i * n + j] * A[(size_t)i * n + j];
    if (off < 1e-22) break;
    for (int p = 0; p < n; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = A[(size_t)p * n + q];
        if (std::fabs(apq) < 1e-300) continue;
        const double app = A[(size_t)p * n + p], aqq = A[(size_t)q * n + q];
        const double theta = (aqq - app) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; ++k) {
          const double akp = A[(size_t)k * n + p], akq = A[(size_t)k * n + q];
          A[(size_t)k * n + p] = c * akp - s * akq;
          A[(size_t)k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {
          const double apk = A[(size_t)p * n + k], aqk = A[(size_t)q * n + k];
          A[(size_t)p * n + k] = c * apk - s * aqk;
          A[(size_t)q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) {
          const double vkp = V[(size_t)k * n + p], vkq = V[(size_t)k * n + q];
          V[(size_t)k * n + p] = c * vkp - s * vkq;
          V[(size_t)k * n + q] = s * vkp + c * vkq;
        }
      }
  }
  std::vector<int> idx(n);
  for (int i = 0; i < n; ++i) idx[i] = i;
  std::sort(idx.begin(), idx.end(), [&](int a, int b) { return A[(size_t)a * n + a] < A[(size_t)b * n + b]; });
  std::vector<double> V2((size_t)n * n);
  w.resize(n);
  for (int k = 0; k < n; ++k) {
    w[k] = A[(size_t)idx[k] * n + idx[k]];
    for (int i = 0; i < n; ++i) V2[(size_t)i * n + k] = V[(size_t)i * n + idx[k]];
  }
  V.swap(V2);
}

// Sweep orders from the Fiedler vector of the check graph (schedule.py:spectral_orders)
static std::vector<std::vector<int>> spectral_orders(const Sim &sim) {
  const int nF = (int)sim.factors.size();
  std::vector<int> ids;
  for (size_t c = 0; c < sim.checks.size(); ++c)
    if (sim.checks[c].kind == 0) ids.push_back((int)c);
  std::vector<std::vector<int>> out;
  const int n = (int)ids.size();
  if (n < 3) return out;
  std::vector<int> loc(sim.checks.size(), -1);
  for (int k = 0; k < n; ++k) loc[ids[k]] = k;
  std::vector<double> A((size_t)n * n, 0.0);
  for (auto &fc0 : sim.f_checks) {
    std::vector<int> fc;
    for (int c : fc0)
      if (loc[c] >= 0) fc.push_back(loc[c]);
    for (int a : fc)
      for (int b : fc)
        if (a != b) A[(size_t)a * n + b] = 1.0;
  }
  std::vector<int> comp(n, -1);
  std::vector<std::vector<int>> comps;
  for (int s0 = 0; s0 < n; ++s0) {
    if (comp[s0] >= 0) continue;
    comp[s0] = (int)comps.size();
    std::vector<int> stack = {s0}, members;
    while (!stack.empty()) {
      const int u = stack.back();
      stack.pop_back();
      members.push_back(u);
      for (int v = 0; v < n; ++v)
        if (A[(size_t)u * n + v] != 0.0 && comp[v] < 0) {
          comp[v] = (int)comps.size();
          stack.push_back(v);
        }
    }
    std::sort(members.begin(), members.end());
    comps.push_back(members);
  }
  std::vector<double> x(n, 0.0);
  double base = 0.0;
  for (auto &members : comps) {
    const int m = (int)members.size();
    std::vector<double> v(m);
    if (m >= 3) {
      std::vector<double> Lp((size_t)m * m, 0.0), w, V;
      for (int i = 0; i < m; ++i) {
        double deg = 0.0;
        for (int j = 0; j < m; ++j) {
          const double a = A[(size_t)members[i] * n + members[j]];
          deg += a;
          if (i != j) Lp[(size_t)i * m + j] = -a;
        }
        Lp[(size_t)i * m + i] = deg;
      }
      jacobi_eigh(Lp, m, w, V);
      int amax = 0;
      for (int i = 0; i < m; ++i) {
        v[i] = V[(size_t)i * m + 1];
        if (std::fabs(v[i]) > std::fabs(v[amax])) amax = i;
      }
      if (v[amax] < 0)
        for (double &t : v) t = -t;
      const double mn = *std::min_element(v.begin(), v.end());
      for (double &t : v) t -= mn;
    } else {
      for (int i = 0; i < m; ++i) v[i] = (double)i;
    }
    double mx = 0.0;
    for (int i = 0; i < m; ++i) {
      x[members[i]] = base + v[i];
      mx = std::max(mx, v[i]);
    }
    base += mx + 1.0;
  }
  for (int sgn = 0; sgn < 2; ++sgn) {
    std::vector<double> hi(nF, 0.0);
    for (int i = 0; i < nF; ++i) {
      bool any = false;
      double h = 0.0;
      for (int c : sim.f_checks[i])
        if (loc[c] >= 0) {
          const double xv = sgn ? -x[loc[c]] : x[loc[c]];
          if (!any || xv > h) h = xv;
          any = true;
        }
      hi[i] = any ? h : 0.0;
    }
    std::vector<int> o(nF);
    for (int i = 0; i < nF; ++i) o[i] = i;
    std::stable_sort(o.begin(), o.end(), [&](int a, int b) { return hi[a] < hi[b] || (hi[a] == hi[b] && a < b); });
    out.push_back(o);
  }
  return out;
}

std::vector<int> choose_order(const std::vector<Factor> &factors, const std::vector<Check> &checks, int max_starts) {
  Sim sim(factors, checks);
  const int nF = (int)factors.size();
  std::vector<std::vector<int>> cands(2);
  for (int i = 0; i < nF; ++i) { cands[0].push_back(i); cands[1].push_back(nF - 1 - i); }
  if (nF > 600) max_starts = 6;
  std::vector<int> deg(nF);
  for (int i = 0; i < nF; ++i) deg[i] = i;
  std::stable_sort(deg.begin(), deg.end(), [&](int a, int b) {
    return sim.f_checks[a].size() < sim.f_checks[b].size() || (sim.f_checks[a].size() == sim.f_checks[b].size() && a < b);
  });
  std::vector<int> starts;
  auto add = [&](int s) { if (!contains(starts, s)) starts.push_back(s); };
  for (int i = 0; i < max_starts / 2 && i < nF; ++i) add(deg[i]);
  add(0);
  add(nF - 1);
  for (int s = 0; s < nF; s += std::max(1, nF / (max_starts / 2))) add(s);
  if ((int)starts.size() > max_starts) starts.resize(max_starts);
  for (int s : starts) cands.push_back(greedy_order(sim, s));
  std::vector<std::pair<int, double>> ev;
  for (auto &o : cands) ev.push_back(evaluate(o, sim));
  int best = 0;
  for (int i = 1; i < (int)cands.size(); ++i) {
    const auto ki = std::make_tuple(score(ev[i].first, ev[i].second, nF), ev[i].first, i);
    const auto kb = std::make_tuple(score(ev[best].first, ev[best].second, nF), ev[best].first, best);
    if (ki < kb) best = i;
  }
  if (ev[best].first > MAX_SMEM_WIDTH) {
    for (auto &o : spectral_orders(sim)) {
      cands.push_back(o);
      ev.push_back(evaluate(o, sim));
    }
    best = 0;
    for (int i = 1; i < (int)cands.size(); ++i)
      if (std::make_tuple(ev[i].first, ev[i].second, i) < std::make_tuple(ev[best].first, ev[best].second, best)) best = i;
  }
  return cands[best];
}

// ---- lower (schedule.py) --------------------------------------------------------------------------------------------------
static void encode_schedule(Schedule &s) {
  const int HDR = TQEC_HDR_INTS;
  s.hdr.assign(s.steps.size() * HDR, 0);
  s.ints.clear();
  s.tables.clear();
  const double zero = s.semiring == TQEC_SEMIRING_MAXPLUS ? -INFINITY : 0.0;
  for (size_t t = 0; t < s.steps.size(); ++t) {
    const Step &st = s.steps[t];
    const int r = (int)st.vars.size(), n_open = (int)st.opened.size(), nk = (int)st.ker.size();
    const int64_t inmask = ((int64_t)1 << st.w_in) - 1;
    int32_t *h = s.hdr.data() + t * HDR;
    h[TQEC_H_R] = r; h[TQEC_H_WIN] = st.w_in; h[TQEC_H_NOPEN] = n_open; h[TQEC_H_NCLOSE] = (int)st.closed.size();
    h[TQEC_H_WOUT] = st.w_out; h[TQEC_H_NK] = nk;
    int kb = 0;
    while ((1 << kb) < nk) ++kb;
    h[TQEC_H_KB] = kb;
    if ((1 << kb) != nk) fail("candidate count is not a power of two");
    h[TQEC_H_OFF_T] = (int32_t)s.tables.size();
    for (int p = 0; p < (1 << n_open); ++p)
      for (int k = 0; k < nk; ++k) s.tables.push_back(st.a0[p] >= 0 ? st.table[(size_t)(st.a0[p] ^ st.ker[k])] : zero);
    h[TQEC_H_OFF_ML] = (int32_t)s.ints.size();
    for (int p = 0; p < (1 << n_open); ++p) s.ints.push_back(st.a0[p] >= 0 ? (int32_t)(st.M[(size_t)st.a0[p]] & inmask) : 0);
    h[TQEC_H_OFF_MK] = (int32_t)s.ints.size();
    for (int k = 0; k < nk; ++k) s.ints.push_back((int32_t)(st.M[(size_t)st.ker[k]] & inmask));
    h[TQEC_H_OFF_A0] = (int32_t)s.ints.size();
    for (int p = 0; p < (1 << n_open); ++p) s.ints.push_back(st.a0[p] >= 0 ? (int32_t)st.a0[p] : 0);
    h[TQEC_H_OFF_KER] = (int32_t)s.ints.size();
    for (int k = 0; k < nk; ++k) s.ints.push_back((int32_t)st.ker[k]);
    h[TQEC_H_OFF_VARS] = (int32_t)s.ints.size();
    for (int v : st.vars) s.ints.push_back(v);
    h[TQEC_H_OFF_CLOSE] = (int32_t)s.ints.size();
    for (auto &c : st.closed) { s.ints.push_back(c.first); s.ints.push_back(c.second); }
    for (int x : st.perm) s.ints.push_back(x);
  }
  if (s.ints.empty()) s.ints.push_back(0);
  if (s.tables.empty()) s.tables.push_back(0.0);
}

struct MergedPairs {
  std::vector<Factor> factors;
  std::vector<int> order;
  std::vector<std::pair<int, int>> pairs;  // (-1, -1) = single
};

static MergedPairs merge_pairs(const std::vector<Factor> &factors, const std::vector<int> &order, const std::set<std::pair<int, int>> &forbidden) {
  MergedPairs out;
  size_t k = 0;
  while (k < order.size()) {
    const int a = order[k];
    const int b = k + 1 < order.size() ? order[k + 1] : -1;
    if (b >= 0 && !forbidden.count({a, b}) && factors[a].vars.size() + factors[b].vars.size() <= 4) {
      const Factor &f = factors[a], &g = factors[b];
      const int r = (int)f.vars.size();
      Factor m;
      m.vars = f.vars;
      m.vars.insert(m.vars.end(), g.vars.begin(), g.vars.end());
      m.table.resize((size_t)1 << m.vars.size());
      for (size_t idx = 0; idx < m.table.size(); ++idx) m.table[idx] = f.table[idx & (((size_t)1 << r) - 1)] * g.table[idx >> r];
      out.factors.push_back(m);
      out.pairs.push_back({a, b});
      k += 2;
    } else {
      out.factors.push_back(factors[a]);
      out.pairs.push_back({-1, -1});
      k += 1;
    }
  }
  for (size_t i = 0; i < out.factors.size(); ++i) out.order.push_back((int)i);
  return out;
}

static Schedule lower_impl(const std::vector<Factor> &factors_in, const std::vector<Check> &checks_in, int semiring, int n_vars,
                           int n_checks, int n_obs, const std::vector<int> *order_in, int max_width, int fuse, bool split, bool stable) {
  // signed factors (Clifford-network inference) are legal for sum-product plans
  std::vector<Factor> factors = merge_overlapping(factors_in, n_vars, checks_in, semiring == TQEC_SEMIRING_SUMPROD);
  std::vector<Check> checks;
  for (auto &c : checks_in) {
    Check d;
    d.kind = c.kind;
    d.index = c.index;
    for (int v : c.vars)
      if (!contains(d.vars, v)) d.vars.push_back(v);
    if (c.kind != 0 && c.kind != 1) fail("unknown check kind");
    checks.push_back(d);
  }
  std::vector<int> order;
  if (order_in) order = *order_in;
  else order = choose_order(factors, checks);
  if (order.size() == factors_in.size() && factors_in.size() != factors.size()) order = map_order(factors_in, factors, order);
  {
    std::vector<int> so(order);
    std::sort(so.begin(), so.end());
    bool ok = so.size() == factors.size();
    for (size_t i = 0; ok && i < so.size(); ++i) ok = so[i] == (int)i;
    if (!ok) fail("order must be a permutation of the (merged) factors");
  }
  if (stable) fuse = 0;
  if (fuse < 0) fuse = (semiring == TQEC_SEMIRING_MAXPLUS && std::getenv("TQEC_NO_FUSE") == nullptr) ? 1 : 0;
  if (fuse && !split) {
    std::set<std::pair<int, int>> forbidden;
    bool have_best = false;
    Schedule best;
    for (size_t it = 0; it < 4 * order.size() + 4; ++it) {
      MergedPairs mp = merge_pairs(factors, order, forbidden);
      bool any_pair = false;
      for (auto &p : mp.pairs) any_pair = any_pair || p.first >= 0;
      if (!any_pair) break;
      Schedule trial = lower_impl(mp.factors, checks, semiring, n_vars, n_checks, n_obs, &mp.order, max_width, 1, true, false);
      bool bad = false;
      std::pair<int, int> first_bad;
      for (size_t i = 0; i < trial.steps.size() && !bad; ++i)
        if (mp.pairs[i].first >= 0 && !(trial.steps[i].quad && trial.steps[i].w_out == 9)) { bad = true; first_bad = mp.pairs[i]; }
      if (!bad) { best = trial; have_best = true; break; }
      forbidden.insert(first_bad);
    }
    if (have_best) {
      bool anyq = false;
      for (auto &st : best.steps) anyq = anyq || st.quad;
      if (anyq) return best;
    }
  }
  Sim sim(factors, checks);
  std::vector<int> remaining;
  for (auto &fs : sim.c_factors) remaining.push_back((int)fs.size());
  std::vector<int> orphan;
  for (size_t ci = 0; ci < checks.size(); ++ci)
    if (sim.c_factors[ci].empty()) orphan.push_back((int)ci);
  struct Role { std::vector<int> touched, opened, closing; };
  std::vector<Role> plan;
  std::set<int> seen;
  for (size_t t = 0; t < order.size(); ++t) {
    const int fi = order[t];
    Role r;
    r.touched = sim.f_checks[fi];
    if (t == 0) r.touched.insert(r.touched.end(), orphan.begin(), orphan.end());
    for (int c : r.touched)
      if (!seen.count(c)) r.opened.push_back(c);
    seen.insert(r.opened.begin(), r.opened.end());
    for (int c : r.touched) {
      if (contains(orphan, c)) {
        if (checks[c].kind == 0) r.closing.push_back(c);
        continue;
      }
      remaining[c] -= 1;
      if (remaining[c] == 0 && checks[c].kind == 0) r.closing.push_back(c);
    }
    plan.push_back(r);
  }

  Schedule S;
  S.semiring = semiring; S.n_vars = n_vars; S.n_checks = n_checks; S.n_obs = n_obs;
  std::vector<int> live;
  double cost = 0.0, log2_run = 0.0;
  int wmax = 0, log2_scale = 0;
  for (size_t t = 0; t < order.size(); ++t) {
    const int fi = order[t];
    const Factor &f = factors[fi];
    const int r = (int)f.vars.size();
    const Role &R = plan[t];
    const int w_in = (int)live.size();
    std::vector<int> full = live;
    full.insert(full.end(), R.opened.begin(), R.opened.end());
    if ((int)full.size() > MAX_WIDE_WIDTH)
      fail("frontier needs " + std::to_string(full.size()) + " bits > 31: no executor holds such a state (choose a sweep-like absorption order)");
    std::map<int, int> pos;
    for (size_t k = 0; k < full.size(); ++k) pos[full[k]] = (int)k;
    std::vector<std::pair<int, int>> closed;
    for (int c : R.closing) closed.push_back({pos[c], checks[c].index});
    std::sort(closed.begin(), closed.end());
    std::vector<int64_t> m;
    for (int v : f.vars) {
      int64_t mv = 0;
      for (int c : sim.f_checks[fi])
        if (sim.in_check(c, v)) mv |= (int64_t)1 << pos[c];
      m.push_back(mv);
    }
    const int NA = 1 << r;
    std::vector<int64_t> M(NA, 0), pat(NA);
    for (int a = 0; a < NA; ++a) {
      for (int j = 0; j < r; ++j)
        if ((a >> j) & 1) M[a] ^= m[j];
      pat[a] = M[a] >> w_in;
    }
    const int n_open = (int)R.opened.size();
    std::vector<int64_t> a0((size_t)1 << n_open, -1), ker;
    for (int a = NA - 1; a >= 0; --a) a0[(size_t)pat[a]] = a;
    for (int a = 0; a < NA; ++a)
      if (pat[a] == 0) ker.push_back(a);
    std::vector<int> kept_old;
    for (int c : live)
      if (!contains(R.closing, c)) kept_old.push_back(c);
    int64_t kmask = 0;
    for (int64_t k : ker) kmask |= M[(size_t)k];
    std::vector<int> km;
    for (int c : kept_old)
      if ((kmask >> pos[c]) & 1) km.push_back(c);
    bool quad = false;
    int q5 = -1, q6 = -1;       // -1 = None
    bool q5_none = true, q6_none = true;
    if (ker.size() == 4 && n_open == 2 && r <= 6) {
      int64_t cmask = 0;
      for (auto &c : closed) cmask |= (int64_t)1 << c.first;
      const int64_t inmask = ((int64_t)1 << w_in) - 1;
      std::vector<int> kc;
      for (int a = 0; a < NA; ++a)
        if ((M[a] & cmask) == 0) kc.push_back(a);
      std::map<int, int> reps;
      for (int a : kc) reps[(int)pat[a]] = a;
      if (kc.size() == 4 && reps.size() == 4) {
        const int64_t c1 = M[reps[1]] & inmask, c2 = M[reps[2]] & inmask;
        const bool ok1 = c1 == 0 || (c1 & (c1 - 1)) == 0, ok2 = c2 == 0 || (c2 & (c2 - 1)) == 0;
        if (ok1 && ok2 && (c1 || c2) && c1 != c2) {
          auto blen = [](int64_t x) { int b = 0; while (x >> b) ++b; return b; };
          q5_none = c1 == 0; q6_none = c2 == 0;
          q5 = q5_none ? -1 : full[blen(c1) - 1];
          q6 = q6_none ? -1 : full[blen(c2) - 1];
          bool all_ok = true;
          if (!q5_none && !contains(kept_old, q5)) all_ok = false;
          if (!q6_none && !contains(kept_old, q6)) all_ok = false;
          if (all_ok) {
            quad = true;
            for (int pp = 0; pp < 4; ++pp) a0[pp] = reps[pp];
            km.clear();
            if (!q5_none) km.push_back(q5);
            if (!q6_none) km.push_back(q6);
          } else {
            q5_none = q6_none = true;
          }
        }
      }
    }
    std::set<int> nxt;
    if (t + 1 < order.size()) nxt.insert(plan[t + 1].closing.begin(), plan[t + 1].closing.end());
    std::vector<int> cn, others;
    for (int c : kept_old)
      if (nxt.count(c) && !contains(km, c)) cn.push_back(c);
    for (int c : kept_old)
      if (!contains(km, c) && !contains(cn, c)) others.push_back(c);
    std::vector<int> out_old;
    if (stable) {
      quad = false;
      out_old = kept_old;
    } else if (quad && (int)(others.size() + cn.size()) >= 5 + (q5_none ? 1 : 0) + (q6_none ? 1 : 0)) {
      std::vector<int> pool = others;
      pool.insert(pool.end(), cn.begin(), cn.end());
      std::vector<int> low(pool.begin(), pool.begin() + 5);
      pool.erase(pool.begin(), pool.begin() + 5);
      int b5, b6;
      if (!q5_none) b5 = q5; else { b5 = pool.front(); pool.erase(pool.begin()); }
      if (!q6_none) b6 = q6; else { b6 = pool.front(); pool.erase(pool.begin()); }
      out_old = low;
      out_old.push_back(b5);
      out_old.push_back(b6);
      for (int c : pool) if (contains(cn, c)) out_old.push_back(c);
      for (int c : pool) if (!contains(cn, c)) out_old.push_back(c);
    } else if (others.size() >= 5) {
      quad = false;
      out_old.assign(others.begin(), others.begin() + 5);
      out_old.insert(out_old.end(), km.begin(), km.end());
      out_old.insert(out_old.end(), cn.begin(), cn.end());
      out_old.insert(out_old.end(), others.begin() + 5, others.end());
    } else {
      quad = false;
      out_old = others;
      out_old.insert(out_old.end(), cn.begin(), cn.end());
      out_old.insert(out_old.end(), km.begin(), km.end());
    }
    live = out_old;
    for (int c : R.opened)
      if (!contains(R.closing, c)) live.push_back(c);
    Step st;
    st.factor = fi; st.vars = f.vars; st.w_in = w_in; st.w_out = (int)live.size();
    st.opened = R.opened; st.closed = closed;
    for (int c : live) st.perm.push_back(pos[c]);
    st.M = M; st.a0 = a0; st.ker = ker; st.quad = quad;
    st.table = f.table;
    if (semiring == TQEC_SEMIRING_MAXPLUS) {
      for (double &x : st.table) x = std::log(x);
    } else {
      double mx = 0.0;
      for (double x : st.table) mx = std::max(mx, std::fabs(x));
      if (mx > 0.0) {
        log2_run += std::log2(mx);
        const int e = (int)std::nearbyint(log2_run);
        log2_run -= e;
        for (double &x : st.table) x = std::ldexp(x, -e);
        log2_scale += e;
      }
    }
    cost += std::ldexp(1.0, st.w_out) * (double)ker.size();
    wmax = std::max(wmax, std::max(w_in, st.w_out));
    S.steps.push_back(st);
  }
  if (wmax > max_width) fail("frontier needs " + std::to_string(wmax) + " bits > " + std::to_string(max_width) + ": the schedule does not fit the on-chip state");
  S.obs_slot.assign(n_obs, -1);
  for (size_t k = 0; k < live.size(); ++k) {
    if (checks[live[k]].kind != 1) fail("a clamped check survived the sweep");
    if (checks[live[k]].index < 0 || checks[live[k]].index >= n_obs) fail("observable index out of range");
    S.obs_slot[checks[live[k]].index] = (int)k;
  }
  for (int s : S.obs_slot)
    if (s < 0) fail("every observable row must be declared exactly once");
  if ((int)live.size() != n_obs) fail("every observable row must be declared exactly once");
  S.order = order; S.factors = factors; S.checks = checks; S.w_max = wmax; S.cost = cost; S.log2_scale = log2_scale;
  encode_schedule(S);
  return S;
}

Schedule lower_schedule(const std::vector<Factor> &factors, const std::vector<Check> &checks, int semiring, int n_vars,
                        int n_checks, int n_obs, const std::vector<int> *order, int max_width, int fuse, bool stable) {
  return lower_impl(factors, checks, semiring, n_vars, n_checks, n_obs, order, max_width, fuse, false, stable);
}

}  // namespace lower
}  // namespace tqec
