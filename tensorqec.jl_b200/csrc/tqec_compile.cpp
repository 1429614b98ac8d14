// C ABI of the lowering: tqec_lower / tqec_lowered_get / tqec_plan_from_lowered / tqec_plan_compile (include/tqec.h).
// Chooses the lowering the way tensorqec.jl_b200/decoding.py does:
//   max-plus (TNMAP)    : unfused schedule + in-place patch sweep when the frontier has 5..10 bits and every step fits a
//                         compiled shape; otherwise the fused schedule for the general kernels;
//   sum-product (TNMMAP): schedule (+ sweep when it fits) up to 13 bits, global-memory passes beyond.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

#include "tqec_common.h"
#include "tqec_lower.h"

using namespace tqec;
using namespace tqec::lower;

struct tqec_lowered {
  int kind = 0;            // 0 schedule, 1 schedule + sweep, 2 wide
  int semiring = 0, n_vars = 0, n_checks = 0, n_obs = 0, table_bits = 0, plan_flags = 0;
  Schedule sch;
  SweepPlan sw;
  WidePlan wd;
  std::vector<int32_t> meta, order32, obs32, head_bits32, out_index32;
  std::vector<double> cost, bf_scale;
};

static int env_int(const char *name, int dflt) {
  const char *e = std::getenv(name);
  return e ? std::atoi(e) : dflt;
}

static void read_problem(const tqec_problem_desc *d, Problem &P) {
  if (!d) throw std::runtime_error("problem descriptor is NULL");
  if (d->semiring != TQEC_SEMIRING_MAXPLUS && d->semiring != TQEC_SEMIRING_SUMPROD) throw std::runtime_error("unknown semiring");
  if (d->n_vars < 0 || d->n_checks < 0 || d->n_obs < 0 || d->n_obs > 16) throw std::runtime_error("bad sizes");
  if (d->semiring == TQEC_SEMIRING_MAXPLUS && d->n_obs != 0) throw std::runtime_error("max-plus plans have no open axes");
  if (d->n_factors < 0 || (d->n_factors > 0 && (!d->factor_ptr || !d->factor_vars || !d->factor_tables)))
    throw std::runtime_error("factor arrays missing");
  if (d->n_rows < 0 || (d->n_rows > 0 && (!d->row_ptr || !d->row_kind || !d->row_index))) throw std::runtime_error("row arrays missing");
  P.semiring = d->semiring; P.n_vars = d->n_vars; P.n_checks = d->n_checks; P.n_obs = d->n_obs;
  size_t toff = 0;
  for (int f = 0; f < d->n_factors; ++f) {
    const int a = d->factor_ptr[f], b = d->factor_ptr[f + 1];
    if (a < 0 || b < a || b - a > 10) throw std::runtime_error("factor " + std::to_string(f) + ": bad variable range (rank <= 10)");
    Factor F;
    for (int k = a; k < b; ++k) {
      const int v = d->factor_vars[k];
      if (v < 0 || v >= d->n_vars) throw std::runtime_error("factor " + std::to_string(f) + ": variable id out of range");
      F.vars.push_back(v);
    }
    F.table.assign(d->factor_tables + toff, d->factor_tables + toff + ((size_t)1 << (b - a)));
    toff += (size_t)1 << (b - a);
    P.factors.push_back(F);
  }
  int n_syn = 0, n_obs = 0;
  for (int r = 0; r < d->n_rows; ++r) {
    const int a = d->row_ptr[r], b = d->row_ptr[r + 1];
    if (a < 0 || b < a) throw std::runtime_error("row " + std::to_string(r) + ": bad variable range");
    Check C;
    C.kind = d->row_kind[r]; C.index = d->row_index[r];
    if (C.kind == 0) { if (C.index < 0 || C.index >= d->n_checks) throw std::runtime_error("row " + std::to_string(r) + ": syndrome bit out of range"); ++n_syn; }
    else if (C.kind == 1) { if (C.index < 0 || C.index >= d->n_obs) throw std::runtime_error("row " + std::to_string(r) + ": observable index out of range"); ++n_obs; }
    else throw std::runtime_error("row " + std::to_string(r) + ": unknown kind");
    for (int k = a; k < b; ++k) {
      const int v = d->row_vars[k];
      if (v < 0 || v >= d->n_vars) throw std::runtime_error("row " + std::to_string(r) + ": variable id out of range");
      C.vars.push_back(v);
    }
    P.checks.push_back(C);
  }
  if (n_obs != d->n_obs) throw std::runtime_error("every observable row must be declared exactly once");
  if (d->order) {
    P.has_order = true;
    P.order.assign(d->order, d->order + d->n_factors);
  }
}

static void lower_problem(const tqec_problem_desc *d, tqec_lowered &L) {
  Problem P;
  read_problem(d, P);
  L.semiring = P.semiring; L.n_vars = P.n_vars; L.n_checks = P.n_checks; L.n_obs = P.n_obs; L.table_bits = d->table_bits;
  const bool no_sweep = (d->flags & TQEC_COMPILE_NO_SWEEP) || std::getenv("TQEC_NO_SWEEP");
  const std::vector<int> *order = P.has_order ? &P.order : nullptr;
  if (P.semiring == TQEC_SEMIRING_MAXPLUS) {
    const int head_bits = env_int("TQEC_HEAD_BITS", d->head_bits > 0 ? d->head_bits : 14);
    if (!no_sweep) {
      bool ok = false;
      Schedule su;
      try {
        su = lower_schedule(P.factors, P.checks, P.semiring, P.n_vars, P.n_checks, 0, order, 13, 0, false);
        if (env_int("TQEC_SWEEP_MINW", 5) <= su.w_max && su.w_max <= 10) ok = lower_sweep(su, head_bits, L.sw);
      } catch (const std::runtime_error &) {
        ok = false;
      }
      if (ok) { L.kind = 1; L.sch = su; return; }
    }
    const int fuse = (d->flags & TQEC_COMPILE_NO_FUSE) ? 0 : -1;
    L.sch = lower_schedule(P.factors, P.checks, P.semiring, P.n_vars, P.n_checks, 0, order, 13, fuse, false);
    L.kind = 0;
    return;
  }
  // sum-product: one order for either executor
  std::vector<Factor> merged = merge_overlapping(P.factors, P.n_vars, P.checks, true);
  std::vector<Check> checks;
  for (auto &c : P.checks) {
    Check q;
    q.kind = c.kind; q.index = c.index;
    for (int v : c.vars)
      if (std::find(q.vars.begin(), q.vars.end(), v) == q.vars.end()) q.vars.push_back(v);
    checks.push_back(q);
  }
  std::vector<int> ord = P.has_order ? map_order(P.factors, merged, P.order) : choose_order(merged, checks);
  const int w_max = evaluate_order(merged, checks, ord).first;
  const bool dynamic = (d->flags & TQEC_COMPILE_DYNAMIC_RESCALE) != 0;
  if (dynamic) L.plan_flags |= TQEC_PLAN_DYNAMIC_RESCALE;
  const bool force_wide = dynamic || (d->flags & TQEC_COMPILE_FORCE_WIDE) || std::getenv("TQEC_FORCE_WIDE");
  // on chip up to 11 bits; from 12 bits on the global-memory executor's tile kernel is faster (a 12-bit plan is one tile)
  // plans made of rank-1 factors only (detector error models) run as register butterflies there (k_wide_bf): measured
  // faster from 10 bits on (phenomenological d = 5 x 5 rounds, 11 bits: 5.4 M/s against 2.6 M/s on chip)
  bool all_rank1 = true;
  for (auto &f : merged) all_rank1 = all_rank1 && f.vars.size() == 1;
  const int onchip = std::min(env_int("TQEC_SUMPROD_ONCHIP_WIDTH", all_rank1 ? 9 : 11), 13);
  if (w_max <= onchip && !force_wide) {
    L.sch = lower_schedule(merged, checks, P.semiring, P.n_vars, P.n_checks, P.n_obs, &ord, 13, 0, false);
    L.kind = 0;
    if (!no_sweep && L.sch.w_max >= 5 && L.sch.w_max <= 10) {
      bool ok = false;
      try { ok = lower_sweep(L.sch, env_int("TQEC_HEAD_BITS_SP", d->head_bits > 0 ? d->head_bits : 14), L.sw); } catch (const std::runtime_error &) { ok = false; }
      if (ok) L.kind = 1;
    }
    return;
  }
  const int t_max = env_int("TQEC_WIDE_TMAX", d->wide_t_max > 0 ? d->wide_t_max : 12);
  // dynamic rescaling acts between passes: keep the worst-case drop inside one pass below 2^-600
  L.wd = lower_wide(merged, checks, P.semiring, P.n_vars, P.n_checks, P.n_obs, &ord, t_max, 4, dynamic ? 600.0 : 0.0);
  L.kind = 2;
}

static void finish(tqec_lowered &L) {
  L.meta.assign(20, 0);
  L.meta[0] = L.kind;
  if (L.kind == 2) {
    L.meta[1] = L.wd.n_steps; L.meta[2] = L.wd.w_peak; L.meta[3] = L.wd.log2_scale;
    L.meta[11] = L.wd.n_pass; L.meta[12] = L.wd.n_steps; L.meta[13] = L.wd.w_cap; L.meta[14] = L.wd.t_max;
    L.cost = {L.wd.cost, L.wd.bytes_per_shot};
    L.bf_scale = {L.wd.bf_mant, (double)L.wd.bf_log2};
    L.order32.assign(L.wd.order.begin(), L.wd.order.end());
    L.obs32.assign(L.wd.obs_pos.begin(), L.wd.obs_pos.end());
  } else {
    L.meta[1] = (int32_t)L.sch.steps.size(); L.meta[2] = L.sch.w_max; L.meta[3] = L.sch.log2_scale;
    L.cost = {L.sch.cost, 0.0};
    L.order32.assign(L.sch.order.begin(), L.sch.order.end());
    L.obs32.assign(L.sch.obs_slot.begin(), L.sch.obs_slot.end());
    if (L.kind == 1) {
      L.meta[4] = L.sw.W; L.meta[5] = L.sw.sg; L.meta[6] = L.sw.n_ss; L.meta[7] = (int32_t)L.sw.head_bits.size();
      L.meta[8] = L.sw.bp_words; L.meta[9] = L.sw.head_steps; L.meta[10] = L.sw.conflicts;
      L.head_bits32.assign(L.sw.head_bits.begin(), L.sw.head_bits.end());
      L.out_index32.assign(L.sw.out_index.begin(), L.sw.out_index.end());
    }
  }
  L.meta[15] = L.table_bits;
  L.meta[16] = L.n_obs; L.meta[17] = L.n_checks; L.meta[18] = L.n_vars; L.meta[19] = L.semiring;
  if (L.obs32.empty()) L.obs32.push_back(0);
  if (L.head_bits32.empty()) L.head_bits32.push_back(0);
}

extern "C" int tqec_lower(const tqec_problem_desc *prob, tqec_lowered **out) {
  tqec::NvtxRange nvtx_range("tqec_lower");
  TQEC_REQUIRE(out != nullptr, "tqec_lower: out is NULL");
  *out = nullptr;
  tqec_lowered *L = new tqec_lowered();
  try {
    lower_problem(prob, *L);
    finish(*L);
  } catch (const std::exception &e) {
    delete L;
    const bool unsupported = std::strstr(e.what(), "frontier needs") != nullptr || std::strstr(e.what(), "tile bits") != nullptr;
    set_error("tqec_lower: %s", e.what());
    return unsupported ? TQEC_ERR_UNSUPPORTED : TQEC_ERR_INVALID;
  }
  *out = L;
  return TQEC_OK;
}

extern "C" int tqec_lowered_destroy(tqec_lowered *lw) {
  delete lw;
  return TQEC_OK;
}

extern "C" int tqec_lowered_get(const tqec_lowered *L, int32_t what, const void **data, int64_t *count) {
  TQEC_REQUIRE(L && data && count, "tqec_lowered_get: NULL argument");
#define LW_RET(vec) do { *data = (vec).data(); *count = (int64_t)(vec).size(); return TQEC_OK; } while (0)
  switch (what) {
    case TQEC_LW_META: LW_RET(L->meta);
    case TQEC_LW_COST: LW_RET(L->cost);
    case TQEC_LW_ORDER: LW_RET(L->order32);
    case TQEC_LW_HDR: LW_RET(L->sch.hdr);
    case TQEC_LW_INTS: LW_RET(L->sch.ints);
    case TQEC_LW_TABLES: LW_RET(L->sch.tables);
    case TQEC_LW_OBS_SLOT: LW_RET(L->obs32);
    case TQEC_LW_SW_REC: LW_RET(L->sw.rec);
    case TQEC_LW_SW_TB: LW_RET(L->sw.tb);
    case TQEC_LW_SW_LANETAB: LW_RET(L->sw.lanetab);
    case TQEC_LW_SW_TVALS: LW_RET(L->sw.tvals);
    case TQEC_LW_SW_HEAD_BITS: LW_RET(L->head_bits32);
    case TQEC_LW_SW_HEAD_STATE: LW_RET(L->sw.head_state);
    case TQEC_LW_SW_HEAD_CFG: LW_RET(L->sw.head_cfg);
    case TQEC_LW_SW_OUT_INDEX: LW_RET(L->out_index32);
    case TQEC_LW_WD_PASS_HDR: LW_RET(L->wd.pass_hdr);
    case TQEC_LW_WD_STEP_HDR: LW_RET(L->wd.step_hdr);
    case TQEC_LW_WD_INTS: LW_RET(L->wd.ints);
    case TQEC_LW_WD_TABLES: LW_RET(L->wd.tables);
    case TQEC_LW_WD_OBS_POS: LW_RET(L->obs32);
    case TQEC_LW_WD_BF_OFF: LW_RET(L->wd.bf_off);
    case TQEC_LW_WD_BF_INTS: LW_RET(L->wd.bf_ints);
    case TQEC_LW_WD_BF_VALS: LW_RET(L->wd.bf_vals);
    case TQEC_LW_WD_BF_SCALE: LW_RET(L->bf_scale);
    default: break;
  }
#undef LW_RET
  set_error("tqec_lowered_get: unknown item %d", what);
  return TQEC_ERR_INVALID;
}

extern "C" int tqec_plan_from_lowered(const tqec_lowered *L, int32_t device, tqec_plan **out) {
  TQEC_REQUIRE(L && out, "tqec_plan_from_lowered: NULL argument");
  tqec_plan_desc d;
  std::memset(&d, 0, sizeof(d));
  d.semiring = L->semiring; d.n_vars = L->n_vars; d.n_checks = L->n_checks; d.n_obs = L->n_obs;
  d.device = device; d.table_bits = L->table_bits; d.flags = L->plan_flags;
  tqec_sweep_desc sd;
  tqec_wide_desc wd;
  if (L->kind == 2) {
    std::memset(&wd, 0, sizeof(wd));
    wd.n_pass = L->wd.n_pass; wd.n_steps = L->wd.n_steps; wd.w_cap = L->wd.w_cap; wd.t_max = L->wd.t_max;
    wd.pass_hdr = L->wd.pass_hdr.data(); wd.step_hdr = L->wd.step_hdr.data();
    wd.ints = L->wd.ints.data(); wd.n_ints = (int64_t)L->wd.ints.size();
    wd.tables = L->wd.tables.data(); wd.n_tables = (int64_t)L->wd.tables.size();
    wd.obs_pos = L->obs32.data();
    wd.bf_off = L->wd.bf_off.empty() ? nullptr : L->wd.bf_off.data();
    wd.bf_ints = L->wd.bf_ints.data(); wd.n_bf_ints = (int64_t)L->wd.bf_ints.size();
    wd.bf_vals = L->wd.bf_vals.data(); wd.n_bf_vals = (int64_t)L->wd.bf_vals.size();
    wd.bf_mant = L->wd.bf_mant; wd.bf_log2 = L->wd.bf_log2;
    d.wide = &wd; d.w_max = L->wd.w_cap; d.log2_scale = L->wd.log2_scale;
  } else {
    d.n_steps = (int32_t)L->sch.steps.size(); d.w_max = L->sch.w_max;
    d.hdr = L->sch.hdr.data(); d.ints = L->sch.ints.data(); d.n_ints = (int64_t)L->sch.ints.size();
    d.tables = L->sch.tables.data(); d.n_tables = (int64_t)L->sch.tables.size();
    d.obs_slot = L->obs32.data(); d.log2_scale = L->sch.log2_scale;
    if (L->kind == 1) {
      std::memset(&sd, 0, sizeof(sd));
      sd.W = L->sw.W; sd.sg = L->sw.sg; sd.n_ss = L->sw.n_ss; sd.n_head_bits = (int32_t)L->sw.head_bits.size();
      sd.bp_words = L->sw.bp_words; sd.n_tvals = (int32_t)L->sw.tvals.size();
      sd.rec = L->sw.rec.data(); sd.tb = L->sw.tb.data(); sd.lanetab = L->sw.lanetab.data(); sd.tvals = L->sw.tvals.data();
      sd.head_bits = L->head_bits32.data(); sd.head_state = L->sw.head_state.data(); sd.head_cfg = L->sw.head_cfg.data();
      sd.out_index = L->out_index32.data();
      d.sweep = &sd;
    }
  }
  return tqec_plan_create(&d, out);
}

// ---- plan serialisation: a lowered plan as one file ----------------------------------------------------------------------------
// Everything tqec_plan_from_lowered and tqec_lowered_get read, as tagged little-endian blobs behind a magic word and the
// library's table-format version; a file written by another table format is refused, not reinterpreted.
namespace {
constexpr char LW_MAGIC[8] = {'T', 'Q', 'E', 'C', 'L', 'W', '0', '1'};
constexpr int32_t LW_FORMAT = 3;   // bump whenever a table layout (records, headers, menu ids) changes

struct Writer {
  FILE *f;
  bool ok = true;
  void raw(const void *p, size_t n) { if (ok && n && std::fwrite(p, 1, n, f) != n) ok = false; }
  template <typename T> void scalar(T v) { raw(&v, sizeof(T)); }
  template <typename T> void vec(const std::vector<T> &v) { scalar<int64_t>((int64_t)v.size()); raw(v.data(), v.size() * sizeof(T)); }
};
struct Reader {
  FILE *f;
  bool ok = true;
  void raw(void *p, size_t n) { if (ok && n && std::fread(p, 1, n, f) != n) ok = false; }
  template <typename T> T scalar() { T v{}; raw(&v, sizeof(T)); return v; }
  template <typename T> void vec(std::vector<T> &v) {
    const int64_t n = scalar<int64_t>();
    if (!ok || n < 0 || n > ((int64_t)1 << 34)) { ok = false; return; }
    v.resize((size_t)n);
    raw(v.data(), (size_t)n * sizeof(T));
  }
};
}  // namespace

extern "C" int tqec_lowered_save(const tqec_lowered *L, const char *path) {
  TQEC_REQUIRE(L && path, "tqec_lowered_save: NULL argument");
  FILE *f = std::fopen(path, "wb");
  if (!f) { set_error("tqec_lowered_save: cannot open %s for writing", path); return TQEC_ERR_INVALID; }
  Writer w{f};
  w.raw(LW_MAGIC, 8);
  w.scalar<int32_t>(LW_FORMAT);
  for (int v : {L->kind, L->semiring, L->n_vars, L->n_checks, L->n_obs, L->table_bits, L->plan_flags}) w.scalar<int32_t>(v);
  w.vec(L->meta); w.vec(L->order32); w.vec(L->obs32); w.vec(L->head_bits32); w.vec(L->out_index32); w.vec(L->cost); w.vec(L->bf_scale);
  // schedule (kinds 0 and 1)
  w.scalar<int32_t>((int32_t)L->sch.steps.size()); w.scalar<int32_t>(L->sch.w_max); w.scalar<int32_t>(L->sch.log2_scale);
  w.vec(L->sch.hdr); w.vec(L->sch.ints); w.vec(L->sch.tables);
  // sweep (kind 1)
  for (int v : {L->sw.W, L->sw.sg, L->sw.head_steps, L->sw.n_ss, L->sw.bp_words, L->sw.conflicts}) w.scalar<int32_t>(v);
  w.vec(L->sw.head_bits); w.vec(L->sw.head_state); w.vec(L->sw.head_cfg); w.vec(L->sw.out_index);
  w.vec(L->sw.rec); w.vec(L->sw.tb); w.vec(L->sw.lanetab); w.vec(L->sw.tvals);
  // global-memory passes (kind 2)
  for (int v : {L->wd.w_cap, L->wd.w_peak, L->wd.t_max, L->wd.n_pass, L->wd.n_steps, L->wd.log2_scale, L->wd.bf_log2}) w.scalar<int32_t>(v);
  w.scalar<double>(L->wd.bf_mant);
  w.vec(L->wd.pass_hdr); w.vec(L->wd.step_hdr); w.vec(L->wd.ints); w.vec(L->wd.tables);
  w.vec(L->wd.bf_off); w.vec(L->wd.bf_ints); w.vec(L->wd.bf_vals);
  w.raw(LW_MAGIC, 8);                                            // trailer: a truncated file does not load
  const bool closed = std::fclose(f) == 0;
  if (!w.ok || !closed) { std::remove(path); set_error("tqec_lowered_save: write to %s failed", path); return TQEC_ERR_INVALID; }
  return TQEC_OK;
}

extern "C" int tqec_lowered_load(const char *path, tqec_lowered **out) {
  TQEC_REQUIRE(path && out, "tqec_lowered_load: NULL argument");
  *out = nullptr;
  FILE *f = std::fopen(path, "rb");
  if (!f) { set_error("tqec_lowered_load: cannot open %s", path); return TQEC_ERR_INVALID; }
  Reader r{f};
  char magic[8] = {0};
  r.raw(magic, 8);
  const int32_t fmt = r.scalar<int32_t>();
  if (!r.ok || std::memcmp(magic, LW_MAGIC, 8) != 0 || fmt != LW_FORMAT) {
    std::fclose(f);
    set_error("tqec_lowered_load: %s is not a lowered plan of this library (table format %d expected)", path, (int)LW_FORMAT);
    return TQEC_ERR_INVALID;
  }
  tqec_lowered *L = new tqec_lowered();
  for (int *v : {&L->kind, &L->semiring, &L->n_vars, &L->n_checks, &L->n_obs, &L->table_bits, &L->plan_flags}) *v = r.scalar<int32_t>();
  r.vec(L->meta); r.vec(L->order32); r.vec(L->obs32); r.vec(L->head_bits32); r.vec(L->out_index32); r.vec(L->cost); r.vec(L->bf_scale);
  const int32_t n_steps = r.scalar<int32_t>();
  L->sch.w_max = r.scalar<int32_t>(); L->sch.log2_scale = r.scalar<int32_t>();
  r.vec(L->sch.hdr); r.vec(L->sch.ints); r.vec(L->sch.tables);
  for (int *v : {&L->sw.W, &L->sw.sg, &L->sw.head_steps, &L->sw.n_ss, &L->sw.bp_words, &L->sw.conflicts}) *v = r.scalar<int32_t>();
  r.vec(L->sw.head_bits); r.vec(L->sw.head_state); r.vec(L->sw.head_cfg); r.vec(L->sw.out_index);
  r.vec(L->sw.rec); r.vec(L->sw.tb); r.vec(L->sw.lanetab); r.vec(L->sw.tvals);
  for (int *v : {&L->wd.w_cap, &L->wd.w_peak, &L->wd.t_max, &L->wd.n_pass, &L->wd.n_steps, &L->wd.log2_scale, &L->wd.bf_log2}) *v = r.scalar<int32_t>();
  L->wd.bf_mant = r.scalar<double>();
  r.vec(L->wd.pass_hdr); r.vec(L->wd.step_hdr); r.vec(L->wd.ints); r.vec(L->wd.tables);
  r.vec(L->wd.bf_off); r.vec(L->wd.bf_ints); r.vec(L->wd.bf_vals);
  char trailer[8] = {0};
  r.raw(trailer, 8);
  std::fclose(f);
  const bool sane = r.ok && std::memcmp(trailer, LW_MAGIC, 8) == 0 && L->kind >= 0 && L->kind <= 2 && n_steps >= 0 && n_steps < (1 << 24) &&
                    (L->semiring == TQEC_SEMIRING_MAXPLUS || L->semiring == TQEC_SEMIRING_SUMPROD) && L->n_vars >= 0 && L->n_checks >= 0 &&
                    L->n_obs >= 0 && L->n_obs <= 16 && (int64_t)L->sch.hdr.size() >= (int64_t)n_steps * TQEC_HDR_INTS;
  if (!sane) {
    delete L;
    set_error("tqec_lowered_load: %s is truncated or corrupt", path);
    return TQEC_ERR_INVALID;
  }
  L->sch.steps.resize((size_t)n_steps);                          // only the count is read after the lowering
  L->sch.semiring = L->sw.semiring = L->wd.semiring = L->semiring;
  L->sch.n_vars = L->sw.n_vars = L->wd.n_vars = L->n_vars;
  L->sch.n_checks = L->sw.n_checks = L->wd.n_checks = L->n_checks;
  L->sch.n_obs = L->sw.n_obs = L->wd.n_obs = L->n_obs;
  *out = L;
  return TQEC_OK;
}

extern "C" int tqec_plan_compile(const tqec_problem_desc *prob, tqec_plan **out) {
  TQEC_REQUIRE(prob && out, "tqec_plan_compile: NULL argument");
  *out = nullptr;
  tqec_lowered *L = nullptr;
  int rc = tqec_lower(prob, &L);
  if (rc) return rc;
  rc = tqec_plan_from_lowered(L, prob->device, out);
  tqec_lowered_destroy(L);
  return rc;
}
