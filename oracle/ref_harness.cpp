/*
 * ref_harness.cpp — runs the REFERENCE'S OWN planner sources (TEST INFRASTRUCTURE, never shipped, never imported by the
 * product).
 *
 * What is compiled: this file + /root/reference/motion_primitive_library/src/mpl_planner/map_planner.cpp, against the
 * reference's unmodified headers (mpl_planner/..., mpl_basis/..., mpl_collision/map_util.h) where they lie under
 * /root/reference.  Eigen and Boost are not installed in this image, so oracle/shim/ provides stand-in headers for the few
 * pieces of them the planner uses (dense small matrices, boost::hash_combine, boost::unordered_map,
 * boost::heap::d_ary_heap); see the headers there for what each stand-in restates.  Consequently this is NOT the
 * reference exactly as its authors build it — the third-party parts are restatements — but every line of planner
 * logic (graph_search.h, state_space.h, env_base.h, env_map.h, primitive.h, waypoint.h, map_util.h, map_planner.cpp) is
 * the reference's.  tests/test_oracle_vs_reference.py uses it to validate oracle/mpl_oracle.cpp on configurations for
 * which the reference publishes no numbers (3D, JRK, yaw, search region / potential map, iterativePlan).
 *
 * Built by oracle/Makefile into oracle/_ref/libmplref.so (git-ignored; it travels to the GPU box with the snapshot,
 * /root/reference itself does not).  C interface: the oracle's structs (mpl_oracle.h) and the same call shapes.
 */
#include <mpl_planner/planner/map_planner.h>

#include <fcntl.h>
#include <pthread.h>
#include <sched.h>
#include <unistd.h>

#include <atomic>
#include <chrono>

#include <cstring>
#include <map>
#include <string>
#include <thread>

#include "mpl_oracle.h"

namespace {

/* lattice ints of a waypoint in hash_value's order (waypoint.h:92-125) */
template <int Dim>
int key_ints(const Waypoint<Dim> &w, int32_t *k) {
  int n = 0;
  for (int i = 0; i < Dim; i++) {
    if (w.use_pos) k[n++] = std::round(w.pos(i) / 0.01);
    if (w.use_vel) k[n++] = std::round(w.vel(i) / 0.1);
    if (w.use_acc) k[n++] = std::round(w.acc(i) / 0.1);
    if (w.use_jrk) k[n++] = std::round(w.jrk(i) / 0.1);
  }
  if (w.use_yaw) k[n++] = std::round(w.yaw / 0.1);
  if (w.enable_t) k[n++] = std::round(w.t / 0.1);
  return n;
}
/* same 64-bit mix as the oracle's key_hash() and the CUDA side's khash_* */
uint64_t key_hash(const int32_t *v, int n) {
  uint64_t h = 0x243F6A8885A308D3ull;
  for (int i = 0; i < n; i++) { h ^= (uint64_t)(uint32_t)v[i]; h *= 0x9E3779B97F4A7C15ull; h ^= h >> 32; }
  h ^= h >> 30; h *= 0xBF58476D1CE4E5B9ull;
  h ^= h >> 27; h *= 0x94D049BB133111EBull;
  h ^= h >> 31;
  return h;
}

/* env_map that records what the search asks of it; the search itself is untouched */
template <int Dim>
struct RecEnv : public MPL::env_map<Dim> {
  using MPL::env_map<Dim>::env_map;
  mutable vec_E<Waypoint<Dim>> pops;
  mutable long long n_prims = 0, n_valid = 0;
  void get_succ(const Waypoint<Dim> &curr, vec_E<Waypoint<Dim>> &succ, std::vector<decimal_t> &succ_cost,
                std::vector<int> &action_idx) const {
    if (this->expanded_nodes_.empty()) { pops.clear(); n_prims = 0; n_valid = 0; } /* plan() cleared it (planner_base.h:310): a new search */
    pops.push_back(curr);
    MPL::env_map<Dim>::get_succ(curr, succ, succ_cost, action_idx);
    n_prims += (long long)this->U_.size();
    for (decimal_t c : succ_cost)
      if (!std::isinf(c)) n_valid++;
  }
};

struct IPlanner {
  virtual ~IPlanner() {}
  virtual void set_map(void *map) = 0;
  virtual void set_param(const std::string &k, double v) = 0;
  virtual void set_controls(const double *U, int n, int udim) = 0;
  virtual void set_vec(const std::string &k, const double *v) = 0;
  virtual void set_search_region(const double *path, int n, bool dense) = 0;
  virtual int64_t get_search_region(uint8_t *out, int64_t cap) = 0;
  virtual void update_potential_map(const double *pos) = 0;
  virtual int plan(const orc_waypoint &s, const orc_waypoint &g, orc_result *out) = 0;
  virtual int iterative_plan(const orc_waypoint &s, const orc_waypoint &g, IPlanner *raw, int max_num, orc_result *out) = 0;
  virtual int get_traj_coeffs(double *out, int cap_seg) = 0;
  virtual int get_pop_keys(int32_t *keys16, int cap) = 0;
  virtual int get_nodes(orc_node *nodes, int cap) = 0;
  virtual void set_prior(IPlanner *raw) = 0;
  virtual IPlanner *clone_config() = 0;
  /* LPA* (graph_search.h:194-365, state_space.h:116-282, map_planner.cpp:125-185) */
  virtual void lpa_reset() = 0;
  virtual int lpa_plan(const orc_waypoint &s, const orc_waypoint &g, orc_result *out) = 0;
  virtual int lpa_sub_state_space(int k) = 0;
  virtual int lpa_linked_nodes(double *pts3, int cap) = 0;
  virtual int lpa_update(const int32_t *pns3, int n, bool blocked) = 0;
  virtual int lpa_dump_nodes(orc_lpa_node *nodes, int cap) = 0;
  virtual int lpa_dump_heap(orc_lpa_heap_entry *e, int cap) = 0;
  virtual int lpa_best_child(int32_t *keys16, int cap) = 0;
  virtual int lpa_best_child_states(double *states13, int cap) = 0;
};

template <int Dim>
struct MapHolder {
  std::shared_ptr<MPL::MapUtil<Dim>> mu;
};

template <int Dim>
struct Access : public MPL::MapPlanner<Dim> {
  Access() : MPL::MapPlanner<Dim>(false) {}
  void set_map_recording(const std::shared_ptr<MPL::MapUtil<Dim>> &mu) { /* setMapUtil (map_planner.cpp:14-18) with the recording env */
    this->ENV_.reset(new RecEnv<Dim>(mu));
    this->map_util_ = mu;
  }
  RecEnv<Dim> *env() { return static_cast<RecEnv<Dim> *>(this->ENV_.get()); }
  MPL::StateSpace<Dim, Waypoint<Dim>> *ss() { return this->ss_ptr_.get(); }
  const Trajectory<Dim> &traj() const { return this->traj_; }
  void clear_traj() { this->traj_ = Trajectory<Dim>(); }
};

template <int Dim>
struct PlannerT : public IPlanner {
  std::unique_ptr<Access<Dim>> pl{new Access<Dim>()};
  std::shared_ptr<MPL::MapUtil<Dim>> mu;
  std::map<std::string, double> params;
  vec_E<VecDf> U;
  Vecf<Dim> search_radius = Vecf<Dim>::Zero(), potential_radius = Vecf<Dim>::Zero(), potential_map_range = Vecf<Dim>::Zero();
  orc_result last;
  bool have_traj = false;

  void apply_param(const std::string &k, double v) {
    if (k == "v_max") pl->setVmax(v); else if (k == "a_max") pl->setAmax(v); else if (k == "j_max") pl->setJmax(v);
    else if (k == "yaw_max") pl->setYawmax(v); else if (k == "dt") pl->setDt(v); else if (k == "w") pl->setW(v);
    else if (k == "epsilon") pl->setEpsilon(v); else if (k == "max_num") pl->setMaxNum((int)v);
    else if (k == "wyaw") pl->setWyaw(v);
    else if (k == "potential_weight") pl->setPotentialWeight(v); else if (k == "gradient_weight") pl->setGradientWeight(v);
  }
  void apply_tol() {
    auto get = [&](const char *k, double d) { auto it = params.find(k); return it == params.end() ? d : it->second; };
    pl->setTol(get("tol_pos", 0.5), get("tol_vel", -1), get("tol_acc", -1)); /* planner_base.h:255-265 */
  }
  void set_map(void *map) override {
    mu = static_cast<MapHolder<Dim> *>(map)->mu;
    pl->set_map_recording(mu); /* a fresh environment: re-apply everything, like map_planner_node.cpp:173-182 does once */
    for (auto &kv : params) apply_param(kv.first, kv.second);
    apply_tol();
    if (!U.empty()) pl->setU(U);
  }
  void set_param(const std::string &k, double v) override {
    params[k] = v;
    if (!mu) return;
    if (k.rfind("tol_", 0) == 0) apply_tol(); else apply_param(k, v);
  }
  void set_controls(const double *Uin, int n, int udim) override {
    U.clear();
    for (int i = 0; i < n; i++) {
      VecDf u(udim);
      for (int k = 0; k < udim; k++) u(k) = Uin[(size_t)i * udim + k];
      U.push_back(u);
    }
    if (mu) pl->setU(U);
  }
  void set_vec(const std::string &k, const double *v) override {
    Vecf<Dim> x;
    for (int i = 0; i < Dim; i++) x(i) = v[i];
    if (k == "search_radius") { search_radius = x; pl->setSearchRadius(x); }
    else if (k == "potential_radius") { potential_radius = x; pl->setPotentialRadius(x); }
    else { potential_map_range = x; pl->setPotentialMapRange(x); }
  }
  void set_search_region(const double *path, int n, bool dense) override {
    vec_Vecf<Dim> p;
    for (int i = 0; i < n; i++) { Vecf<Dim> x; for (int k = 0; k < Dim; k++) x(k) = path[(size_t)i * 3 + k]; p.push_back(x); }
    pl->setSearchRegion(p, dense);
  }
  int64_t get_search_region(uint8_t *out, int64_t cap) override {
    const std::vector<bool> r = pl->env()->get_search_region();
    for (int64_t i = 0; i < (int64_t)r.size() && i < cap; i++) out[i] = r[i] ? 1 : 0;
    return (int64_t)r.size();
  }
  void update_potential_map(const double *pos) override {
    Vecf<Dim> x;
    for (int i = 0; i < Dim; i++) x(i) = pos[i];
    pl->updatePotentialMap(x);
  }
  static Waypoint<Dim> to_wp(const orc_waypoint &w) {
    Waypoint<Dim> p;
    for (int i = 0; i < Dim; i++) { p.pos(i) = w.pos[i]; p.vel(i) = w.vel[i]; p.acc(i) = w.acc[i]; p.jrk(i) = w.jrk[i]; }
    p.yaw = w.yaw; p.t = w.t; p.control = (Control::Control)w.control; p.enable_t = w.enable_t != 0;
    return p;
  }
  void fill_result(bool ok, bool start_free, orc_result *out) {
    std::memset(&last, 0, sizeof(last));
    RecEnv<Dim> *env = pl->env();
    have_traj = false;
    if (!start_free) { last.status = 1; last.cost = std::numeric_limits<double>::infinity(); }
    else {
      const bool searched = !env->pops.empty();
      if (ok && !searched) { last.status = 5; last.cost = 0; }       /* graph_search.h:44 */
      else if (ok) { last.status = 0; last.cost = pl->getTrajCost(); last.n_seg = (int)pl->traj().segs.size(); have_traj = true; }
      else { last.status = -1; last.cost = std::numeric_limits<double>::infinity(); } /* the reference's bool does not say why */
      last.pops = (int)env->pops.size();
      last.n_prims = env->n_prims; last.n_valid = env->n_valid; last.n_samples = -1;
      if (searched && pl->ss()) {
        last.n_nodes = (int)pl->ss()->hm_.size();
        last.n_open = (int)pl->ss()->pq_.size();
        uint64_t ch = 0;
        for (const auto &it : pl->ss()->hm_)
          if (it.second && it.second->iterationclosed) { int32_t k[16]; int n = key_ints(it.second->coord, k); ch += key_hash(k, n); last.n_closed++; }
        last.closed_hash = ch;
        uint64_t ph = 0xCBF29CE484222325ull;
        for (const auto &w : env->pops) { int32_t k[16]; int n = key_ints(w, k); ph = (ph ^ key_hash(k, n)) * 0x100000001B3ull; }
        last.pop_hash = ph;
      }
    }
    if (out) *out = last;
  }
  int plan(const orc_waypoint &s, const orc_waypoint &g, orc_result *out) override {
    RecEnv<Dim> *env = pl->env();
    env->pops.clear(); env->n_prims = 0; env->n_valid = 0;
    const Waypoint<Dim> ws = to_wp(s), wg = to_wp(g);
    const bool start_free = env->is_free(ws.pos);
    pl->clear_traj();
    bool ok = false;
    if (start_free) ok = pl->plan(ws, wg);
    fill_result(ok, start_free, out);
    return last.status;
  }
  int iterative_plan(const orc_waypoint &s, const orc_waypoint &g, IPlanner *raw, int max_num, orc_result *out) override {
    RecEnv<Dim> *env = pl->env();
    env->pops.clear(); env->n_prims = 0; env->n_valid = 0;
    const Trajectory<Dim> raw_traj = static_cast<PlannerT<Dim> *>(raw)->pl->traj();
    const bool ok = pl->iterativePlan(to_wp(s), to_wp(g), raw_traj, max_num);
    /* the recording restarts with every inner plan(): what remains is the last iteration, like the planner's own state */
    fill_result(ok, true, out);
    return last.status;
  }
  void set_prior(IPlanner *raw) override { pl->setPriorTrajectory(static_cast<PlannerT<Dim> *>(raw)->pl->traj()); } /* planner_base.h:249-252 */
  int get_traj_coeffs(double *out, int cap_seg) override { /* rows cx, cy, cz, cyaw of toPrimitiveROSMsg, 6 doubles each */
    if (!have_traj) return 0;
    const auto &segs = pl->traj().segs;
    for (int i = 0; i < (int)segs.size() && i < cap_seg; i++) {
      double *o = out + (size_t)i * 24;
      for (int k = 0; k < 24; k++) o[k] = 0;
      for (int ax = 0; ax < Dim; ax++) { const Vec6f c = segs[i].pr(ax).coeff(); for (int k = 0; k < 6; k++) o[ax * 6 + k] = c(k); }
      const Vec6f cy = segs[i].pr_yaw().coeff();
      for (int k = 0; k < 6; k++) o[18 + k] = cy(k);
    }
    return (int)segs.size();
  }
  int get_pop_keys(int32_t *keys16, int cap) override {
    const auto &p = pl->env()->pops;
    for (int i = 0; i < (int)p.size() && i < cap; i++) {
      int32_t *k = keys16 + (size_t)i * 16;
      std::memset(k, 0, 16 * sizeof(int32_t));
      k[15] = key_ints(p[i], k);
    }
    return (int)p.size();
  }
  int get_nodes(orc_node *nodes, int cap) override {
    if (!pl->ss()) return 0;
    int n = 0;
    for (const auto &it : pl->ss()->hm_) {
      if (!it.second) continue;
      if (n < cap) {
        orc_node &o = nodes[n];
        std::memset(&o, 0, sizeof(o));
        const Waypoint<Dim> &w = it.second->coord;
        for (int i = 0; i < Dim; i++) { o.state[i] = w.pos(i); o.state[3 + i] = w.vel(i); o.state[6 + i] = w.acc(i); o.state[9 + i] = w.jrk(i); }
        o.state[12] = w.yaw; o.t = w.t; o.g = it.second->g; o.h = it.second->h;
        o.key[15] = key_ints(w, o.key);
        o.opened = it.second->iterationopened; o.closed = it.second->iterationclosed;
      }
      n++;
    }
    return n;
  }
  /* ---- LPA*: the reference's own setLPAstar / plan / getSubStateSpace / getLinkedNodes / update*Nodes, recorded */
  bool lpa_on = false;
  void lpa_reset() override { pl->reset(); lpa_on = false; }
  int lpa_plan(const orc_waypoint &s, const orc_waypoint &g, orc_result *out) override {
    if (!lpa_on) { pl->setLPAstar(true); lpa_on = true; }
    RecEnv<Dim> *env = pl->env();
    env->pops.clear(); env->n_prims = 0; env->n_valid = 0;
    const Waypoint<Dim> ws = to_wp(s), wg = to_wp(g);
    const bool start_free = env->is_free(ws.pos);
    bool ok = false;
    if (start_free) ok = pl->plan(ws, wg);
    std::memset(&last, 0, sizeof(last));
    have_traj = false;
    last.cost = std::numeric_limits<double>::infinity();
    if (!start_free) last.status = 1;
    else {
      if (ok && pl->getTrajCost() == 0 && env->is_goal(ws)) { last.status = 5; last.cost = 0; } /* graph_search.h:200-205 */
      else if (ok) { last.status = 0; last.cost = pl->getTrajCost(); last.n_seg = (int)pl->traj().segs.size(); have_traj = true; last.pops = pl->ss()->expand_iteration_; }
      else last.status = -1; /* the reference's bool does not say why */
      last.n_prims = env->n_prims; last.n_valid = env->n_valid; last.n_samples = -1;
      if (pl->ss()) {
        last.n_nodes = (int)pl->ss()->hm_.size();
        last.n_open = (int)pl->ss()->pq_.size();
        uint64_t ch = 0;
        for (const auto &it : pl->ss()->hm_)
          if (it.second && it.second->iterationclosed) { int32_t k[16]; int n = key_ints(it.second->coord, k); ch += key_hash(k, n); last.n_closed++; }
        last.closed_hash = ch;
        uint64_t ph = 0xCBF29CE484222325ull; /* over the nodes whose successors were generated in this call, in order */
        for (const auto &w : env->pops) { int32_t k[16]; int n = key_ints(w, k); ph = (ph ^ key_hash(k, n)) * 0x100000001B3ull; }
        last.pop_hash = ph;
      }
    }
    if (out) *out = last;
    return last.status;
  }
  int lpa_sub_state_space(int k) override {
    if (!pl->ss()) return 0;
    pl->getSubStateSpace(k);
    return (int)pl->ss()->hm_.size();
  }
  int lpa_linked_nodes(double *pts3, int cap) override {
    const vec_Vecf<Dim> pts = pl->getLinkedNodes();
    for (int i = 0; i < (int)pts.size() && i < cap; i++) { for (int k = 0; k < 3; k++) pts3[(size_t)i * 3 + k] = 0; for (int k = 0; k < Dim; k++) pts3[(size_t)i * 3 + k] = pts[i](k); }
    return (int)pts.size();
  }
  int lpa_update(const int32_t *pns3, int n, bool blocked) override {
    vec_Veci<Dim> pns;
    for (int i = 0; i < n; i++) { Veci<Dim> v; for (int k = 0; k < Dim; k++) v(k) = pns3[(size_t)i * 3 + k]; pns.push_back(v); }
    if (blocked) pl->updateBlockedNodes(pns); else pl->updateClearedNodes(pns);
    return 0;
  }
  static uint64_t mix(uint64_t h, uint64_t x) { return (h ^ x) * 0x100000001B3ull; }
  static uint64_t bits(double d) { uint64_t u; std::memcpy(&u, &d, 8); return u; }
  int lpa_dump_nodes(orc_lpa_node *nodes, int cap) override {
    if (!pl->ss()) return 0;
    int n = 0;
    for (const auto &it : pl->ss()->hm_) {
      if (n < cap) {
        orc_lpa_node &o = nodes[n];
        std::memset(&o, 0, sizeof(o));
        if (it.second) {
          const auto &st = *it.second;
          o.key[15] = key_ints(st.coord, o.key);
          o.g = st.g; o.rhs = st.rhs; o.h = st.h; o.opened = st.iterationopened; o.closed = st.iterationclosed;
          o.n_succ = (int)st.succ_coord.size(); o.n_pred = (int)st.pred_coord.size();
          uint64_t hs = 0xCBF29CE484222325ull, hp = hs;
          for (size_t i = 0; i < st.succ_coord.size(); i++) { int32_t k[16]; int m = key_ints(st.succ_coord[i], k); hs = mix(mix(mix(hs, key_hash(k, m)), (uint64_t)st.succ_action_id[i]), bits(st.succ_action_cost[i])); }
          for (size_t i = 0; i < st.pred_coord.size(); i++) { int32_t k[16]; int m = key_ints(st.pred_coord[i], k); hp = mix(mix(mix(hp, key_hash(k, m)), (uint64_t)st.pred_action_id[i]), bits(st.pred_action_cost[i])); }
          o.succ_hash = hs; o.pred_hash = hp;
        }
      }
      n++;
    }
    return n;
  }
  int lpa_dump_heap(orc_lpa_heap_entry *e, int cap) override {
    if (!pl->ss()) return 0;
    int n = 0;
    for (const auto &it : pl->ss()->pq_) {
      if (n < cap) { int32_t k[16]; int m = key_ints(it.second->coord, k); e[n].fval = it.first; e[n].key_hash = key_hash(k, m); }
      n++;
    }
    return n;
  }
  int lpa_best_child(int32_t *keys16, int cap) override {
    if (!pl->ss()) return 0;
    const auto &bc = pl->ss()->best_child_;
    for (int i = 0; i < (int)bc.size() && i < cap; i++) { int32_t *k = keys16 + (size_t)i * 16; std::memset(k, 0, 64); k[15] = key_ints(bc[i]->coord, k); }
    return (int)bc.size();
  }
  int lpa_best_child_states(double *states13, int cap) override {
    if (!pl->ss()) return 0;
    const auto &bc = pl->ss()->best_child_;
    for (int i = 0; i < (int)bc.size() && i < cap; i++) {
      double *o = states13 + (size_t)i * 13;
      for (int k = 0; k < 13; k++) o[k] = 0;
      const Waypoint<Dim> &w = bc[i]->coord;
      for (int k = 0; k < Dim; k++) { o[k] = w.pos(k); o[3 + k] = w.vel(k); o[6 + k] = w.acc(k); o[9 + k] = w.jrk(k); }
      o[12] = w.yaw;
    }
    return (int)bc.size();
  }
  IPlanner *clone_config() override { /* same map, parameters, controls and cost shaping; private search state */
    PlannerT<Dim> *c = new PlannerT<Dim>();
    c->params = params; c->U = U;
    MapHolder<Dim> h{mu};
    c->set_map(&h);
    c->pl->env()->set_search_region(pl->env()->get_search_region());
    if (pot_set) { c->pl->env()->set_potential_map(pot_copy); c->pot_set = true; c->pot_copy = pot_copy; }
    return c;
  }
  bool pot_set = false;
  std::vector<int8_t> pot_copy;
};

template <int Dim>
struct PlannerWithPot : public PlannerT<Dim> {
  void update_potential_map(const double *pos) override {
    PlannerT<Dim>::update_potential_map(pos);
    this->pot_set = true;
    this->pot_copy = this->mu->getMap(); /* what updatePotentialMap installed (map_planner.cpp:387-388) */
  }
};

/* graph_search.h:157-160 prints "Priority queue is empty" unconditionally; callers that own stdout (bench.py prints one
 * JSON line) get it silenced for the duration of a call. */
struct QuietStdout {
  int saved = -1;
  QuietStdout() {
    std::fflush(stdout);
    saved = dup(1);
    int nul = open("/dev/null", O_WRONLY);
    if (nul >= 0) { dup2(nul, 1); close(nul); }
  }
  ~QuietStdout() {
    std::fflush(stdout);
    if (saved >= 0) { dup2(saved, 1); close(saved); }
  }
};

struct MapAny {
  int dim;
  MapHolder<2> m2;
  MapHolder<3> m3;
};

}  // namespace

extern "C" {

void *ref_map_create(int dim, const int32_t *nd, const double *origin, double res, const int8_t *data) {
  MapAny *m = new MapAny();
  m->dim = dim;
  size_t n = 1;
  for (int i = 0; i < dim; i++) n *= (size_t)nd[i];
  std::vector<signed char> v(data, data + n);
  if (dim == 2) {
    m->m2.mu = std::make_shared<MPL::MapUtil<2>>();
    m->m2.mu->setMap(Vec2f(origin[0], origin[1]), Vec2i(nd[0], nd[1]), v, res);
  } else {
    m->m3.mu = std::make_shared<MPL::MapUtil<3>>();
    m->m3.mu->setMap(Vec3f(origin[0], origin[1], origin[2]), Vec3i(nd[0], nd[1], nd[2]), v, res);
  }
  return m;
}
void ref_map_destroy(void *map) { delete (MapAny *)map; }
void ref_map_free_unknown(void *map) { MapAny *m = (MapAny *)map; if (m->dim == 2) m->m2.mu->freeUnknown(); else m->m3.mu->freeUnknown(); }
void ref_map_dilate(void *map, const int32_t *ns, int n) { /* MapUtil::dilate, map_util.h:221-257; ns = n rows of Dim ints */
  MapAny *m = (MapAny *)map;
  if (m->dim == 2) {
    vec_Vec2i v;
    for (int i = 0; i < n; i++) v.push_back(Vec2i(ns[i * 2], ns[i * 2 + 1]));
    m->m2.mu->dilate(v);
  } else {
    vec_Vec3i v;
    for (int i = 0; i < n; i++) v.push_back(Vec3i(ns[i * 3], ns[i * 3 + 1], ns[i * 3 + 2]));
    m->m3.mu->dilate(v);
  }
}
int64_t ref_map_get_data(void *map, int8_t *out, int64_t cap) {
  MapAny *m = (MapAny *)map;
  const auto v = m->dim == 2 ? m->m2.mu->getMap() : m->m3.mu->getMap();
  for (int64_t i = 0; i < (int64_t)v.size() && i < cap; i++) out[i] = v[i];
  return (int64_t)v.size();
}

struct RefPlanner {
  int dim;
  IPlanner *p;
};
void *ref_planner_create(int dim) {
  RefPlanner *r = new RefPlanner();
  r->dim = dim;
  r->p = dim == 2 ? (IPlanner *)new PlannerWithPot<2>() : (IPlanner *)new PlannerWithPot<3>();
  return r;
}
void ref_planner_destroy(void *p) { RefPlanner *r = (RefPlanner *)p; delete r->p; delete r; }
void ref_planner_set_map(void *p, void *map) {
  RefPlanner *r = (RefPlanner *)p;
  MapAny *m = (MapAny *)map;
  if (r->dim == 2) r->p->set_map(&m->m2); else r->p->set_map(&m->m3);
}
int ref_planner_set_param(void *p, const char *key, double v) { ((RefPlanner *)p)->p->set_param(key, v); return 0; }
void ref_planner_set_controls(void *p, const double *U, int n, int udim) { ((RefPlanner *)p)->p->set_controls(U, n, udim); }
void ref_planner_set_vec(void *p, const char *key, const double *v) { ((RefPlanner *)p)->p->set_vec(key, v); }
void ref_planner_set_search_region(void *p, const double *path, int n, int dense) { ((RefPlanner *)p)->p->set_search_region(path, n, dense != 0); }
int64_t ref_planner_get_search_region(void *p, uint8_t *out, int64_t cap) { return ((RefPlanner *)p)->p->get_search_region(out, cap); }
void ref_planner_update_potential_map(void *p, const double *pos) { ((RefPlanner *)p)->p->update_potential_map(pos); }
int ref_plan(void *p, const orc_waypoint *s, const orc_waypoint *g, orc_result *out) {
  QuietStdout q;
  return ((RefPlanner *)p)->p->plan(*s, *g, out);
}
int ref_iterative_plan(void *p, void *p_raw, const orc_waypoint *s, const orc_waypoint *g, int max_num, orc_result *out) {
  QuietStdout q;
  return ((RefPlanner *)p)->p->iterative_plan(*s, *g, ((RefPlanner *)p_raw)->p, max_num, out);
}
void ref_planner_set_prior_trajectory(void *p, void *p_raw) {
  QuietStdout q; /* env_map.h:209,217 print every prior cost unconditionally */
  ((RefPlanner *)p)->p->set_prior(((RefPlanner *)p_raw)->p);
}
void ref_map_set_cells(void *map, const int32_t *cells3, int n, int8_t value) { /* getMap / edit / setMap, like map_replanner_node.cpp:181-196 */
  MapAny *m = (MapAny *)map;
  if (m->dim == 2) {
    auto &mu = m->m2.mu; auto d = mu->getMap(); const auto nd = mu->getDim();
    for (int i = 0; i < n; i++) d[cells3[i * 3] + nd(0) * cells3[i * 3 + 1]] = value;
    mu->setMap(mu->getOrigin(), nd, d, mu->getRes());
  } else {
    auto &mu = m->m3.mu; auto d = mu->getMap(); const auto nd = mu->getDim();
    for (int i = 0; i < n; i++) d[cells3[i * 3] + nd(0) * cells3[i * 3 + 1] + nd(0) * nd(1) * cells3[i * 3 + 2]] = value;
    mu->setMap(mu->getOrigin(), nd, d, mu->getRes());
  }
}
void ref_lpa_reset(void *p) { ((RefPlanner *)p)->p->lpa_reset(); }
int ref_lpa_plan(void *p, const orc_waypoint *s, const orc_waypoint *g, orc_result *out) { QuietStdout q; return ((RefPlanner *)p)->p->lpa_plan(*s, *g, out); }
int ref_lpa_get_sub_state_space(void *p, int k) { QuietStdout q; return ((RefPlanner *)p)->p->lpa_sub_state_space(k); }
int ref_lpa_get_linked_nodes(void *p, double *pts3, int cap) { return ((RefPlanner *)p)->p->lpa_linked_nodes(pts3, cap); }
int ref_lpa_update_blocked_nodes(void *p, const int32_t *pns3, int n) { return ((RefPlanner *)p)->p->lpa_update(pns3, n, true); }
int ref_lpa_update_cleared_nodes(void *p, const int32_t *pns3, int n) { return ((RefPlanner *)p)->p->lpa_update(pns3, n, false); }
int ref_lpa_dump_nodes(void *p, orc_lpa_node *nodes, int cap) { return ((RefPlanner *)p)->p->lpa_dump_nodes(nodes, cap); }
int ref_lpa_dump_heap(void *p, orc_lpa_heap_entry *e, int cap) { return ((RefPlanner *)p)->p->lpa_dump_heap(e, cap); }
int ref_lpa_best_child_states(void *p, double *states13, int cap) { return ((RefPlanner *)p)->p->lpa_best_child_states(states13, cap); }
int ref_lpa_best_child(void *p, int32_t *keys16, int cap) { return ((RefPlanner *)p)->p->lpa_best_child(keys16, cap); }
int ref_get_traj_coeffs(void *p, double *out, int cap_seg) { return ((RefPlanner *)p)->p->get_traj_coeffs(out, cap_seg); }
int ref_get_pop_keys(void *p, int32_t *keys16, int cap) { return ((RefPlanner *)p)->p->get_pop_keys(keys16, cap); }
int ref_get_nodes(void *p, orc_node *nodes, int cap) { return ((RefPlanner *)p)->p->get_nodes(nodes, cap); }

/* n independent plans pulled from an atomic work queue (in `order` when given) by nthreads std::threads, one private
 * planner per thread (the reference is single-threaded per plan), threads optionally pinned to cores. */
int ref_plan_batch_dyn(void *p, const orc_waypoint *starts, const orc_waypoint *goals, int n, int nthreads, orc_result *results,
                       const int32_t *order, int pin, double *busy_s) {
  RefPlanner *base = (RefPlanner *)p;
  if (nthreads < 1) nthreads = 1;
  QuietStdout q;
  std::atomic<int> next{0};
  auto worker = [&](int tid) {
    if (pin) {
      cpu_set_t set;
      CPU_ZERO(&set);
      long nc = sysconf(_SC_NPROCESSORS_ONLN);
      CPU_SET((int)(tid % (nc > 0 ? nc : 1)), &set);
      pthread_setaffinity_np(pthread_self(), sizeof(set), &set);
    }
    std::unique_ptr<IPlanner> local(base->p->clone_config());
    double busy = 0.0;
    while (true) {
      const int k = next.fetch_add(1);
      if (k >= n) break;
      const int i = order ? order[k] : k;
      auto t0 = std::chrono::steady_clock::now();
      local->plan(starts[i], goals[i], &results[i]);
      const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      busy += el;
      results[i].device_ms = el * 1e3;
    }
    if (busy_s) busy_s[tid] = busy;
  };
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; t++) th.emplace_back(worker, t);
  for (auto &t : th) t.join();
  return 0;
}
int ref_plan_batch(void *p, const orc_waypoint *starts, const orc_waypoint *goals, int n, int nthreads, orc_result *results) {
  return ref_plan_batch_dyn(p, starts, goals, n, nthreads, results, nullptr, 0, nullptr);
}

}  // extern "C"
