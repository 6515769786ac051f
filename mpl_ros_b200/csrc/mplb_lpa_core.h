/*
 * mplb_lpa_core.h — LPA* (Lifelong Planning A*) replanning core of libmplb.so.
 *
 * What it implements (paths under motion_primitive_library/include/mpl_planner, src/mpl_planner):
 *   GraphSearch::LPAstar                               common/graph_search.h:194-365      (gs)
 *   GraphSearch::recoverTraj                           common/graph_search.h:369-455
 *   StateSpace::getSubStateSpace / increaseCost / decreaseCost / updateNode / calculateKey
 *                                                      common/state_space.h:116-282       (ss)
 *   MapPlanner::getLinkedNodes / updateBlockedNodes / updateClearedNodes   map_planner.cpp:125-185
 *   env_map::get_succ / traverse_primitive / is_free(Primitive) / is_goal  env/env_map.h:25-172 (em)
 * for the plain occupancy map (no potential map / search region / yaw / prior trajectory — those stay A*-only).
 *
 * Shape on the GPU.  One replanner ("session") is one CTA of one warp; a batch of sessions (multi-robot replanning) is a
 * grid.  LPA* is a serial algorithm over a pointer graph; what is parallel inside one session is the successor
 * generation of a popped node (one lane per control: polynomial end state, dynamic validation, lattice key, the
 * collision samples of its primitive), the voxel -> edge link table (one thread per node, count / scan / fill) and the
 * matching of changed voxels against that table (one thread per link).  Everything that decides ORDER (heap, hm_
 * insertion order, updateNode sequence) runs on lane 0 in the reference's statement order, so that the priority queue
 * array, the key ties and therefore the expanded set are the reference's, state by state.
 *
 * Layout.  A node is identified with its lattice key for the whole life of the session: a record that the reference
 * would drop in getSubStateSpace and later re-create through `hm_[coord]` is reset in place (same id).  Successor lists
 * store the successor's node id instead of its coordinate (the coordinate is a function of (parent coord, action) and is
 * recomputed when a dropped successor has to be re-created), predecessor lists are singly linked records in insertion
 * order (recoverTraj's tie rule depends on that order).  hm_ iteration order — which Boost leaves unspecified and which
 * getSubStateSpace / getLinkedNodes observe — is defined as insertion order (`order[]`), like the stand-in container the
 * reference's own sources are compiled against for the parity tests.
 *
 * This header is plain C++ with MPLB_HD functions and no CUDA-only construct outside `#ifdef __CUDA_ARCH__`, so that the
 * test suite can compile the very same statements for the host and compare them with the CPU checker where there is no
 * GPU (tests/cpp/lpa_emul.cpp; test infrastructure, never loaded by the product).  The product path is the kernels of
 * mplb_lpa.cu; there is no CPU fallback.
 *
 * Arithmetic: IEEE double, every operation in the reference's order, explicit round-to-nearest intrinsics on the device
 * (no FMA contraction), plain operators on the host build (-ffp-contract=off).
 */
#ifndef MPLB_LPA_CORE_H
#define MPLB_LPA_CORE_H
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define MPLB_HD __host__ __device__ __forceinline__
#define MPLB_HDN __host__ __device__
#else
#define MPLB_HD inline
#define MPLB_HDN inline
#endif

namespace mplb_lpa {

#define LPA_INF (__builtin_huge_val())
#define LPA_MAXU 128

/* statuses of a plan (same numbers as mplb_result.status) + internal ones */
enum { LPA_OK = 0, LPA_START_NOT_FREE = 1, LPA_MAX_EXPAND = 2, LPA_QUEUE_EMPTY = 3, LPA_TRACEBACK_FAILED = 4, LPA_START_IS_GOAL = 5,
       LPA_NEED_GROW = 100, LPA_FAULT = 101 };

MPLB_HD double fA(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}
MPLB_HD double fS(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dsub_rn(a, b);
#else
  return a - b;
#endif
}
MPLB_HD double fM(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}
MPLB_HD double fD(double a, double b) {
#ifdef __CUDA_ARCH__
  return __ddiv_rn(a, b);
#else
  return a / b;
#endif
}
MPLB_HD double fSqrt(double a) {
#ifdef __CUDA_ARCH__
  return __dsqrt_rn(a);
#else
  return sqrt(a);
#endif
}
/* std::round: half away from zero; x - trunc(x) is exact */
MPLB_HD double fRound(double x) {
  double r = trunc(x);
  if (fabs(fS(x, r)) >= 0.5) r = fA(r, copysign(1.0, x));
  return r;
}
MPLB_HD double fMin(double a, double b) { return b < a ? b : a; } /* std::min */
MPLB_HD double fMax(double a, double b) { return a < b ? b : a; } /* std::max */
MPLB_HD bool fIsInf(double x) { return x == LPA_INF || x == -LPA_INF; }
MPLB_HD double fPower(double t, int n) { /* math.h:197-203 */
  double tn = 1;
  while (n > 0) { tn = fM(tn, t); n--; }
  return tn;
}

/* ------------------------------------------------------------------ configuration and storage */
struct Cfg {
  int dim, ord, control, nU, nkey, max_num;
  double dt, w, eps, v_max, a_max, j_max, tol_pos, tol_vel, tol_acc;
  int nd[3];
  double origin[3];
  double res;
  const int8_t *grid; /* int8 cells, x fastest (map_util.h:33-41) */
  const double *U;    /* nU rows of 3 */
};

struct Node {
  double st[13]; /* pos3 vel3 acc3 jrk3 yaw: the State's coord (first creator's values) */
  double t;
  double g, rhs, h;
  int key[12];
  int heap_pos; /* -1 = not in pq_ */
  int n_succ;   /* stored successor entries; 0 = never expanded (gs:265) */
  int pred_head, pred_tail, n_pred;
  unsigned char opened, closed, in_hm, pad;
};
struct Succ { double cost; int node; int act; };
struct Pred { double cost; int node; int act; int next; int pad; };
struct Link { int vox; int node; int pred_idx; int cell[3]; };

struct Row { /* one control's result of get_succ for the popped node (staging, written by the lane that owns the control) */
  double st[13];
  double t;
  double cost;
  int key[12];
  int verdict; /* 0 = no successor emitted (same state / dynamically infeasible), 1 = emitted */
  int n_samples;
  unsigned long long hash;
};

struct Hdr { /* scalar state of one session */
  int n_nodes, cap_nodes, n_pred, cap_pred, n_order, n_heap, tsize;
  int start_node, goal_node; /* goal_node < 0: the detached State(Coord()) of gs:224-241 (g = rhs = inf, h = 0) */
  int expand_iteration, status, resume, initialized;
  int n_best, n_links, cap_links, n_match, cap_match, n_epq, cap_epq;
  int curr, has_rows; /* the pop in flight between the serial and the parallel half */
  int n_explored, fault;
  double start_g, start_rhs, start_t, eps;
  double goal[13]; /* requested goal (pos vel acc) */
  long long n_prims, n_valid, n_samples;
  unsigned long long pop_hash;
  int start_key[12];
  double start_st[13];
  double start_tt;
};

struct Ctx { /* device view of one session: header + arrays */
  Cfg cfg;
  Hdr *h;
  Node *nodes;
  Succ *succ;   /* cap_nodes * nU */
  Pred *preds;  /* cap_pred */
  int *table;   /* tsize slots: node id or -1 */
  int *order;   /* hm_ iteration order */
  int *order2;  /* scratch of getSubStateSpace */
  double *heap_f;
  int *heap_node;
  int *best;    /* best_child_ (start .. goal) */
  int *traj_act;
  Row *rows;    /* nU */
  double *epq_f; int *epq_node;             /* scratch heap of getSubStateSpace (duplicates allowed) */
  unsigned char *mark;                      /* per node: member of new_hm */
  Link *links; int *link_count;             /* lhm_ as a flat table in insertion order; per order position counts/offsets */
  unsigned long long *match;                /* (changed voxel position, link index) pairs */
};

/* ------------------------------------------------------------------ lattice key (waypoint.h:92-125) */
MPLB_HD unsigned long long key_hash(const int *k, int n) { /* same definition as the checker's key_hash */
  unsigned long long h = 0x243F6A8885A308D3ull;
  for (int i = 0; i < n; i++) {
    h ^= (unsigned long long)(unsigned int)k[i];
    h *= 0x9E3779B97F4A7C15ull;
    h ^= h >> 32;
  }
  h ^= h >> 30; h *= 0xBF58476D1CE4E5B9ull;
  h ^= h >> 27; h *= 0x94D049BB133111EBull;
  h ^= h >> 31;
  return h;
}
MPLB_HD void make_key(const Cfg &c, const double *st, int *key) {
  int n = 0;
  for (int i = 0; i < c.dim; i++) {
    key[n++] = (int)fRound(fD(st[i], 0.01));
    if (c.ord >= 2) key[n++] = (int)fRound(fD(st[3 + i], 0.1));
    if (c.ord >= 3) key[n++] = (int)fRound(fD(st[6 + i], 0.1));
    if (c.ord >= 4) key[n++] = (int)fRound(fD(st[9 + i], 0.1));
  }
  for (; n < 12; n++) key[n] = 0;
}
MPLB_HD bool key_eq(const int *a, const int *b, int n) {
  for (int i = 0; i < n; i++) if (a[i] != b[i]) return false;
  return true;
}

/* ------------------------------------------------------------------ Primitive (primitive.h:21-198,205-431) */
struct Prim1 { double c[6]; };
MPLB_HD double pr_p(const Prim1 &q, double t) { /* pr:128-131 */
  const double *c = q.c;
  double s = fM(fD(c[0], 120), fPower(t, 5));
  s = fA(s, fM(fD(c[1], 24), fPower(t, 4)));
  s = fA(s, fM(fD(c[2], 6), fPower(t, 3)));
  s = fA(s, fM(fM(fD(c[3], 2), t), t));
  s = fA(s, fM(c[4], t));
  return fA(s, c[5]);
}
MPLB_HD double pr_v(const Prim1 &q, double t) { /* pr:134-137 */
  const double *c = q.c;
  double s = fM(fD(c[0], 24), fPower(t, 4));
  s = fA(s, fM(fD(c[1], 6), fPower(t, 3)));
  s = fA(s, fM(fM(fD(c[2], 2), t), t));
  s = fA(s, fM(c[3], t));
  return fA(s, c[4]);
}
MPLB_HD double pr_a(const Prim1 &q, double t) { /* pr:140-142 */
  const double *c = q.c;
  double s = fM(fD(c[0], 6), fPower(t, 3));
  s = fA(s, fM(fM(fD(c[1], 2), t), t));
  s = fA(s, fM(c[2], t));
  return fA(s, c[3]);
}
MPLB_HD double pr_j(const Prim1 &q, double t) { /* pr:145 */
  const double *c = q.c;
  return fA(fA(fM(fM(fD(c[0], 2), t), t), fM(c[1], t)), c[2]);
}
/* solve(0, 0, c, d, e) of math.h:117-131 (a = b = 0 for every control-built primitive: c0 = 0, pr:35-52); roots in the
 * order quad() returns them; returns the count */
MPLB_HD int solve_low(double c, double d, double e, double *r) {
  if (c != 0) { /* quad, math.h:22-33 */
    const double p = fS(fM(d, d), fM(fM(4, c), e));
    if (p < 0) return 0;
    const double sq = fSqrt(p);
    r[0] = fD(fS(-d, sq), fM(2, c));
    r[1] = fD(fA(-d, sq), fM(2, c));
    return 2;
  } else if (d != 0) {
    r[0] = fD(-e, d);
    return 1;
  }
  return 0;
}
/* max over [0, T] of |derivative|: end points and the interior extrema (pr:353-394 with extrema_* pr:152-193) */
MPLB_HD double pr_max(const Prim1 &q, double T, int which) {
  const double *c = q.c;
  double r[2] = {0, 0};
  int nr = 0;
  double m;
  if (which == 1) { nr = solve_low(fD(c[1], 2), c[2], c[3], r); m = fMax(fabs(pr_v(q, 0)), fabs(pr_v(q, T))); }
  else if (which == 2) { nr = solve_low(0, c[1], c[2], r); m = fMax(fabs(pr_a(q, 0)), fabs(pr_a(q, T))); } /* solve(0,0,c0/2,c1,c2) */
  else { nr = 0; m = fMax(fabs(pr_j(q, 0)), fabs(pr_j(q, T))); } /* extrema_j needs c0 != 0 */
  for (int i = 0; i < nr; i++) { /* extrema_*: keep roots in (0, T), stop at the first root >= T */
    const double it = r[i];
    if (it > 0 && it < T) {
      const double x = fabs(which == 1 ? pr_v(q, it) : pr_a(q, it));
      m = x > m ? x : m;
    } else if (it >= T) break;
  }
  return m;
}
MPLB_HD double pr_J(const Prim1 &q, double t, int control) { /* pr:92-122 */
  const double *c = q.c;
  const int cc = control & 15;
  if (cc == 1) {
    double s = fM(fD(fM(c[0], c[0]), 5184), fPower(t, 9));
    s = fA(s, fM(fD(fM(c[0], c[1]), 576), fPower(t, 8)));
    s = fA(s, fM(fA(fD(fM(c[1], c[1]), 252), fD(fM(c[0], c[2]), 168)), fPower(t, 7)));
    s = fA(s, fM(fA(fD(fM(c[0], c[3]), 72), fD(fM(c[1], c[2]), 36)), fPower(t, 6)));
    s = fA(s, fM(fA(fA(fD(fM(c[2], c[2]), 20), fD(fM(c[0], c[4]), 60)), fD(fM(c[1], c[3]), 15)), fPower(t, 5)));
    s = fA(s, fM(fA(fD(fM(c[2], c[3]), 4), fD(fM(c[1], c[4]), 12)), fPower(t, 4)));
    s = fA(s, fM(fA(fD(fM(c[3], c[3]), 3), fD(fM(c[2], c[4]), 3)), fPower(t, 3)));
    s = fA(s, fM(fM(fM(c[3], c[4]), t), t));
    return fA(s, fM(fM(c[4], c[4]), t));
  } else if (cc == 3) {
    double s = fM(fD(fM(c[0], c[0]), 252), fPower(t, 7));
    s = fA(s, fM(fD(fM(c[0], c[1]), 36), fPower(t, 6)));
    s = fA(s, fM(fA(fD(fM(c[1], c[1]), 20), fD(fM(c[0], c[2]), 15)), fPower(t, 5)));
    s = fA(s, fM(fA(fD(fM(c[0], c[3]), 12), fD(fM(c[1], c[2]), 4)), fPower(t, 4)));
    s = fA(s, fM(fA(fD(fM(c[2], c[2]), 3), fD(fM(c[1], c[3]), 3)), fPower(t, 3)));
    s = fA(s, fM(fM(fM(c[2], c[3]), t), t));
    return fA(s, fM(fM(c[3], c[3]), t));
  } else if (cc == 7) {
    double s = fM(fD(fM(c[0], c[0]), 20), fPower(t, 5));
    s = fA(s, fM(fD(fM(c[0], c[1]), 4), fPower(t, 4)));
    s = fA(s, fM(fD(fA(fM(c[1], c[1]), fM(c[0], c[2])), 3), fPower(t, 3)));
    s = fA(s, fM(fM(fM(c[1], c[2]), t), t));
    return fA(s, fM(fM(c[2], c[2]), t));
  } else if (cc == 15) {
    double s = fM(fD(fM(c[0], c[0]), 3), fPower(t, 3));
    s = fA(s, fM(fM(fM(c[0], c[1]), t), t));
    return fA(s, fM(fM(c[1], c[1]), t));
  }
  return 0;
}

struct Prim { Prim1 ax[3]; };
MPLB_HD void prim_build(const Cfg &c, const double *st, const double *u, Prim &pr) { /* pr:220-256 */
  for (int i = 0; i < c.dim; i++) {
    double *k = pr.ax[i].c;
    k[0] = k[1] = k[2] = k[3] = k[4] = k[5] = 0;
    if (c.ord == 4) { k[1] = u[i]; k[2] = st[9 + i]; k[3] = st[6 + i]; k[4] = st[3 + i]; k[5] = st[i]; }
    else if (c.ord == 3) { k[2] = u[i]; k[3] = st[6 + i]; k[4] = st[3 + i]; k[5] = st[i]; }
    else if (c.ord == 2) { k[3] = u[i]; k[4] = st[3 + i]; k[5] = st[i]; }
    else { k[4] = u[i]; k[5] = st[i]; }
  }
}
MPLB_HD void prim_eval(const Cfg &c, const Prim &pr, double t, double *st) { /* pr:321-331 (yaw stays 0: no yaw control here) */
  for (int k = 0; k < 13; k++) st[k] = 0;
  for (int k = 0; k < c.dim; k++) {
    st[k] = pr_p(pr.ax[k], t);
    st[3 + k] = pr_v(pr.ax[k], t);
    st[6 + k] = pr_a(pr.ax[k], t);
    st[9 + k] = pr_j(pr.ax[k], t);
  }
}
MPLB_HD bool validate_xxx(const Cfg &c, const Prim &pr, double mx, int which) { /* pr:483-496 */
  if (mx <= 0) return true;
  for (int i = 0; i < c.dim; i++)
    if (pr_max(pr.ax[i], c.dt, which) > mx) return false;
  return true;
}
MPLB_HD bool validate_primitive(const Cfg &c, const Prim &pr) { /* pr:449-475 */
  if (c.ord == 2) return validate_xxx(c, pr, c.v_max, 1);
  if (c.ord == 3) return validate_xxx(c, pr, c.v_max, 1) && validate_xxx(c, pr, c.a_max, 2);
  if (c.ord == 4) return validate_xxx(c, pr, c.v_max, 1) && validate_xxx(c, pr, c.a_max, 2) && validate_xxx(c, pr, c.j_max, 3);
  return true;
}
MPLB_HD double prim_J(const Cfg &c, const Prim &pr) { /* pr:403-407 */
  double j = 0;
  for (int k = 0; k < c.dim; k++) j = fA(j, pr_J(pr.ax[k], c.dt, c.control));
  return j;
}
MPLB_HD double prim_max_v(const Cfg &c, const Prim &pr) { /* em:91-94 */
  double mv = 0;
  for (int i = 0; i < c.dim; i++) {
    const double x = pr_max(pr.ax[i], c.dt, 1);
    if (x > mv) mv = x;
  }
  return mv;
}

/* ------------------------------------------------------------------ MapUtil (map_util.h) */
MPLB_HD void float_to_int(const Cfg &c, const double *pt, int *pn) { /* mu:103-108 */
  pn[0] = pn[1] = pn[2] = 0;
  for (int i = 0; i < c.dim; i++) pn[i] = (int)fRound(fS(fD(fS(pt[i], c.origin[i]), c.res), 0.5));
}
MPLB_HD bool outside(const Cfg &c, const int *pn) {
  for (int i = 0; i < c.dim; i++) if (pn[i] < 0 || pn[i] >= c.nd[i]) return true;
  return false;
}
MPLB_HD int cell_index(const Cfg &c, const int *pn) { /* mu:33-41, int arithmetic like the reference (no bounds check) */
  return c.dim == 2 ? pn[0] + c.nd[0] * pn[1] : pn[0] + c.nd[0] * pn[1] + c.nd[0] * c.nd[1] * pn[2];
}
MPLB_HD bool occupied(const Cfg &c, const int *pn) { return outside(c, pn) ? false : c.grid[cell_index(c, pn)] == 100; }
MPLB_HD bool cell_free(const Cfg &c, const int *pn) {
  if (outside(c, pn)) return false;
  const int8_t v = c.grid[cell_index(c, pn)];
  return v < 100 && v >= 0;
}
MPLB_HD bool ray_hits_occupied(const Cfg &c, const double *p1, const double *p2) { /* mu:117-134 as em:38-42 uses it */
  double diff[3] = {0, 0, 0}, q = 0;
  for (int i = 0; i < c.dim; i++) {
    diff[i] = fS(p2[i], p1[i]);
    const double a = fabs(fD(diff[i], c.res));
    if (i == 0 || a > q) q = a;
  }
  const int max_diff = (int)fD(q, 0.8);
  const double s = fD(1.0, (double)max_diff);
  double step[3] = {0, 0, 0};
  for (int i = 0; i < c.dim; i++) step[i] = fM(diff[i], s);
  for (int n = 1; n < max_diff; n++) {
    double pt[3] = {0, 0, 0};
    int pn[3];
    for (int i = 0; i < c.dim; i++) pt[i] = fA(p1[i], fM(step[i], (double)n));
    float_to_int(c, pt, pn);
    if (outside(c, pn)) break;
    if (c.grid[cell_index(c, pn)] == 100) return true;
  }
  return false;
}

/* ------------------------------------------------------------------ env_map */
MPLB_HD bool is_goal(const Cfg &c, const double *goal, const double *st) { /* em:25-45 */
  double m = 0;
  for (int i = 0; i < c.dim; i++) m = fMax(m, fabs(fS(st[i], goal[i])));
  bool goaled = m <= c.tol_pos;
  if (goaled && c.tol_vel >= 0) {
    m = 0;
    for (int i = 0; i < c.dim; i++) m = fMax(m, fabs(fS(st[3 + i], goal[3 + i])));
    goaled = m <= c.tol_vel;
  }
  if (goaled && c.tol_acc >= 0) {
    m = 0;
    for (int i = 0; i < c.dim; i++) m = fMax(m, fabs(fS(st[6 + i], goal[6 + i])));
    goaled = m <= c.tol_acc;
  }
  if (goaled && ray_hits_occupied(c, st, goal)) return false;
  return goaled;
}
MPLB_HD double heur(const Cfg &c, const double *goal, const int *goal_key, const double *st, const int *key) { /* eb:46-64 */
  if (key_eq(goal_key, key, c.nkey)) return 0;
  double m = 0;
  for (int i = 0; i < c.dim; i++) m = fMax(m, fabs(fS(st[i], goal[i])));
  if (c.v_max > 0) return fD(fM(c.w, m), c.v_max);
  return fM(c.w, m);
}
/* em:90-132 on the plain map: +inf when a sample is outside or occupied, else 0; samples at the accumulated times */
MPLB_HD double traverse(const Cfg &c, const Prim &pr, int *n_samples) {
  const double max_v = prim_max_v(c, pr);
  int n = (int)ceil(fD(fM(max_v, c.dt), c.res));
  if (n < 5) n = 5;
  const double dts = fD(c.dt, (double)n);
  int tested = 0;
  for (double t = 0; t < c.dt; t = fA(t, dts)) {
    double st[13];
    int pn[3];
    prim_eval(c, pr, t, st);
    float_to_int(c, st, pn);
    tested++;
    if (outside(c, pn) || c.grid[cell_index(c, pn)] == 100) { *n_samples += tested; return LPA_INF; }
  }
  *n_samples += tested;
  return 0;
}
/* em:60-76: Primitive::sample(n) = n + 1 points at i * (T / n), occupied or outside -> false */
MPLB_HD bool prim_is_free(const Cfg &c, const Prim &pr) {
  const double max_v = prim_max_v(c, pr);
  const int n = (int)ceil(fD(fM(max_v, c.dt), c.res));
  const double dts = fD(c.dt, (double)n);
  for (int i = 0; i <= n; i++) {
    double st[13];
    int pn[3];
    prim_eval(c, pr, fM((double)i, dts), st);
    float_to_int(c, st, pn);
    if (occupied(c, pn) || outside(c, pn)) return false;
  }
  return true;
}
/* One control of get_succ (em:147-172) for the node with coord (st, t, key): any lane */
MPLB_HDN void succ_row(const Cfg &c, const double *st, double t, const int *key, int u, Row *row) {
  Prim pr;
  prim_build(c, st, c.U + 3 * u, pr);
  prim_eval(c, pr, c.dt, row->st);
  make_key(c, row->st, row->key);
  row->verdict = 0;
  row->n_samples = 0;
  row->cost = 0;
  if (key_eq(row->key, key, c.nkey)) return;          /* em:158 tn == curr */
  if (!validate_primitive(c, pr)) return;            /* em:159 */
  row->t = fA(t, c.dt);                              /* em:161 */
  row->verdict = 1;
  bool same = true;
  for (int k = 0; k < c.dim; k++) same = same && (st[k] == row->st[k]);
  double cost = same ? 0 : traverse(c, pr, &row->n_samples); /* em:163 */
  if (!fIsInf(cost)) cost = fA(cost, fA(prim_J(c, pr), fM(c.w, c.dt))); /* em:164-165, eb:343-345 */
  row->cost = cost;
  row->hash = key_hash(row->key, c.nkey);
}

/* ------------------------------------------------------------------ node table (hm_ lookup by lattice key) */
MPLB_HD int table_find(const Ctx &x, const int *key, unsigned long long hash) {
  const int mask = x.h->tsize - 1;
  int s = (int)(hash & (unsigned long long)mask);
  while (true) {
    const int id = x.table[s];
    if (id < 0) return -1;
    if (key_eq(x.nodes[id].key, key, x.cfg.nkey)) return id;
    s = (s + 1) & mask;
  }
}
MPLB_HD void table_insert(const Ctx &x, int id) {
  const int mask = x.h->tsize - 1;
  int s = (int)(key_hash(x.nodes[id].key, x.cfg.nkey) & (unsigned long long)mask);
  while (x.table[s] >= 0) s = (s + 1) & mask;
  x.table[s] = id;
}
/* make_shared<State>(coord): a record (new or recycled in place) with the fresh State's defaults */
MPLB_HD void node_init(Node &n, const double *st, double t, const int *key) {
  for (int k = 0; k < 13; k++) n.st[k] = st[k];
  n.t = t;
  for (int k = 0; k < 12; k++) n.key[k] = key[k];
  n.g = LPA_INF; n.rhs = LPA_INF; n.h = LPA_INF;
  n.heap_pos = -1; n.n_succ = 0; n.pred_head = n.pred_tail = -1; n.n_pred = 0;
  n.opened = 0; n.closed = 0; n.in_hm = 0; n.pad = 0;
}
/* hm_[coord] for a successor / start coordinate: existing member, or a fresh State inserted at the end of the iteration order */
MPLB_HD int hm_get_or_create(const Ctx &x, const double *st, double t, const int *key, unsigned long long hash, const int *goal_key, bool *created) {
  Hdr &h = *x.h;
  int id = table_find(x, key, hash);
  *created = false;
  if (id >= 0 && x.nodes[id].in_hm) return id;
  if (id < 0) {
    id = h.n_nodes++;
    node_init(x.nodes[id], st, t, key);
    table_insert(x, id);
  } else node_init(x.nodes[id], st, t, key); /* dropped by an earlier getSubStateSpace: the reference builds a new State */
  x.nodes[id].in_hm = 1;
  x.order[h.n_order++] = id;
  (void)goal_key;
  *created = true;
  return id;
}

/* ------------------------------------------------------------------ pq_: d_ary_heap<arity 2, mutable> with compare_pair (ss:15-34) */
MPLB_HD bool heap_worse(const Ctx &x, const double *hf, const int *hn, int a, int b) { /* cmp(a, b): a has lower priority */
  if (hf[a] == hf[b]) {
    const Node &na = x.nodes[hn[a]], &nb = x.nodes[hn[b]];
    return fMin(na.g, na.rhs) > fMin(nb.g, nb.rhs);
  }
  return hf[a] > hf[b];
}
MPLB_HD void heap_swap(const Ctx &x, double *hf, int *hn, int a, int b, bool track) {
  const double f = hf[a]; hf[a] = hf[b]; hf[b] = f;
  const int n = hn[a]; hn[a] = hn[b]; hn[b] = n;
  if (track) { x.nodes[hn[a]].heap_pos = a; x.nodes[hn[b]].heap_pos = b; }
}
MPLB_HD void heap_sift_up(const Ctx &x, double *hf, int *hn, int pos, bool force, bool track) {
  while (pos != 0) {
    const int parent = (pos - 1) / 2;
    if (force || heap_worse(x, hf, hn, parent, pos)) { heap_swap(x, hf, hn, parent, pos, track); pos = parent; }
    else return;
  }
}
MPLB_HD void heap_sift_down(const Ctx &x, double *hf, int *hn, int n, int pos, bool track) {
  while (2 * pos + 1 < n) {
    int c = 2 * pos + 1;
    if (c + 1 < n && heap_worse(x, hf, hn, c, c + 1)) c = c + 1; /* std::max_element: the first of equally good children */
    if (!heap_worse(x, hf, hn, c, pos)) { heap_swap(x, hf, hn, pos, c, track); pos = c; }
    else return;
  }
}
MPLB_HD void pq_push(const Ctx &x, double f, int node) {
  Hdr &h = *x.h;
  const int pos = h.n_heap++;
  x.heap_f[pos] = f; x.heap_node[pos] = node;
  x.nodes[node].heap_pos = pos;
  heap_sift_up(x, x.heap_f, x.heap_node, pos, false, true);
}
MPLB_HD void pq_pop(const Ctx &x) {
  Hdr &h = *x.h;
  const int last = --h.n_heap;
  x.nodes[x.heap_node[0]].heap_pos = -1;
  if (last > 0) {
    x.heap_f[0] = x.heap_f[last]; x.heap_node[0] = x.heap_node[last];
    x.nodes[x.heap_node[0]].heap_pos = 0;
    heap_sift_down(x, x.heap_f, x.heap_node, last, 0, true);
  }
}
MPLB_HD void pq_erase(const Ctx &x, int node) { /* Boost: sift up unconditionally to the root, then pop */
  heap_sift_up(x, x.heap_f, x.heap_node, x.nodes[node].heap_pos, true, true);
  pq_pop(x);
}
MPLB_HD double calc_key(const Ctx &x, int id) { /* ss:270-272 */
  const Node &n = x.nodes[id];
  return fA(fMin(n.g, n.rhs), fM(x.h->eps, n.h));
}
MPLB_HD void update_node(const Ctx &x, int id) { /* ss:242-267 */
  Node &n = x.nodes[id];
  if (n.rhs != x.h->start_rhs) {
    n.rhs = LPA_INF;
    for (int p = n.pred_head; p >= 0; p = x.preds[p].next) {
      const double v = fA(x.nodes[x.preds[p].node].g, x.preds[p].cost);
      if (n.rhs > v) n.rhs = v;
    }
  }
  if (n.opened && !n.closed) { pq_erase(x, id); n.closed = 1; }
  if (n.g != n.rhs) {
    pq_push(x, calc_key(x, id), id);
    n.opened = 1;
    n.closed = 0;
  }
}
MPLB_HD int pred_find(const Ctx &x, int node, int pred_node) { /* index of pred_node in node's list, or -1 */
  int i = 0;
  for (int p = x.nodes[node].pred_head; p >= 0; p = x.preds[p].next, i++)
    if (x.preds[p].node == pred_node) return i;
  return -1;
}
MPLB_HD void pred_append(const Ctx &x, int node, int pred_node, double cost, int act) {
  Hdr &h = *x.h;
  const int r = h.n_pred++;
  x.preds[r].cost = cost; x.preds[r].node = pred_node; x.preds[r].act = act; x.preds[r].next = -1; x.preds[r].pad = 0;
  Node &n = x.nodes[node];
  if (n.pred_tail >= 0) x.preds[n.pred_tail].next = r; else n.pred_head = r;
  n.pred_tail = r;
  n.n_pred++;
}
MPLB_HD int pred_at(const Ctx &x, int node, int idx) {
  int p = x.nodes[node].pred_head;
  for (int i = 0; i < idx && p >= 0; i++) p = x.preds[p].next;
  return p;
}

/* ------------------------------------------------------------------ LPAstar (gs:194-365), split around the parallel get_succ */
MPLB_HD void goal_values(const Ctx &x, double *g, double *rhs, double *key) {
  const Hdr &h = *x.h;
  if (h.goal_node < 0) { *g = LPA_INF; *rhs = LPA_INF; *key = fA(LPA_INF, fM(h.eps, 0.0)); }
  else { *g = x.nodes[h.goal_node].g; *rhs = x.nodes[h.goal_node].rhs; *key = calc_key(x, h.goal_node); }
}
/* Entry of a plan (lane 0): pre-checks, start node, goal node.  Returns a final status, or -1 to enter the loop. */
MPLB_HDN int plan_begin(const Ctx &x, const double *start_st, double start_t, const double *goal_st) {
  Hdr &h = *x.h;
  const Cfg &c = x.cfg;
  h.n_prims = 0; h.n_valid = 0; h.n_samples = 0; h.n_explored = 0; h.pop_hash = 0xCBF29CE484222325ull; h.expand_iteration = 0;
  h.has_rows = 0; h.curr = -1;
  int pn[3];
  float_to_int(c, start_st, pn);
  if (!cell_free(c, pn)) return LPA_START_NOT_FREE; /* pb:283-287 */
  if (!h.initialized) { /* pb:296-304: a new StateSpace(epsilon_) only at the first plan */
    h.initialized = 1; h.eps = c.eps; h.start_g = 0; h.start_rhs = 0; h.start_t = 0; h.n_best = 0;
  }
  for (int k = 0; k < 13; k++) h.goal[k] = goal_st[k]; /* pb:306 */
  if (is_goal(c, h.goal, start_st)) return LPA_START_IS_GOAL; /* gs:200-205 */
  int key[12], gkey[12];
  make_key(c, start_st, key);
  make_key(c, h.goal, gkey);
  for (int k = 0; k < 12; k++) h.start_key[k] = key[k];
  bool created;
  const int s = hm_get_or_create(x, start_st, start_t, key, key_hash(key, c.nkey), gkey, &created); /* gs:208 */
  if (created) { /* gs:209-221 */
    Node &n = x.nodes[s];
    n.g = LPA_INF; n.rhs = 0;
    n.h = h.eps == 0 ? 0 : heur(c, h.goal, gkey, start_st, key);
    pq_push(x, calc_key(x, s), s);
    n.opened = 1; n.closed = 0;
  }
  h.start_node = s;
  if (h.n_best > 0 && is_goal(c, h.goal, x.nodes[x.best[h.n_best - 1]].st)) h.goal_node = x.best[h.n_best - 1]; /* gs:224-232 */
  else h.goal_node = -1;
  return -1;
}
/* Head of one iteration (lane 0).  Returns -1 when the popped node needs its successors generated (the lanes then fill
 * rows[] and pop_finish follows), -2 when the node's stored successor list is used (pop_finish follows directly), or a
 * final / internal status when the loop ends here. */
MPLB_HDN int pop_begin(const Ctx &x) {
  Hdr &h = *x.h;
  const Cfg &c = x.cfg;
  if (h.n_heap == 0) return LPA_QUEUE_EMPTY; /* the reference reads pq_.top() of an empty heap here (undefined) */
  double gg, grhs, gkey;
  goal_values(x, &gg, &grhs, &gkey);
  if (!(x.heap_f[0] < gkey || grhs != gg)) return LPA_OK; /* gs:244-245 */
  if (h.n_nodes + c.nU > h.cap_nodes || h.n_pred + c.nU > h.cap_pred || h.n_order + c.nU > h.cap_nodes) return LPA_NEED_GROW;
  h.expand_iteration++;
  const int curr = x.heap_node[0];
  pq_pop(x);
  Node &n = x.nodes[curr];
  n.closed = 1;
  if (n.g > n.rhs) n.g = n.rhs; /* gs:252-257 */
  else { n.g = LPA_INF; update_node(x, curr); }
  h.curr = curr;
  if (n.n_succ == 0) { /* gs:265-271: get_succ */
    h.has_rows = 1;
    h.n_explored++;
    h.pop_hash = (h.pop_hash ^ key_hash(n.key, c.nkey)) * 0x100000001B3ull;
    h.n_prims += c.nU;
    return -1;
  }
  h.has_rows = 0;
  return -2;
}
/* Tail of one iteration (lane 0): gs:287-336.  Returns -1 to continue, else the final status. */
MPLB_HDN int pop_finish(const Ctx &x) {
  Hdr &h = *x.h;
  const Cfg &c = x.cfg;
  const int curr = h.curr;
  int gkey[12];
  make_key(c, h.goal, gkey);
  Succ *sl = x.succ + (size_t)curr * c.nU;
  if (h.has_rows) { /* first expansion: the emitted rows, in control order, become the stored list */
    int ns = 0;
    for (int u = 0; u < c.nU; u++) {
      const Row &r = x.rows[u];
      h.n_samples += r.n_samples;
      if (!r.verdict) continue;
      if (!fIsInf(r.cost)) h.n_valid++;
      bool created;
      const int sid = hm_get_or_create(x, r.st, r.t, r.key, r.hash, gkey, &created);
      if (created) x.nodes[sid].h = h.eps == 0 ? 0 : heur(c, h.goal, gkey, r.st, r.key); /* gs:279-281 */
      sl[ns].node = sid; sl[ns].act = u; sl[ns].cost = r.cost;
      ns++;
      if (pred_find(x, sid, curr) < 0) pred_append(x, sid, curr, r.cost, u); /* gs:296-309 */
      update_node(x, sid);
    }
    x.nodes[curr].n_succ = ns;
  } else {
    const int ns = x.nodes[curr].n_succ;
    for (int s = 0; s < ns; s++) {
      int sid = sl[s].node;
      if (!x.nodes[sid].in_hm) { /* dropped since: hm_[succ_coord[s]] builds a new State from the stored coordinate */
        Prim pr;
        double st[13];
        int key[12];
        prim_build(c, x.nodes[curr].st, c.U + 3 * sl[s].act, pr);
        prim_eval(c, pr, c.dt, st);
        make_key(c, st, key);
        bool created;
        sid = hm_get_or_create(x, st, fA(x.nodes[curr].t, c.dt), key, key_hash(key, c.nkey), gkey, &created);
        x.nodes[sid].h = h.eps == 0 ? 0 : heur(c, h.goal, gkey, st, key);
        sl[s].node = sid;
      }
      if (pred_find(x, sid, curr) < 0) pred_append(x, sid, curr, sl[s].cost, sl[s].act);
      update_node(x, sid);
    }
  }
  if (is_goal(c, h.goal, x.nodes[curr].st)) h.goal_node = curr;                 /* gs:319 */
  if (c.max_num > 0 && h.expand_iteration >= c.max_num) return LPA_MAX_EXPAND;  /* gs:322-328 */
  if (h.n_heap == 0) return LPA_QUEUE_EMPTY;                                     /* gs:331-336 */
  return -1;
}
/* ------------------------------------------------------------------ the same tail, spread over the warp
 * pop_finish walks the successors one by one on lane 0.  Most of what it does per successor is independent of the other
 * successors of the same pop: the key-table probe, building a new State, the scan of the successor's predecessor list for the
 * popped node, the append, and the recomputation of rhs over that list (it reads g of predecessors, and g changes only at the head
 * of a pop).  Only three things are order-defining: the ids / iteration-order slots / pool records handed out to new objects,
 * and the priority-queue operations.  pop_finish_warp therefore runs chunks of 32 successors: the lanes probe and detect
 * siblings with equal keys (the later one must see the earlier one's node, exactly as `hm_[coord]` would), lane 0 hands out ids
 * and slots in control order, the lanes build nodes / append predecessor records / recompute rhs for the first occurrence of
 * each node, and lane 0 then applies rhs and the heap operations in control order (a repeated sibling takes the plain serial
 * updateNode).  Statement for statement this leaves every array as pop_finish leaves it; the host build (tests/cpp/lpa_emul.cpp)
 * runs the same source with the lane loops serialised and is compared with the checker.
 * Written once for both builds: LPA_LANES(l) is the lane loop (one iteration per thread on the device), LPA_SYNC a warp barrier. */
#ifdef __CUDA_ARCH__
#define LPA_LANES(l) for (int l = (int)threadIdx.x, l##_go = 1; l##_go; l##_go = 0)
#define LPA_SYNC() __syncwarp()
#define LPA_LANE0 (threadIdx.x == 0)
#else
#ifdef LPA_REVERSE_LANES /* test builds: the result must not depend on the order in which lanes run inside a phase */
#define LPA_LANES(l) for (int l = 31; l >= 0; l--)
#else
#define LPA_LANES(l) for (int l = 0; l < 32; l++)
#endif
#define LPA_SYNC() ((void)0)
#define LPA_LANE0 true
#endif

struct PopScratch { /* per-warp staging (shared memory on the device) */
  int sid[32], state[32], leader[32], item[32], opos[32], prec[32];
  unsigned char active[32], need_append[32], has_rhs[32];
  double new_rhs[32];
  int ret;
};

MPLB_HD void table_insert_shared(const Ctx &x, int id) { /* several lanes insert different ids at once */
#ifdef __CUDA_ARCH__
  const int mask = x.h->tsize - 1;
  int s = (int)(key_hash(x.nodes[id].key, x.cfg.nkey) & (unsigned long long)mask);
  while (atomicCAS(&x.table[s], -1, id) != -1) s = (s + 1) & mask;
#else
  table_insert(x, id);
#endif
}
MPLB_HD void update_node_heap(const Ctx &x, int id) { /* the queue half of updateNode (ss:256-266) */
  Node &n = x.nodes[id];
  if (n.opened && !n.closed) { pq_erase(x, id); n.closed = 1; }
  if (n.g != n.rhs) {
    pq_push(x, calc_key(x, id), id);
    n.opened = 1;
    n.closed = 0;
  }
}
/* Called by all lanes of the warp (device) / once (host).  Result in S->ret: -1 to continue, else the final status. */
MPLB_HDN void pop_finish_warp(const Ctx &x, PopScratch *S) {
  Hdr &h = *x.h;
  const Cfg &c = x.cfg;
  const int curr = h.curr;
  const bool fresh = h.has_rows != 0;
  Succ *sl = x.succ + (size_t)curr * c.nU;
  const int n_in = fresh ? c.nU : x.nodes[curr].n_succ; /* controls (fresh) or stored entries */
  int gkey[12];
  make_key(c, h.goal, gkey);
  int ns = 0; /* lane 0's running count of stored entries (fresh case) */
  for (int base = 0; base < n_in; base += 32) {
    /* ---- 1: probe.  state 0 = member of hm_, 1 = known key whose State was dropped, 2 = never seen */
    LPA_LANES(l) {
      const int i = base + l;
      S->active[l] = 0; S->need_append[l] = 0; S->has_rhs[l] = 0; S->leader[l] = l; S->sid[l] = -1; S->state[l] = 0;
      if (i < n_in) {
        if (fresh) {
          const Row &r = x.rows[i];
          if (r.verdict) {
            S->active[l] = 1;
            const int id = table_find(x, r.key, r.hash);
            S->sid[l] = id;
            S->state[l] = id < 0 ? 2 : (x.nodes[id].in_hm ? 0 : 1);
          }
        } else {
          S->active[l] = 1;
          const int id = sl[i].node;
          S->sid[l] = id;
          S->state[l] = x.nodes[id].in_hm ? 0 : 1;
          if (S->state[l] == 1) { /* hm_[succ_coord[s]] will build a new State from the stored coordinate: recompute it */
            Row &r = x.rows[l];
            Prim pr;
            prim_build(c, x.nodes[curr].st, c.U + 3 * sl[i].act, pr);
            prim_eval(c, pr, c.dt, r.st);
            make_key(c, r.st, r.key);
            r.t = fA(x.nodes[curr].t, c.dt);
            r.hash = key_hash(r.key, c.nkey);
          }
        }
      }
    }
    LPA_SYNC();
    /* ---- 2: the first sibling of this chunk with the same key (ids for stored entries) */
    LPA_LANES(l) {
      if (S->active[l]) {
        for (int k = 0; k < l; k++) {
          if (!S->active[k]) continue;
          bool same;
          if (fresh) same = x.rows[base + k].hash == x.rows[base + l].hash && key_eq(x.rows[base + k].key, x.rows[base + l].key, c.nkey);
          else same = S->sid[k] == S->sid[l];
          if (same) { S->leader[l] = k; break; }
        }
      }
    }
    LPA_SYNC();
    /* ---- 3: lane 0 hands out node ids, iteration-order slots and list positions in control order */
    if (LPA_LANE0) {
      for (int l = 0; l < 32; l++) {
        const int i = base + l;
        if (fresh && i < n_in) h.n_samples += x.rows[i].n_samples;
        if (!S->active[l]) continue;
        if (fresh) { if (!fIsInf(x.rows[i].cost)) h.n_valid++; S->item[l] = ns++; }
        if (S->leader[l] != l) continue;
        if (S->state[l] == 2) S->sid[l] = h.n_nodes++;
        if (S->state[l] != 0) S->opos[l] = h.n_order++;
      }
    }
    LPA_SYNC();
    /* ---- 4: the lanes build the new States */
    LPA_LANES(l) {
      if (S->active[l] && S->leader[l] == l && S->state[l] != 0) {
        const Row &r = fresh ? x.rows[base + l] : x.rows[l];
        Node &n = x.nodes[S->sid[l]];
        node_init(n, r.st, r.t, r.key);
        n.in_hm = 1;
        n.h = h.eps == 0 ? 0 : heur(c, h.goal, gkey, r.st, r.key); /* gs:279-281 */
        x.order[S->opos[l]] = S->sid[l];
        if (S->state[l] == 2) table_insert_shared(x, S->sid[l]);
      }
    }
    LPA_SYNC();
    /* ---- 5: stored list entry, and does the successor already list the popped node as a predecessor? (gs:296-303) */
    LPA_LANES(l) {
      if (S->active[l]) {
        if (S->leader[l] != l) S->sid[l] = S->sid[S->leader[l]];
        const int i = base + l;
        if (fresh) { Succ &e = sl[S->item[l]]; e.node = S->sid[l]; e.act = i; e.cost = x.rows[i].cost; }
        if (S->leader[l] == l) S->need_append[l] = pred_find(x, S->sid[l], curr) < 0;
      }
    }
    LPA_SYNC();
    if (LPA_LANE0)
      for (int l = 0; l < 32; l++)
        if (S->active[l] && S->need_append[l]) S->prec[l] = h.n_pred++;
    LPA_SYNC();
    /* ---- 6: append (one lane per node, so the lists never collide), then rhs over the node's own list (ss:245-253) */
    LPA_LANES(l) {
      if (S->active[l] && S->leader[l] == l) {
        const int sid = S->sid[l];
        Node &n = x.nodes[sid];
        if (S->need_append[l]) {
          const int i = base + l;
          const int r = S->prec[l];
          x.preds[r].cost = fresh ? x.rows[i].cost : sl[i].cost;
          x.preds[r].node = curr;
          x.preds[r].act = fresh ? i : sl[i].act;
          x.preds[r].next = -1; x.preds[r].pad = 0;
          if (n.pred_tail >= 0) x.preds[n.pred_tail].next = r; else n.pred_head = r;
          n.pred_tail = r;
          n.n_pred++;
        }
        if (n.rhs != h.start_rhs) {
          double v = LPA_INF;
          for (int p = n.pred_head; p >= 0; p = x.preds[p].next) {
            const double w = fA(x.nodes[x.preds[p].node].g, x.preds[p].cost);
            if (v > w) v = w;
          }
          S->new_rhs[l] = v;
          S->has_rhs[l] = 1;
        }
      }
    }
    LPA_SYNC();
    /* ---- 7: lane 0 commits rhs and runs the queue operations in control order */
    if (LPA_LANE0) {
      for (int l = 0; l < 32; l++) {
        if (!S->active[l]) continue;
        const int sid = S->sid[l];
        if (S->leader[l] == l) {
          if (S->has_rhs[l]) x.nodes[sid].rhs = S->new_rhs[l];
          update_node_heap(x, sid);
        } else {
          update_node(x, sid); /* a repeated sibling: the plain serial statement */
        }
      }
    }
    LPA_SYNC();
  }
  if (LPA_LANE0) {
    if (fresh) x.nodes[curr].n_succ = ns;
    int ret = -1;
    if (is_goal(c, h.goal, x.nodes[curr].st)) h.goal_node = curr;            /* gs:319 */
    if (c.max_num > 0 && h.expand_iteration >= c.max_num) ret = LPA_MAX_EXPAND; /* gs:322-328 */
    else if (h.n_heap == 0) ret = LPA_QUEUE_EMPTY;                              /* gs:331-336 */
    S->ret = ret;
  }
  LPA_SYNC();
}

/* recoverTraj (gs:369-455); returns LPA_OK or LPA_TRACEBACK_FAILED; cost = goal g - start_g_ (gs:362) */
MPLB_HDN int recover(const Ctx &x, int *n_seg, double *cost) {
  Hdr &h = *x.h;
  const Cfg &c = x.cfg;
  h.n_best = 0;
  *n_seg = 0;
  *cost = LPA_INF;
  if (h.goal_node < 0) return LPA_TRACEBACK_FAILED; /* the detached goal State has no predecessors */
  int cur = h.goal_node, na = 0;
  bool found = false, cycle = false;
  while (x.nodes[cur].pred_head >= 0) {
    if (h.n_best > h.n_order) { cycle = true; break; } /* more steps than hm_ has members: a cycle of best predecessors.  The
                                                          reference's loop (gs:377-438) would never return; the trace-back fails
                                                          and best_child_ is left empty */
    x.best[h.n_best++] = cur;
    int min_p = -1;
    double min_rhs = LPA_INF, min_g = LPA_INF;
    for (int p = x.nodes[cur].pred_head; p >= 0; p = x.preds[p].next) {
      const double pg = x.nodes[x.preds[p].node].g;
      const double v = fA(pg, x.preds[p].cost);
      if (min_rhs > v) { min_rhs = v; min_g = pg; min_p = p; }
      else if (!fIsInf(x.preds[p].cost) && min_rhs == v) {
        if (min_g < pg) { min_g = pg; min_p = p; }
      }
    }
    if (min_p < 0) break;
    x.traj_act[na++] = x.preds[min_p].act;
    cur = x.preds[min_p].node;
    if (key_eq(x.nodes[cur].key, h.start_key, c.nkey)) { x.best[h.n_best++] = cur; found = true; break; }
  }
  if (cycle) { h.n_best = 0; h.fault |= 2; return LPA_TRACEBACK_FAILED; }
  for (int i = 0; i < h.n_best / 2; i++) { const int t = x.best[i]; x.best[i] = x.best[h.n_best - 1 - i]; x.best[h.n_best - 1 - i] = t; }
  if (!found) return LPA_TRACEBACK_FAILED;
  for (int i = 0; i < na / 2; i++) { const int t = x.traj_act[i]; x.traj_act[i] = x.traj_act[na - 1 - i]; x.traj_act[na - 1 - i] = t; }
  *n_seg = na;
  *cost = fS(x.nodes[h.goal_node].g, h.start_g);
  return LPA_OK;
}

/* ------------------------------------------------------------------ getSubStateSpace (ss:116-204), lane 0 */
MPLB_HD void epq_push(const Ctx &x, double f, int node) {
  Hdr &h = *x.h;
  const int pos = h.n_epq++;
  x.epq_f[pos] = f; x.epq_node[pos] = node;
  heap_sift_up(x, x.epq_f, x.epq_node, pos, false, false);
}
MPLB_HD int epq_pop(const Ctx &x) {
  Hdr &h = *x.h;
  const int top = x.epq_node[0];
  const int last = --h.n_epq;
  if (last > 0) {
    x.epq_f[0] = x.epq_f[last]; x.epq_node[0] = x.epq_node[last];
    heap_sift_down(x, x.epq_f, x.epq_node, last, 0, false);
  }
  return top;
}
MPLB_HDN int sub_state_space(const Ctx &x, int time_step) {
  Hdr &h = *x.h;
  const Cfg &c = x.cfg;
  if (h.n_best == 0 || time_step < 0 || time_step >= h.n_best) return LPA_OK;
  int curr = x.best[time_step];
  h.start_g = x.nodes[curr].g; h.start_rhs = x.nodes[curr].rhs; h.start_t = x.nodes[curr].t;
  for (int i = 0; i < h.n_order; i++) { /* ss:126-136: every member of hm_ (the root included) */
    Node &n = x.nodes[x.order[i]];
    n.g = LPA_INF; n.rhs = LPA_INF;
    n.pred_head = n.pred_tail = -1; n.n_pred = 0;
    x.mark[x.order[i]] = 0;
  }
  h.n_pred = 0; /* every live list is empty now: the record pool restarts */
  x.nodes[curr].g = h.start_g; x.nodes[curr].rhs = h.start_rhs;
  int n_new = 0;
  h.n_epq = 0;
  epq_push(x, x.nodes[curr].rhs, curr);
  x.mark[curr] = 1; x.order2[n_new++] = curr;
  int fault = 0;
  while (h.n_epq > 0) {
    curr = epq_pop(x);
    const Succ *sl = x.succ + (size_t)curr * c.nU;
    const int ns = x.nodes[curr].n_succ;
    for (int i = 0; i < ns; i++) {
      const int sid = sl[i].node;
      if (!x.nodes[sid].in_hm) { fault = 1; continue; } /* "critical bug!!!!" (ss:160-163): the reference dereferences a null State */
      if (!x.mark[sid]) { x.mark[sid] = 1; x.order2[n_new++] = sid; }
      if (pred_find(x, sid, curr) < 0) pred_append(x, sid, curr, sl[i].cost, sl[i].act);
      const double tentative = fA(x.nodes[curr].rhs, sl[i].cost);
      if (tentative < x.nodes[sid].rhs) {
        x.nodes[sid].rhs = tentative;
        if (x.nodes[sid].closed) {
          x.nodes[sid].g = x.nodes[sid].rhs;
          epq_push(x, x.nodes[sid].rhs, sid);
        }
      }
    }
  }
  for (int i = 0; i < h.n_order; i++) { /* hm_ = new_hm */
    const int id = x.order[i];
    if (!x.mark[id]) { x.nodes[id].in_hm = 0; x.nodes[id].heap_pos = -1; }
  }
  for (int i = 0; i < n_new; i++) x.order[i] = x.order2[i];
  h.n_order = n_new;
  h.n_heap = 0; /* pq_.clear(), then the open members in iteration order (ss:190-199) */
  for (int i = 0; i < n_new; i++) x.nodes[x.order[i]].heap_pos = -1;
  for (int i = 0; i < n_new; i++) {
    const int id = x.order[i];
    if (x.nodes[id].opened && !x.nodes[id].closed) pq_push(x, calc_key(x, id), id);
  }
  if (fault) h.fault = 1;
  return fault ? LPA_FAULT : LPA_OK;
}

/* ------------------------------------------------------------------ getLinkedNodes (map_planner.cpp:125-158)
 * One call per position of the iteration order (any thread): the links of that node's predecessor edges, in (pred index,
 * sample) order.  out == nullptr counts, otherwise fills starting at out[0]; returns the count. */
MPLB_HDN int link_node(const Ctx &x, int order_pos, Link *out) {
  const Cfg &c = x.cfg;
  const int nid = x.order[order_pos];
  int cnt = 0, i = 0;
  for (int p = x.nodes[nid].pred_head; p >= 0; p = x.preds[p].next, i++) {
    Prim pr;
    prim_build(c, x.nodes[x.preds[p].node].st, c.U + 3 * x.preds[p].act, pr);
    const double max_v = prim_max_v(c, pr); /* std::max over the axes */
    const int n = (int)(1.0 * ceil(fD(fM(max_v, c.dt), c.res)));
    const double dts = fD(c.dt, (double)n);
    int prev_id = -1;
    for (int s = 0; s <= n; s++) {
      double st[13];
      int pn[3];
      prim_eval(c, pr, fM((double)s, dts), st);
      float_to_int(c, st, pn);
      const int id = cell_index(c, pn);
      if (id != prev_id) {
        if (out) { Link &l = out[cnt]; l.vox = id; l.node = nid; l.pred_idx = i; l.cell[0] = pn[0]; l.cell[1] = pn[1]; l.cell[2] = pn[2]; }
        cnt++;
        prev_id = id;
      }
    }
  }
  return cnt;
}

/* ------------------------------------------------------------------ increaseCost / decreaseCost (ss:207-240) for one affected
 * (node, pred index) pair, lane 0, in the order updateBlockedNodes / updateClearedNodes produce (map_planner.cpp:160-185) */
MPLB_HDN void apply_change(const Ctx &x, int node, int pred_idx, bool blocked) {
  const Cfg &c = x.cfg;
  const int p = pred_at(x, node, pred_idx);
  if (p < 0) return;
  Pred &pr_ = x.preds[p];
  double new_cost;
  if (blocked) {
    if (fIsInf(pr_.cost)) return;
    new_cost = LPA_INF;
  } else {
    if (!fIsInf(pr_.cost)) return;
    Prim pr;
    prim_build(c, x.nodes[pr_.node].st, c.U + 3 * pr_.act, pr);
    if (!prim_is_free(c, pr)) return;
    new_cost = fA(prim_J(c, pr), fM(c.w, c.dt)); /* eb:343-345 */
  }
  pr_.cost = new_cost;
  update_node(x, node);
  Succ *sl = x.succ + (size_t)pr_.node * c.nU;
  const int ns = x.nodes[pr_.node].n_succ;
  for (int j = 0; j < ns; j++)
    if (sl[j].act == pr_.act) { sl[j].cost = new_cost; break; }
}

}  // namespace mplb_lpa
#endif
