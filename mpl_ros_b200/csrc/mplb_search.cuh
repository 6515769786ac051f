/*
 * mplb_search.cuh — the fused A* kernel: one CTA owns one plan; a persistent grid drains a batch.
 *
 * Reference path restated per pop (paths under motion_primitive_library/include/):
 *   GraphSearch::Astar            mpl_planner/common/graph_search.h:39-182   (pop, relax, terminate)
 *   env_map::get_succ             mpl_planner/env/env_map.h:147-172          (phase B1: one thread per control u)
 *   env_map::traverse_primitive   mpl_planner/env/env_map.h:90-132           (phase B2: one lane per sample)
 *   env_map::is_goal + rayTrace   env_map.h:25-45, mpl_collision/map_util.h:117-134
 *   priorityQueue / compare_pair  mpl_planner/common/state_space.h:15-34     (binary heap, same sift rules as
 *                                 boost::heap::d_ary_heap<arity<2>, mutable_<true>> so pop order is identical)
 *   recoverTraj                   graph_search.h:369-455 (best-predecessor rule kept as a running argmin)
 *
 * Search state of a plan lives in a private HBM arena (nodes, heap, open-addressing table); the only
 * data shared between CTAs is the read-only map (bit-bricks) and the control / sample-time tables.
 */
#pragma once
#include "mplb_device.cuh"
#include "../../include/mplb.h"

namespace mplb {

#define MPLB_INTERNAL_OVERFLOW 100 /* arena too small: host retries the plan in a larger tier */
#define MPLB_INTERNAL_BADCTRL 9

struct __align__(16) NodeHot {
  unsigned long long k0, k1; /* packed lattice key */
  unsigned long long kh;     /* 64-bit hash of the lattice ints */
  double g, h;
  double pg;     /* g of the best predecessor (tie rule of recoverTraj, gs:391-405) */
  int parent;    /* best predecessor node */
  int heap_pos;  /* position of the live heap entry while open */
  short action;  /* action id of parent -> this */
  unsigned char flags; /* 1 = iterationopened, 2 = iterationclosed */
  unsigned char pad0;
  int pad1;
};
static_assert(sizeof(NodeHot) == 64, "NodeHot must be 64 bytes");

struct HeapEnt {
  double f; /* heap key (gs:54,119) */
  double g; /* copy of node g for compare_pair's tie-break (ss:19-24) */
  int node;
  int pad;
};
static_assert(sizeof(HeapEnt) == 24, "HeapEnt must be 24 bytes");

/* cmp(a,b) of compare_pair: true iff a is worse (lower priority) than b. */
__device__ __forceinline__ bool heap_worse(const HeapEnt &a, const HeapEnt &b) {
  return (a.f == b.f) ? (a.g > b.g) : (a.f > b.f);
}

struct BatchArgs {
  const mplb_waypoint *starts, *goals;
  mplb_result *results;
  int *actions;       /* [n * max_seg] or null */
  double *seg_states; /* [n * max_seg * 13] or null */
  int max_seg;
  const int *work;    /* plan ids of this tier (null = identity) */
  int n_work;
  int *work_counter;
  unsigned char *arena; /* slot s at arena + s*stride */
  size_t stride;
  int cap;       /* nodes (and heap entries, pop-log entries) per slot */
  int tsize_max; /* table slots per slot arena (power of two) */
  size_t off_state, off_heap, off_table, off_poplog;
  int want_poplog;
  int *slot_of_plan;   /* optional: which slot ran plan i (retained single plan) */
  int *overflow_count; /* plans whose arena overflowed in this tier ... */
  int *overflow_list;  /* ... and their ids, for the next (larger) tier */
};

template <int DIM, int ORD>
struct PlanSmem {
  static constexpr int NS = DIM * ORD;
  double cur[NS];
  int cur_ints[NS];
  double cur_g;
  int cur_node;
  double goal_pos[3], goal_vel[3], goal_acc[3];
  unsigned long long gk0, gk1;
  int goal_key_ok;
  double U[MPLB_MAXU * 3];
  double es[MPLB_MAXU * NS]; /* end states, [u][d*DIM+ax] */
  double cost[MPLB_MAXU];
  unsigned long long k0[MPLB_MAXU], k1[MPLB_MAXU], kh[MPLB_MAXU];
  int verdict[MPLB_MAXU]; /* 0 self, 1 dyn, 2 blocked, 3 valid, 4 valid-no-motion, 5 needs sampling (transient) */
  int nsamp[MPLB_MAXU];   /* divisor n */
  int first[MPLB_MAXU];   /* first blocked sample index or INT_MAX */
  int n_nodes, n_heap, tsize, pops, n_closed, status, done, plan_idx, goal_hit, key_bad;
  long long n_samples, n_valid;
  unsigned long long pop_hash, closed_hash;
};

template <int DIM, int ORD>
struct Arena {
  NodeHot *hot;
  double *state;
  HeapEnt *heap;
  unsigned long long *table;
  int *poplog;
};

/* ---------------------------------------------------------------- heap (single-thread sift, hole method) */
__device__ __forceinline__ void heap_sift_up(HeapEnt *heap, NodeHot *hot, int pos, HeapEnt e) {
  while (pos != 0) {
    int par = (pos - 1) >> 1;
    HeapEnt pe = heap[par];
    if (heap_worse(pe, e)) {
      heap[pos] = pe;
      hot[pe.node].heap_pos = pos;
      pos = par;
    } else break;
  }
  heap[pos] = e;
  hot[e.node].heap_pos = pos;
}

__device__ __forceinline__ void heap_sift_down(HeapEnt *heap, NodeHot *hot, int n, int pos, HeapEnt e) {
  while (true) {
    int c = 2 * pos + 1;
    if (c >= n) break;
    HeapEnt ce = heap[c];
    if (c + 1 < n) {
      HeapEnt re = heap[c + 1];
      if (heap_worse(ce, re)) { c = c + 1; ce = re; } /* right child only if strictly better */
    }
    if (!heap_worse(ce, e)) { /* ties still move the element down (boost siftdown) */
      heap[pos] = ce;
      hot[ce.node].heap_pos = pos;
      pos = c;
    } else break;
  }
  heap[pos] = e;
  hot[e.node].heap_pos = pos;
}

/* ---------------------------------------------------------------- open-addressing table of (fingerprint, node+1) */
__device__ __forceinline__ int table_find(const unsigned long long *table, int tsize, const NodeHot *hot,
                                          unsigned long long kh, unsigned long long k0, unsigned long long k1) {
  unsigned mask = (unsigned)tsize - 1u;
  unsigned i = (unsigned)kh & mask;
  unsigned fp = (unsigned)(kh >> 32);
  while (true) {
    unsigned long long s = table[i];
    if (s == 0ull) return -1;
    if ((unsigned)(s >> 32) == fp) {
      int id = (int)(unsigned)s - 1;
      if (hot[id].k0 == k0 && hot[id].k1 == k1) return id;
    }
    i = (i + 1) & mask;
  }
}

__device__ __forceinline__ void table_insert_serial(unsigned long long *table, int tsize, unsigned long long kh, int id) {
  unsigned mask = (unsigned)tsize - 1u;
  unsigned i = (unsigned)kh & mask;
  while (table[i] != 0ull) i = (i + 1) & mask;
  table[i] = ((kh >> 32) << 32) | (unsigned long long)(unsigned)(id + 1);
}

__device__ __forceinline__ void table_insert_atomic(unsigned long long *table, int tsize, unsigned long long kh, int id) {
  unsigned mask = (unsigned)tsize - 1u;
  unsigned i = (unsigned)kh & mask;
  unsigned long long val = ((kh >> 32) << 32) | (unsigned long long)(unsigned)(id + 1);
  while (atomicCAS(&table[i], 0ull, val) != 0ull) i = (i + 1) & mask;
}

/* ---------------------------------------------------------------- phase B1: one thread per control (em:155-160,163-165) */
template <int DIM, int ORD>
__device__ __forceinline__ void expand_b1(const DevCfg &c, PlanSmem<DIM, ORD> &S, int i) {
  constexpr int NS = DIM * ORD;
  const double T = c.dt;
  double es[NS];
  double max_v = 0.0, J = 0.0;
  bool dyn_ok = true, same_pos = true;
#pragma unroll
  for (int ax = 0; ax < DIM; ax++) {
    Axis<ORD> A(&S.cur[ax], DIM, S.U[i * 3 + ax]);
    es[0 * DIM + ax] = A.p(T);
    if (ORD >= 2) es[1 * DIM + ax] = A.v(T);
    if (ORD >= 3) es[2 * DIM + ax] = A.a(T);
    if (ORD >= 4) es[3 * DIM + ax] = A.j(T);
    double mv = A.max_vel(T);
    if (mv > max_v) max_v = mv; /* em:91-94 */
    /* validate_primitive pr:449-475: only derivatives below the control order, bound <= 0 disables (pr:485) */
    if (ORD >= 2 && c.v_max > 0.0 && mv > c.v_max) dyn_ok = false;
    if (ORD >= 3 && c.a_max > 0.0 && A.max_acc(T) > c.a_max) dyn_ok = false;
    if (ORD >= 4 && c.j_max > 0.0 && A.max_jrk(T) > c.j_max) dyn_ok = false;
    J = dadd(J, A.J(T)); /* pr:403-407 */
    same_pos = same_pos && (S.cur[ax] == es[ax]); /* em:163 */
  }
  int ints[NS];
  lattice_ints<DIM, ORD>(es, ints);
  bool self = true;
#pragma unroll
  for (int f = 0; f < NS; f++) self = self && (ints[f] == S.cur_ints[f]);
#pragma unroll
  for (int f = 0; f < NS; f++) S.es[i * NS + f] = es[f];
  unsigned long long k0, k1, kh;
  bool key_ok = pack_key<DIM, ORD>(c, ints, k0, k1, kh);
  S.k0[i] = k0; S.k1[i] = k1; S.kh[i] = kh;
  S.first[i] = 0x7fffffff;
  S.nsamp[i] = 0;
  S.cost[i] = dadd(J, dmul(c.w, T)); /* eb:343-345; traverse contributes 0 on a plain map */
  int verdict;
  if (self) verdict = 0;
  else if (!dyn_ok) verdict = 1;
  else if (same_pos) verdict = 4;
  else {
    verdict = 5;
    int n = __double2int_rz(ceil(ddiv(dmul(max_v, T), c.res))); /* em:95 */
    S.nsamp[i] = n < 5 ? 5 : n;
  }
  if ((verdict >= 3) && !key_ok) S.key_bad = 1;
  S.verdict[i] = verdict;
}

/* One collision sample (em:100-104,119): returns true when sample k of control i is outside or occupied. */
template <int DIM, int ORD>
__device__ __forceinline__ bool sample_blocked(const DevCfg &c, const PlanSmem<DIM, ORD> &S, int i, double t, int *cell_idx) {
  int pn[3] = {0, 0, 0};
  bool outside = false;
#pragma unroll
  for (int ax = 0; ax < DIM; ax++) {
    Axis<ORD> A(&S.cur[ax], DIM, S.U[i * 3 + ax]);
    pn[ax] = float_to_cell(A.p(t), c.origin[ax], c.res);
    outside = outside || pn[ax] < 0 || pn[ax] >= c.nd[ax];
  }
  if (outside) { if (cell_idx) *cell_idx = -1; return true; }
  if (cell_idx) *cell_idx = (DIM == 2) ? pn[0] + c.nd[0] * pn[1] : pn[0] + c.nd[0] * pn[1] + c.nd[0] * c.nd[1] * pn[2];
  return brick_occupied<DIM>(c, pn[0], pn[1], pn[2]);
}

/* phase B2: warps take controls round-robin; lanes take samples (em:99 loop, order-free on a plain map). */
template <int DIM, int ORD>
__device__ __forceinline__ void expand_b2(const DevCfg &c, PlanSmem<DIM, ORD> &S, int warp, int lane, int nwarps) {
  for (int i = warp; i < c.nU; i += nwarps) {
    if (S.verdict[i] != 5) continue;
    int n = S.nsamp[i];
    int cnt = c.tcnt[n];
    const double *tt = c.ttab + c.toff[n];
    int first = 0x7fffffff;
    for (int base = 0; base < cnt; base += 32) {
      int k = base + lane;
      bool blocked = false;
      if (k < cnt) blocked = sample_blocked<DIM, ORD>(c, S, i, __ldg(&tt[k]), nullptr);
      unsigned m = __ballot_sync(0xffffffffu, blocked);
      if (m) { first = base + __ffs(m) - 1; break; }
    }
    if (lane == 0) {
      S.first[i] = first;
      S.verdict[i] = (first == 0x7fffffff) ? 3 : 2;
      S.nsamp[i] = n;
    }
  }
}

/* em:25-45 for the popped state (tolerances, then ray trace mu:117-134); executed by one warp. */
template <int DIM, int ORD>
__device__ __forceinline__ bool goal_test_warp(const DevCfg &c, const PlanSmem<DIM, ORD> &S, const double *st, int lane) {
  double m = 0.0;
#pragma unroll
  for (int ax = 0; ax < DIM; ax++) m = fmax(m, fabs(dsub(st[ax], S.goal_pos[ax])));
  bool goaled = m <= c.tol_pos;
  if (goaled && c.tol_vel >= 0.0) {
    m = 0.0;
#pragma unroll
    for (int ax = 0; ax < DIM; ax++) m = fmax(m, fabs(dsub((ORD >= 2) ? st[DIM + ax] : 0.0, S.goal_vel[ax])));
    goaled = m <= c.tol_vel;
  }
  if (goaled && c.tol_acc >= 0.0) {
    m = 0.0;
#pragma unroll
    for (int ax = 0; ax < DIM; ax++) m = fmax(m, fabs(dsub((ORD >= 3) ? st[2 * DIM + ax] : 0.0, S.goal_acc[ax])));
    goaled = m <= c.tol_acc;
  }
  if (!goaled) return false;
  /* rayTrace(state.pos, goal.pos) */
  double diff[3] = {0, 0, 0}, q = 0.0;
#pragma unroll
  for (int ax = 0; ax < DIM; ax++) {
    diff[ax] = dsub(S.goal_pos[ax], st[ax]);
    q = fmax(q, fabs(ddiv(diff[ax], c.res)));
  }
  int max_diff = __double2int_rz(ddiv(q, 0.8));
  double s = ddiv(1.0, (double)max_diff);
  bool hit = false;
  for (int base = 1; base < max_diff; base += 32) {
    int n = base + lane;
    bool outside = false, occ = false;
    if (n < max_diff) {
      int pn[3] = {0, 0, 0};
#pragma unroll
      for (int ax = 0; ax < DIM; ax++) {
        double pt = dadd(st[ax], dmul(dmul(diff[ax], s), (double)n));
        pn[ax] = float_to_cell(pt, c.origin[ax], c.res);
        outside = outside || pn[ax] < 0 || pn[ax] >= c.nd[ax];
      }
      if (!outside) occ = brick_occupied<DIM>(c, pn[0], pn[1], pn[2]);
    }
    unsigned mo = __ballot_sync(0xffffffffu, outside);
    unsigned mh = __ballot_sync(0xffffffffu, occ);
    if (mo) { /* break at the first outside cell: only hits before it count */
      unsigned before = (1u << (__ffs(mo) - 1)) - 1u;
      hit = (mh & before) != 0u;
      break;
    }
    if (mh) { hit = true; break; }
  }
  return !hit;
}

/* eb:46-64 (heur_ignore_dynamics_ = true, no prior trajectory) */
template <int DIM, int ORD>
__device__ __forceinline__ double heuristic(const DevCfg &c, const PlanSmem<DIM, ORD> &S, const double *st,
                                            unsigned long long k0, unsigned long long k1) {
  if (c.eps == 0.0) return 0.0; /* gs:53,87 */
  if (S.goal_key_ok && k0 == S.gk0 && k1 == S.gk1) return 0.0;
  double m = 0.0;
#pragma unroll
  for (int ax = 0; ax < DIM; ax++) m = fmax(m, fabs(dsub(st[ax], S.goal_pos[ax])));
  if (c.v_max > 0.0) return ddiv(dmul(c.w, m), c.v_max);
  return dmul(c.w, m);
}

/* ---------------------------------------------------------------- the kernel */
template <int DIM, int ORD>
__global__ void __launch_bounds__(MPLB_NT) astar_batch_kernel(const DevCfg c, const BatchArgs a) {
  constexpr int NS = DIM * ORD;
  constexpr int NW = MPLB_NT / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PlanSmem<DIM, ORD> &S = *reinterpret_cast<PlanSmem<DIM, ORD> *>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  unsigned char *base = a.arena + (size_t)blockIdx.x * a.stride;
  NodeHot *hot = reinterpret_cast<NodeHot *>(base);
  double *state = reinterpret_cast<double *>(base + a.off_state);
  HeapEnt *heap = reinterpret_cast<HeapEnt *>(base + a.off_heap);
  unsigned long long *table = reinterpret_cast<unsigned long long *>(base + a.off_table);
  int *poplog = reinterpret_cast<int *>(base + a.off_poplog);

  for (int i = tid; i < c.nU * 3; i += MPLB_NT) S.U[i] = c.U[i];

  while (true) {
    __syncthreads();
    if (tid == 0) S.plan_idx = atomicAdd(a.work_counter, 1);
    __syncthreads();
    const int w = S.plan_idx;
    if (w >= a.n_work) break;
    const int pid = a.work ? a.work[w] : w;
    if (a.slot_of_plan && tid == 0) a.slot_of_plan[pid] = blockIdx.x;

    /* ---------------- per-plan init (pb:275-306, gs:44-60) */
    for (int i = tid; i < 1024; i += MPLB_NT) table[i] = 0ull;
    if (tid == 0) {
      const mplb_waypoint &st = a.starts[pid];
      const mplb_waypoint &gl = a.goals[pid];
      S.tsize = 1024; S.n_nodes = 0; S.n_heap = 0; S.pops = 0; S.n_closed = 0; S.status = -1; S.done = 0;
      S.goal_hit = 0; S.key_bad = 0; S.n_samples = 0; S.n_valid = 0;
      S.pop_hash = 0xCBF29CE484222325ull; S.closed_hash = 0ull;
      for (int ax = 0; ax < 3; ax++) { S.goal_pos[ax] = gl.pos[ax]; S.goal_vel[ax] = gl.vel[ax]; S.goal_acc[ax] = gl.acc[ax]; }
      double s0[NS];
      for (int ax = 0; ax < DIM; ax++) {
        s0[ax] = st.pos[ax];
        if (ORD >= 2) s0[DIM + ax] = st.vel[ax];
        if (ORD >= 3) s0[2 * DIM + ax] = st.acc[ax];
        if (ORD >= 4) s0[3 * DIM + ax] = st.jrk[ax];
      }
      for (int f = 0; f < NS; f++) S.cur[f] = s0[f];
      /* goal lattice key: comparable only when the goal carries the same control flags (wp:92-125) */
      S.goal_key_ok = 0;
      if (gl.control == c.control && gl.enable_t == 0) {
        double gs[NS];
        for (int ax = 0; ax < DIM; ax++) {
          gs[ax] = gl.pos[ax];
          if (ORD >= 2) gs[DIM + ax] = gl.vel[ax];
          if (ORD >= 3) gs[2 * DIM + ax] = gl.acc[ax];
          if (ORD >= 4) gs[3 * DIM + ax] = gl.jrk[ax];
        }
        int gi[NS];
        lattice_ints<DIM, ORD>(gs, gi);
        unsigned long long gh;
        S.goal_key_ok = pack_key<DIM, ORD>(c, gi, S.gk0, S.gk1, gh) ? 1 : 0;
      }
      if (st.control != c.control || st.enable_t != 0) S.status = MPLB_INTERNAL_BADCTRL;
      else {
        /* pb:283: ENV_->is_free(start.pos) -> mu:44,57-62 on the int8 grid */
        int pn[3] = {0, 0, 0};
        bool outside = false;
        for (int ax = 0; ax < DIM; ax++) {
          pn[ax] = float_to_cell(s0[ax], c.origin[ax], c.res);
          outside = outside || pn[ax] < 0 || pn[ax] >= c.nd[ax];
        }
        bool is_free = false;
        if (!outside) {
          size_t idx = (DIM == 2) ? (size_t)pn[0] + (size_t)c.nd[0] * pn[1]
                                  : (size_t)pn[0] + (size_t)c.nd[0] * pn[1] + (size_t)c.nd[0] * c.nd[1] * pn[2];
          int8_t v = c.grid[idx];
          is_free = v < 100 && v >= 0;
        }
        if (!is_free) S.status = MPLB_PLAN_START_NOT_FREE;
      }
    }
    __syncthreads();
    if (S.status < 0 && warp == 0) { /* gs:44: is_goal(start) */
      bool g0 = goal_test_warp<DIM, ORD>(c, S, S.cur, lane);
      if (lane == 0 && g0) S.status = MPLB_PLAN_START_IS_GOAL;
    }
    __syncthreads();
    if (S.status < 0 && tid == 0) {
      int ints[NS];
      lattice_ints<DIM, ORD>(S.cur, ints);
      unsigned long long k0, k1, kh;
      if (!pack_key<DIM, ORD>(c, ints, k0, k1, kh)) S.status = MPLB_PLAN_KEY_RANGE;
      else {
        NodeHot n0;
        n0.k0 = k0; n0.k1 = k1; n0.kh = kh; n0.g = 0.0;
        n0.h = heuristic<DIM, ORD>(c, S, S.cur, k0, k1);
        n0.pg = 0.0; n0.parent = -1; n0.heap_pos = 0; n0.action = -1; n0.flags = 1; n0.pad0 = 0; n0.pad1 = 0;
        hot[0] = n0;
        for (int f = 0; f < NS; f++) state[f] = S.cur[f];
        table_insert_serial(table, S.tsize, kh, 0);
        HeapEnt e; e.f = dadd(0.0, dmul(c.eps, n0.h)); e.g = 0.0; e.node = 0; e.pad = 0;
        heap[0] = e;
        S.n_nodes = 1; S.n_heap = 1;
      }
    }
    __syncthreads();

    /* ---------------- main loop (gs:63-162) */
    while (S.status < 0) {
      /* capacity: this expansion can add at most nU nodes / heap entries */
      if (S.n_nodes + c.nU > a.cap || S.n_heap + c.nU > a.cap) {
        __syncthreads();
        if (tid == 0) S.status = MPLB_INTERNAL_OVERFLOW;
        __syncthreads();
        break;
      }
      if ((S.n_nodes + c.nU) * 2 > S.tsize) { /* grow the table in place and re-insert every node */
        int nt = S.tsize;
        while ((S.n_nodes + c.nU) * 2 > nt) nt <<= 1;
        __syncthreads();
        if (nt > a.tsize_max) { if (tid == 0) S.status = MPLB_INTERNAL_OVERFLOW; __syncthreads(); break; }
        for (int i = tid; i < nt; i += MPLB_NT) table[i] = 0ull;
        __syncthreads();
        for (int i = tid; i < S.n_nodes; i += MPLB_NT) table_insert_atomic(table, nt, hot[i].kh, i);
        if (tid == 0) S.tsize = nt;
        __syncthreads();
      }
      /* ---- pop (gs:64-68) */
      if (tid == 0) {
        HeapEnt top = heap[0];
        int n = S.n_heap - 1;
        if (n > 0) heap_sift_down(heap, hot, n, 0, heap[n]);
        S.n_heap = n;
        int cn = top.node;
        NodeHot *hn = &hot[cn];
        S.cur_node = cn;
        S.cur_g = hn->g;
        unsigned long long kh = hn->kh;
        S.pop_hash = (S.pop_hash ^ kh) * 0x100000001B3ull;
        unsigned char fl = hn->flags;
        if (!(fl & 2)) { S.n_closed++; S.closed_hash += kh; }
        hn->flags = fl | 2;
        if (a.want_poplog && S.pops < a.cap) poplog[S.pops] = cn;
        S.pops++;
        for (int f = 0; f < NS; f++) S.cur[f] = state[(size_t)cn * NS + f];
        lattice_ints<DIM, ORD>(S.cur, S.cur_ints);
      }
      __syncthreads();
      /* ---- get_succ (em:147-172) */
      for (int i = tid; i < c.nU; i += MPLB_NT) expand_b1<DIM, ORD>(c, S, i);
      __syncthreads();
      expand_b2<DIM, ORD>(c, S, warp, lane, NW);
      __syncthreads();
      /* ---- warp 1: goal test of the popped node (gs:146); warp 0: relax successors in control order (gs:79-143) */
      if (warp == NW - 1) {
        bool gh = goal_test_warp<DIM, ORD>(c, S, S.cur, lane);
        if (lane == 0) S.goal_hit = gh ? 1 : 0;
      }
      if (warp == 0) {
        const int cn = S.cur_node;
        const double cg = S.cur_g;
        for (int b0 = 0; b0 < c.nU; b0 += 32) {
          int i = b0 + lane;
          bool valid = (i < c.nU) && (S.verdict[i] >= 3);
          int nid = -1;
          if (valid) nid = table_find(table, S.tsize, hot, S.kh[i], S.k0[i], S.k1[i]);
          unsigned vm = __ballot_sync(0xffffffffu, valid);
          while (vm) {
            int j = __ffs(vm) - 1;
            vm &= vm - 1;
            int nj = __shfl_sync(0xffffffffu, nid, j);
            if (lane == 0) {
              int idx = b0 + j;
              S.n_valid++;
              if (nj < 0) nj = table_find(table, S.tsize, hot, S.kh[idx], S.k0[idx], S.k1[idx]); /* sibling may have created it */
              if (nj < 0) { /* gs:84-88: create the node; its coord is this (first) discoverer's state */
                nj = S.n_nodes++;
                NodeHot nn;
                nn.k0 = S.k0[idx]; nn.k1 = S.k1[idx]; nn.kh = S.kh[idx];
                nn.g = __longlong_as_double(0x7ff0000000000000ll);
                nn.h = heuristic<DIM, ORD>(c, S, &S.es[idx * NS], nn.k0, nn.k1);
                nn.pg = 0.0; nn.parent = -1; nn.heap_pos = -1; nn.action = -1; nn.flags = 0; nn.pad0 = 0; nn.pad1 = 0;
                hot[nj] = nn;
                for (int f = 0; f < NS; f++) state[(size_t)nj * NS + f] = S.es[idx * NS + f];
                table_insert_serial(table, S.tsize, nn.kh, nj);
              }
              NodeHot *sn = &hot[nj];
              double tentative = dadd(cg, S.cost[idx]); /* gs:107 */
              double gold = sn->g;
              if (tentative < gold) { /* gs:109-141 */
                sn->g = tentative; sn->parent = cn; sn->action = (short)idx; sn->pg = cg;
                HeapEnt e; e.f = dadd(tentative, dmul(c.eps, sn->h)); e.g = tentative; e.node = nj; e.pad = 0;
                unsigned char fl = sn->flags;
                if ((fl & 1) && !(fl & 2)) {
                  heap_sift_up(heap, hot, sn->heap_pos, e); /* increase(): f lowered, sift up only (gs:131-133) */
                } else {
                  if (fl & 2) { /* closed node re-pushed (gs:135-141): refresh g copies of its stale entries */
                    for (int q = 0; q < S.n_heap; q++) if (heap[q].node == nj) heap[q].g = tentative;
                  }
                  sn->flags = fl | 1;
                  heap_sift_up(heap, hot, S.n_heap, e);
                  S.n_heap++;
                }
              } else if (tentative == gold && cg > sn->pg) { /* recoverTraj tie: larger predecessor g wins (gs:398-403) */
                sn->parent = cn; sn->action = (short)idx; sn->pg = cg;
              }
            }
            __syncwarp();
          }
        }
        if (lane == 0) { /* sample counter with the reference's early-exit semantics */
          long long ns = 0;
          for (int i = 0; i < c.nU; i++) {
            int v = S.verdict[i];
            if (v == 3) ns += c.tcnt[S.nsamp[i]];
            else if (v == 2) ns += S.first[i] + 1;
          }
          S.n_samples += ns;
        }
      }
      __syncthreads();
      if (tid == 0) {
        if (S.key_bad) S.status = MPLB_PLAN_KEY_RANGE;
        else if (S.goal_hit) S.status = MPLB_PLAN_OK;
        else if (c.max_num > 0 && S.pops >= c.max_num) S.status = MPLB_PLAN_MAX_EXPAND;
        else if (S.n_heap == 0) S.status = MPLB_PLAN_QUEUE_EMPTY;
      }
      __syncthreads();
    }

    /* ---------------- results + recoverTraj (gs:369-455) */
    if (tid == 0) {
      mplb_result r;
      r.status = S.status; r.n_seg = 0; r.cost = __longlong_as_double(0x7ff0000000000000ll);
      r.pops = S.pops; r.n_nodes = S.n_nodes; r.n_open = S.n_heap; r.n_closed = S.n_closed;
      r.n_prims = (long long)S.pops * c.nU; r.n_samples = S.n_samples; r.n_valid = S.n_valid;
      r.pop_hash = S.pop_hash; r.closed_hash = S.closed_hash;
      if (S.status == MPLB_PLAN_START_IS_GOAL) r.cost = 0.0;
      int *acts = a.actions ? a.actions + (size_t)pid * a.max_seg : nullptr;
      if (acts) for (int k = 0; k < a.max_seg; k++) acts[k] = -1;
      if (S.status == MPLB_PLAN_OK) {
        int n = 0, cnode = S.cur_node;
        bool ok = true;
        while (cnode != 0) {
          int p = hot[cnode].parent;
          if (p < 0 || n > S.n_nodes) { ok = false; break; }
          n++; cnode = p;
        }
        if (!ok) r.status = MPLB_PLAN_TRACEBACK_FAILED;
        else {
          r.n_seg = n;
          r.cost = hot[S.cur_node].g; /* gs:179 */
          cnode = S.cur_node;
          for (int k = n - 1; k >= 0; k--) {
            int p = hot[cnode].parent;
            if (k < a.max_seg) {
              if (acts) acts[k] = hot[cnode].action;
              if (a.seg_states) {
                double *row = a.seg_states + ((size_t)pid * a.max_seg + k) * 13;
                for (int q = 0; q < 13; q++) row[q] = 0.0;
                for (int d = 0; d < ORD; d++)
                  for (int ax = 0; ax < DIM; ax++) row[d * 3 + ax] = state[(size_t)p * NS + d * DIM + ax];
              }
            }
            cnode = p;
          }
        }
      }
      a.results[pid] = r;
      if (r.status == MPLB_INTERNAL_OVERFLOW) a.overflow_list[atomicAdd(a.overflow_count, 1)] = pid;
    }
  }
}

/* ---------------------------------------------------------------- get_succ for arbitrary states (parity artefact) */
template <int DIM, int ORD>
__global__ void __launch_bounds__(MPLB_NT) expand_trace_kernel(const DevCfg c, const mplb_waypoint *states, int n_states,
                                                               mplb_prim_trace *rows) {
  constexpr int NS = DIM * ORD;
  constexpr int NW = MPLB_NT / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PlanSmem<DIM, ORD> &S = *reinterpret_cast<PlanSmem<DIM, ORD> *>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < c.nU * 3; i += MPLB_NT) S.U[i] = c.U[i];
  for (int s = blockIdx.x; s < n_states; s += gridDim.x) {
    __syncthreads();
    if (tid == 0) {
      const mplb_waypoint &st = states[s];
      for (int ax = 0; ax < DIM; ax++) {
        S.cur[ax] = st.pos[ax];
        if (ORD >= 2) S.cur[DIM + ax] = st.vel[ax];
        if (ORD >= 3) S.cur[2 * DIM + ax] = st.acc[ax];
        if (ORD >= 4) S.cur[3 * DIM + ax] = st.jrk[ax];
      }
      lattice_ints<DIM, ORD>(S.cur, S.cur_ints);
      S.key_bad = 0;
    }
    __syncthreads();
    for (int i = tid; i < c.nU; i += MPLB_NT) expand_b1<DIM, ORD>(c, S, i);
    __syncthreads();
    expand_b2<DIM, ORD>(c, S, warp, lane, NW);
    __syncthreads();
    for (int i = tid; i < c.nU; i += MPLB_NT) {
      mplb_prim_trace r;
      int v = S.verdict[i];
      r.verdict = v;
      r.n = (v == 2 || v == 3) ? S.nsamp[i] : 0;
      r.n_tested = (v == 3) ? c.tcnt[S.nsamp[i]] : (v == 2 ? S.first[i] + 1 : 0);
      r.block_idx = -1;
      if (v == 2) {
        int cell = -1;
        sample_blocked<DIM, ORD>(c, S, i, c.ttab[c.toff[S.nsamp[i]] + S.first[i]], &cell);
        r.block_idx = cell;
      }
      r.cost = (v >= 3) ? S.cost[i] : (v == 2 ? __longlong_as_double(0x7ff0000000000000ll) : 0.0);
      for (int q = 0; q < 13; q++) r.succ[q] = 0.0;
      for (int ax = 0; ax < DIM; ax++) { /* tn carries every derivative (pr:321-331) */
        Axis<ORD> A(&S.cur[ax], DIM, S.U[i * 3 + ax]);
        r.succ[ax] = A.p(c.dt); r.succ[3 + ax] = A.v(c.dt); r.succ[6 + ax] = A.a(c.dt); r.succ[9 + ax] = A.j(c.dt);
      }
      int ints[NS];
      lattice_ints<DIM, ORD>(&S.es[i * NS], ints);
      for (int q = 0; q < 16; q++) r.key[q] = 0;
      for (int f = 0; f < NS; f++) r.key[f] = ints[f];
      r.key[15] = NS;
      rows[(size_t)s * c.nU + i] = r;
    }
  }
}

}  // namespace mplb
