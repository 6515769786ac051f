/*
 * mplb_search.cuh — the fused A* kernel: one CTA owns one plan; a persistent grid drains a batch.
 *
 * Reference path restated per pop (paths under motion_primitive_library/include/):
 *   GraphSearch::Astar            mpl_planner/common/graph_search.h:39-182   (pop, relax, terminate)
 *   env_map::get_succ             mpl_planner/env/env_map.h:147-172          (phase B1: one thread per control u)
 *   env_map::traverse_primitive   mpl_planner/env/env_map.h:90-132           (phase B2: one thread per sample)
 *   env_map::is_goal + rayTrace   env_map.h:25-45, mpl_collision/map_util.h:117-134
 *   priorityQueue / compare_pair  mpl_planner/common/state_space.h:15-34     (binary heap, same sift rules as
 *                                 boost::heap::d_ary_heap<arity<2>, mutable_<true>> so pop order is identical)
 *   recoverTraj                   graph_search.h:369-455 (best-predecessor rule kept as a running argmin)
 *
 * Latency structure of one pop (the search of one plan is a serial chain of pops, so the kernel is
 * latency-bound per plan and throughput comes from the batch):
 *   P1  warp 0, one lane per control: end state, dynamic validation, lattice key, sample count     (registers/smem)
 *   P2  all threads: every collision sample of every control in one flat pass (bit-brick loads, L2/L1);
 *       concurrently lanes of warp 0 probe the hash table and prefetch the touched node records    (HBM/L2, overlapped)
 *   P3  warp 0: goal test, relax successors in control order against the SHARED-MEMORY heap using the
 *       prefetched records, pop the next node (its state row was prefetched or is forwarded from smem)
 * Three __syncthreads per pop; no global load sits on the serial chain in the common case.
 *
 * Search state of a plan lives in a private HBM arena (node records, state rows, table, heap spill); the only
 * data shared between CTAs is the read-only map (bit-bricks) and the control / sample-time tables.
 */
#pragma once
#include "mplb_device.cuh"
#include "../../include/mplb.h"

namespace mplb {

#define MPLB_INTERNAL_OVERFLOW 100 /* arena too small: host retries the plan in a larger tier */
#define MPLB_INTERNAL_BADCTRL 9
#define MPLB_OWNER_CAP 2304 /* flat sample list capacity (>= 27*41 and 125*16) */

/* Relax-path half of a node: one 32-byte sector. */
struct __align__(32) NodeHot {
  double g, h;
  double pg;       /* g of the best predecessor (tie rule of recoverTraj, gs:391-405) */
  int heap_pos;    /* position of the live heap entry while open */
  short action;    /* action id of parent -> this */
  unsigned char flags; /* 1 = iterationopened, 2 = iterationclosed */
  unsigned char pad0;
};
static_assert(sizeof(NodeHot) == 32, "NodeHot must be 32 bytes");

/* Pop-path half of a node: header of the state row, followed by NS doubles of state. */
struct __align__(16) RowHdr {
  unsigned long long k0, k1; /* packed lattice key */
  unsigned long long kh;     /* 64-bit hash of the lattice ints */
  int parent;                /* best predecessor node */
  int pad;
};
static_assert(sizeof(RowHdr) == 32, "RowHdr must be 32 bytes");

/* Table slot: exact for keys up to 96 bits; wider keys also compare RowHdr::k1. node1 == 0 means empty. */
struct __align__(16) Slot {
  unsigned long long k0;
  unsigned int k1lo;
  unsigned int node1;
};
static_assert(sizeof(Slot) == 16, "Slot must be 16 bytes");

struct HeapEnt {
  double f; /* heap key (gs:54,119) */
  double g; /* copy of node g for compare_pair's tie-break (ss:19-24) */
  int node; /* bit 31 set: entry is a re-push of an already closed node (gs:135-141) */
  int pad;
};
static_assert(sizeof(HeapEnt) == 24, "HeapEnt must be 24 bytes");

/* cmp(a,b) of compare_pair: true iff a is worse (lower priority) than b. */
__device__ __forceinline__ bool heap_worse(double af, double ag, double bf, double bg) {
  return (af == bf) ? (ag > bg) : (af > bf);
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
  size_t off_rows, off_heap, off_table, off_poplog;
  int want_poplog;
  int *slot_of_plan;   /* optional: which slot ran plan i (retained single plan) */
  int *overflow_count; /* plans whose arena overflowed in this tier ... */
  int *overflow_list;  /* ... and their ids, for the next (larger) tier */
};

template <int DIM, int ORD, int MAXU>
struct PlanSmem {
  static constexpr int NS = DIM * ORD;
  static constexpr int HCAP = (MAXU <= 32) ? 2048 : 1024; /* heap entries kept in shared memory */
  /* heap top (SoA) */
  double hf[HCAP], hg[HCAP];
  int hn[HCAP];
  /* current node */
  double cur[NS];
  int cur_ints[NS];
  double cur_g;
  int cur_node;
  double goal_pos[3], goal_vel[3], goal_acc[3];
  unsigned long long gk0, gk1;
  int goal_key_ok;
  double U[MAXU * 3];
  /* per-control results of get_succ */
  double es[MAXU * NS]; /* end states, [u][d*DIM+ax] */
  double cost[MAXU];
  unsigned long long k0[MAXU], k1[MAXU], kh[MAXU];
  int verdict[MAXU]; /* 0 self, 1 dyn, 2 blocked, 3 valid, 4 valid-no-motion, 5 needs sampling (transient) */
  int nsamp[MAXU];   /* divisor n */
  int cnt[MAXU];     /* samples to test */
  int pre[MAXU];     /* exclusive prefix of cnt */
  int first[MAXU];   /* first blocked sample index or INT_MAX */
  /* prefetched node records for the relax phase */
  int nid[MAXU];     /* node id or -1 */
  int slot[MAXU];    /* empty table slot where the probe ended (when nid == -1) */
  double ng[MAXU], nh[MAXU], npg[MAXU];
  int npos[MAXU];
  int nfl[MAXU];
  int cr_idx[MAXU];  /* successors that created a node in this expansion */
  int n_created;
  int n_before;      /* n_nodes before this expansion */
  int m_total;       /* flat sample count */
  unsigned char owner[MPLB_OWNER_CAP];
  int n_nodes, n_heap, tsize, pops, n_closed, status, plan_idx, key_bad;
  long long n_samples, n_valid;
  unsigned long long pop_hash, closed_hash;
};

/* ---------------------------------------------------------------- heap in shared memory with global spill */
template <class SM>
struct HeapView {
  SM &S;
  HeapEnt *spill; /* global array indexed by heap position (entries >= HCAP live here) */
  int *heap_pos_base; /* &hot[0].heap_pos, stride sizeof(NodeHot) */
  __device__ __forceinline__ void get(int i, double &f, double &g, int &n) const {
    if (i < SM::HCAP) { f = S.hf[i]; g = S.hg[i]; n = S.hn[i]; }
    else { HeapEnt e = spill[i]; f = e.f; g = e.g; n = e.node; }
  }
  __device__ __forceinline__ void set(int i, double f, double g, int n, NodeHot *hot) const {
    if (i < SM::HCAP) { S.hf[i] = f; S.hg[i] = g; S.hn[i] = n; }
    else { HeapEnt e; e.f = f; e.g = g; e.node = n; e.pad = 0; spill[i] = e; }
    hot[n & 0x7fffffff].heap_pos = i;
  }
  __device__ __forceinline__ int node_at(int i) const { return (i < SM::HCAP) ? S.hn[i] : spill[i].node; }
  __device__ __forceinline__ void set_g(int i, double g) const { if (i < SM::HCAP) S.hg[i] = g; else spill[i].g = g; }

  /* push/increase: sift up while the parent is strictly worse (boost siftup) */
  __device__ __forceinline__ void sift_up(int pos, double f, double g, int n, NodeHot *hot) const {
    while (pos != 0) {
      int par = (pos - 1) >> 1;
      double pf, pg; int pn;
      get(par, pf, pg, pn);
      if (heap_worse(pf, pg, f, g)) { set(pos, pf, pg, pn, hot); pos = par; }
      else break;
    }
    set(pos, f, g, n, hot);
  }
  /* pop: sift the former last element down from the root; ties still move down (boost siftdown) */
  __device__ __forceinline__ void sift_down(int n_heap, int pos, double f, double g, int n, NodeHot *hot) const {
    while (true) {
      int c = 2 * pos + 1;
      if (c >= n_heap) break;
      double cf, cg; int cn;
      get(c, cf, cg, cn);
      if (c + 1 < n_heap) {
        double rf, rg; int rn;
        get(c + 1, rf, rg, rn);
        if (heap_worse(cf, cg, rf, rg)) { c = c + 1; cf = rf; cg = rg; cn = rn; } /* right child only if strictly better */
      }
      if (!heap_worse(cf, cg, f, g)) { set(pos, cf, cg, cn, hot); pos = c; }
      else break;
    }
    set(pos, f, g, n, hot);
  }
};

/* ---------------------------------------------------------------- hash table */
__device__ __forceinline__ bool slot_matches(const Slot &s, unsigned long long k0, unsigned long long k1, bool wide,
                                             const unsigned char *rows, size_t row_bytes) {
  if (s.k0 != k0 || s.k1lo != (unsigned int)k1) return false;
  if (!wide) return true;
  const RowHdr *h = reinterpret_cast<const RowHdr *>(rows + (size_t)(s.node1 - 1) * row_bytes);
  return h->k1 == k1;
}

/* returns node id or -1; *end_slot = slot index where the probe stopped (empty slot when -1) */
__device__ __forceinline__ int table_find(const Slot *table, int tsize, unsigned long long kh, unsigned long long k0,
                                          unsigned long long k1, bool wide, const unsigned char *rows, size_t row_bytes,
                                          int *end_slot) {
  unsigned mask = (unsigned)tsize - 1u;
  unsigned i = (unsigned)kh & mask;
  while (true) {
    Slot s = table[i];
    if (s.node1 == 0u) { *end_slot = (int)i; return -1; }
    if (slot_matches(s, k0, k1, wide, rows, row_bytes)) { *end_slot = (int)i; return (int)s.node1 - 1; }
    i = (i + 1) & mask;
  }
}

__device__ __forceinline__ void table_insert_atomic(Slot *table, int tsize, unsigned long long kh, unsigned long long k0,
                                                    unsigned long long k1, int id) {
  unsigned mask = (unsigned)tsize - 1u;
  unsigned i = (unsigned)kh & mask;
  unsigned long long w1 = ((unsigned long long)(unsigned)(id + 1) << 32) | (unsigned long long)(unsigned int)k1;
  unsigned long long *t = reinterpret_cast<unsigned long long *>(table);
  while (atomicCAS(&t[2 * i + 1], 0ull, w1) != 0ull) i = (i + 1) & mask; /* claim by the (k1lo,node1) word */
  t[2 * i] = k0;
}

/* ---------------------------------------------------------------- phase B1: one thread per control (em:155-160,163-165) */
template <int DIM, int ORD, class SM>
__device__ __forceinline__ void expand_b1(const DevCfg &c, SM &S, int i) {
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
  S.cost[i] = dadd(J, dmul(c.w, T)); /* eb:343-345; traverse contributes 0 on a plain map */
  int verdict, n = 0, cnt = 0;
  if (self) verdict = 0;
  else if (!dyn_ok) verdict = 1;
  else if (same_pos) verdict = 4;
  else {
    verdict = 5;
    n = __double2int_rz(ceil(ddiv(dmul(max_v, T), c.res))); /* em:95 */
    n = n < 5 ? 5 : n;
    cnt = c.tcnt[n];
  }
  S.nsamp[i] = n;
  S.cnt[i] = cnt;
  if ((verdict >= 3) && !key_ok) S.key_bad = 1;
  S.verdict[i] = verdict;
}

/* One collision sample (em:100-104,119): true when the sample at time t of control i is outside or occupied. */
template <int DIM, int ORD, class SM>
__device__ __forceinline__ bool sample_blocked(const DevCfg &c, const SM &S, int i, double t, int *cell_idx) {
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

/* B2, per-control form: warps take controls round-robin, lanes take samples (used by the trace kernel and when
 * the flat sample list would overflow). */
template <int DIM, int ORD, class SM>
__device__ __forceinline__ void expand_b2_percontrol(const DevCfg &c, SM &S, int warp, int lane, int nwarps) {
  for (int i = warp; i < c.nU; i += nwarps) {
    if (S.verdict[i] != 5) continue;
    int n = S.nsamp[i];
    int cnt = S.cnt[i];
    const double *tt = c.ttab + c.toff[n];
    int first = 0x7fffffff;
    for (int base = 0; base < cnt; base += 32) {
      int k = base + lane;
      bool blocked = false;
      if (k < cnt) blocked = sample_blocked<DIM, ORD>(c, S, i, __ldg(&tt[k]), nullptr);
      unsigned m = __ballot_sync(0xffffffffu, blocked);
      if (m) { first = base + __ffs(m) - 1; break; }
    }
    if (lane == 0) S.first[i] = first;
  }
}

/* em:25-45 for a state (tolerances, then ray trace mu:117-134); executed by one full warp. */
template <int DIM, int ORD, class SM>
__device__ __forceinline__ bool goal_test_warp(const DevCfg &c, const SM &S, const double *st, int lane) {
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
template <int DIM, int ORD, class SM>
__device__ __forceinline__ double heuristic(const DevCfg &c, const SM &S, const double *st, unsigned long long k0,
                                            unsigned long long k1) {
  if (c.eps == 0.0) return 0.0; /* gs:53,87 */
  if (S.goal_key_ok && k0 == S.gk0 && k1 == S.gk1) return 0.0;
  double m = 0.0;
#pragma unroll
  for (int ax = 0; ax < DIM; ax++) m = fmax(m, fabs(dsub(st[ax], S.goal_pos[ax])));
  if (c.v_max > 0.0) return ddiv(dmul(c.w, m), c.v_max);
  return dmul(c.w, m);
}

/* unpack the lattice ints of a node from its packed key */
template <int NS>
__device__ __forceinline__ void unpack_ints(const DevCfg &c, unsigned long long k0, unsigned long long k1, int *ints) {
#pragma unroll
  for (int f = 0; f < NS; f++) {
    unsigned long long wv = c.kword[f] ? k1 : k0;
    unsigned long long v = (wv >> c.kshift[f]) & ((1ull << c.kbits[f]) - 1ull);
    ints[f] = (int)((long long)v + (long long)c.koff[f]);
  }
}

/* ---------------------------------------------------------------- the kernel */
template <int DIM, int ORD, int MAXU>
__global__ void __launch_bounds__(MPLB_NT) astar_batch_kernel(const DevCfg c, const BatchArgs a) {
  constexpr int NS = DIM * ORD;
  constexpr int NW = MPLB_NT / 32;
  using SM = PlanSmem<DIM, ORD, MAXU>;
  constexpr size_t ROWB = sizeof(RowHdr) + NS * sizeof(double);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SM &S = *reinterpret_cast<SM *>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double kInf = __longlong_as_double(0x7ff0000000000000ll);

  unsigned char *base = a.arena + (size_t)blockIdx.x * a.stride;
  NodeHot *hot = reinterpret_cast<NodeHot *>(base);
  unsigned char *rows = base + a.off_rows;
  HeapEnt *spill = reinterpret_cast<HeapEnt *>(base + a.off_heap);
  Slot *table = reinterpret_cast<Slot *>(base + a.off_table);
  int *poplog = reinterpret_cast<int *>(base + a.off_poplog);
  const bool wide = c.key_wide != 0;
  HeapView<SM> H{S, spill, nullptr};

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
    {
      unsigned long long *t64 = reinterpret_cast<unsigned long long *>(table);
      for (int i = tid; i < 2 * 1024; i += MPLB_NT) t64[i] = 0ull;
    }
    if (tid == 0) {
      const mplb_waypoint &st = a.starts[pid];
      const mplb_waypoint &gl = a.goals[pid];
      S.tsize = 1024; S.n_nodes = 0; S.n_heap = 0; S.pops = 0; S.n_closed = 0; S.status = -1;
      S.key_bad = 0; S.n_samples = 0; S.n_valid = 0; S.n_created = 0; S.n_before = 0;
      S.pop_hash = 0ull; S.closed_hash = 0ull;
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
    if (S.status < 0 && tid == 0) { /* gs:47-60: start node, pushed and immediately popped as the first current node */
      int ints[NS];
      lattice_ints<DIM, ORD>(S.cur, ints);
      unsigned long long k0, k1, kh;
      if (!pack_key<DIM, ORD>(c, ints, k0, k1, kh)) S.status = MPLB_PLAN_KEY_RANGE;
      else {
        NodeHot n0;
        n0.g = 0.0; n0.h = heuristic<DIM, ORD>(c, S, S.cur, k0, k1); n0.pg = 0.0; n0.heap_pos = 0; n0.action = -1;
        n0.flags = 3; n0.pad0 = 0; /* opened, and closed by the first pop below */
        hot[0] = n0;
        RowHdr *rh = reinterpret_cast<RowHdr *>(rows);
        rh->k0 = k0; rh->k1 = k1; rh->kh = kh; rh->parent = -1; rh->pad = 0;
        double *rs = reinterpret_cast<double *>(rows + sizeof(RowHdr));
        for (int f = 0; f < NS; f++) rs[f] = S.cur[f];
        table_insert_atomic(table, S.tsize, kh, k0, k1, 0);
        S.n_nodes = 1;
        /* first pop (gs:64-68): the heap holds exactly the start node */
        S.n_heap = 0;
        S.cur_node = 0; S.cur_g = 0.0;
        for (int f = 0; f < NS; f++) S.cur_ints[f] = ints[f];
        S.pop_hash = (0xCBF29CE484222325ull ^ kh) * 0x100000001B3ull;
        S.n_closed = 1; S.closed_hash = kh;
        if (a.want_poplog) poplog[0] = 0;
        S.pops = 1;
      }
    }
    __syncthreads();

    /* ---------------- main loop (gs:63-162): S.cur* always holds the node popped last */
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
        unsigned long long *t64 = reinterpret_cast<unsigned long long *>(table);
        for (int i = tid; i < 2 * nt; i += MPLB_NT) t64[i] = 0ull;
        __syncthreads();
        for (int i = tid; i < S.n_nodes; i += MPLB_NT) {
          const RowHdr *rh = reinterpret_cast<const RowHdr *>(rows + (size_t)i * ROWB);
          table_insert_atomic(table, nt, rh->kh, rh->k0, rh->k1, i);
        }
        if (tid == 0) S.tsize = nt;
        __syncthreads();
      }

      /* ---- P1: get_succ, one lane per control (em:147-172) */
      if (MAXU <= 32) {
        if (warp == 0) {
          int cnt = 0;
          if (lane < c.nU) { expand_b1<DIM, ORD>(c, S, lane); cnt = S.cnt[lane]; }
          int incl = cnt; /* warp scan of the sample counts */
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
          int excl = incl - cnt;
          if (lane < c.nU) {
            S.pre[lane] = excl;
            if (incl <= MPLB_OWNER_CAP) for (int k = 0; k < cnt; k++) S.owner[excl + k] = (unsigned char)lane;
          }
          if (lane == 31) S.m_total = incl;
        }
        __syncthreads();
      } else {
        for (int i = tid; i < c.nU; i += MPLB_NT) expand_b1<DIM, ORD>(c, S, i);
        __syncthreads();
        if (tid < c.nU) {
          int excl = 0;
          for (int j = 0; j < tid; j++) excl += S.cnt[j];
          S.pre[tid] = excl;
          int cnt = S.cnt[tid];
          if (excl + cnt <= MPLB_OWNER_CAP) for (int k = 0; k < cnt; k++) S.owner[excl + k] = (unsigned char)tid;
          if (tid == c.nU - 1) S.m_total = excl + cnt;
        }
        __syncthreads();
      }

      /* ---- P2: table probes + node prefetch (lanes of the first warps), then all collision samples flat */
      for (int i = tid; i < c.nU; i += MPLB_NT) {
        int v = S.verdict[i];
        int nid = -1, slot = -1;
        if (v >= 4) { /* 4 = valid without sampling, 5 = outcome pending: probe speculatively */
          nid = table_find(table, S.tsize, S.kh[i], S.k0[i], S.k1[i], wide, rows, ROWB, &slot);
          if (nid >= 0) {
            const NodeHot hn = hot[nid];
            S.ng[i] = hn.g; S.nh[i] = hn.h; S.npg[i] = hn.pg; S.npos[i] = hn.heap_pos; S.nfl[i] = hn.flags;
          }
        }
        S.nid[i] = nid; S.slot[i] = slot;
      }
      if (S.m_total <= MPLB_OWNER_CAP) {
        for (int m = tid; m < S.m_total; m += MPLB_NT) {
          int u = S.owner[m];
          int k = m - S.pre[u];
          double t = __ldg(&c.ttab[c.toff[S.nsamp[u]] + k]);
          if (sample_blocked<DIM, ORD>(c, S, u, t, nullptr)) atomicMin(&S.first[u], k);
        }
      } else {
        expand_b2_percontrol<DIM, ORD>(c, S, warp, lane, NW);
      }
      __syncthreads();

      /* ---- P3 (warp 0): goal test of the current node (gs:146), relax (gs:79-143), terminate, pop the next node */
      if (warp == 0) {
        const int cn = S.cur_node;
        const double cg = S.cur_g;
        const bool goal_hit = goal_test_warp<DIM, ORD>(c, S, S.cur, lane);
        /* prefetch the state row of the present heap top: it is the next pop unless a successor overtakes it */
        int pf_node = -1;
        unsigned long long pf_k0 = 0, pf_k1 = 0, pf_kh = 0;
        double pf_st[NS];
        if (lane == 0 && S.n_heap > 0) {
          pf_node = S.hn[0] & 0x7fffffff;
          const RowHdr *rh = reinterpret_cast<const RowHdr *>(rows + (size_t)pf_node * ROWB);
          pf_k0 = rh->k0; pf_k1 = rh->k1; pf_kh = rh->kh;
          const double *rs = reinterpret_cast<const double *>(rows + (size_t)pf_node * ROWB + sizeof(RowHdr));
#pragma unroll
          for (int f = 0; f < NS; f++) pf_st[f] = rs[f];
        }
        if (lane == 0) { S.n_before = S.n_nodes; S.n_created = 0; }
        __syncwarp();
        long long ns_acc = 0;
        for (int idx = 0; idx < c.nU; idx++) {
          int v = S.verdict[idx];
          if (v == 5) { /* finalize the collision outcome */
            int first = S.first[idx];
            v = (first == 0x7fffffff) ? 3 : 2;
            ns_acc += (v == 3) ? S.cnt[idx] : first + 1;
          }
          if (v < 3) continue;
          /* resolve the node: prefetched id, or one created earlier in this expansion, or a new one */
          int nid = S.nid[idx];
          const unsigned long long k0 = S.k0[idx], k1 = S.k1[idx];
          if (nid < 0 && S.n_created > 0) {
            int hit = -1;
            for (int q = lane; q < S.n_created; q += 32) {
              int j = S.cr_idx[q];
              if (S.k0[j] == k0 && S.k1[j] == k1) hit = j;
            }
            unsigned bm = __ballot_sync(0xffffffffu, hit >= 0);
            if (bm) {
              int j = __shfl_sync(0xffffffffu, hit, __ffs(bm) - 1);
              nid = S.nid[j];
              if (lane == 0) { S.nid[idx] = nid; S.ng[idx] = S.ng[j]; S.nh[idx] = S.nh[j]; S.npg[idx] = S.npg[j]; S.npos[idx] = S.npos[j]; S.nfl[idx] = S.nfl[j]; }
              __syncwarp();
            }
          }
          if (nid < 0) { /* gs:84-88: create the node; its coord is this (first) discoverer's state */
            int slot = S.slot[idx];
            /* the probe's end slot may have been taken by a node created earlier in this expansion */
            bool taken = false;
            for (int q = lane; q < S.n_created; q += 32) taken = taken || (S.slot[S.cr_idx[q]] == slot);
            taken = __any_sync(0xffffffffu, taken);
            if (lane == 0) {
              if (taken) {
                unsigned mask = (unsigned)S.tsize - 1u;
                unsigned i = (unsigned)slot;
                while (table[i].node1 != 0u) i = (i + 1) & mask;
                slot = (int)i;
                S.slot[idx] = slot;
              }
              nid = S.n_nodes++;
              double hval = heuristic<DIM, ORD>(c, S, &S.es[idx * NS], k0, k1);
              RowHdr *rh = reinterpret_cast<RowHdr *>(rows + (size_t)nid * ROWB);
              rh->k0 = k0; rh->k1 = k1; rh->kh = S.kh[idx]; rh->parent = -1; rh->pad = 0;
              double *rs = reinterpret_cast<double *>(rows + (size_t)nid * ROWB + sizeof(RowHdr));
#pragma unroll
              for (int f = 0; f < NS; f++) rs[f] = S.es[idx * NS + f];
              Slot sl; sl.k0 = k0; sl.k1lo = (unsigned int)k1; sl.node1 = (unsigned)(nid + 1);
              table[slot] = sl;
              S.nid[idx] = nid; S.ng[idx] = kInf; S.nh[idx] = hval; S.npg[idx] = 0.0; S.npos[idx] = -1; S.nfl[idx] = 0;
              S.cr_idx[S.n_created] = idx;
              S.n_created = S.n_created + 1;
            }
            __syncwarp();
            nid = S.nid[idx];
          }
          /* relax (lane 0) */
          double new_g = 0.0, new_pg = 0.0; int new_fl = 0, new_pos = 0; bool changed = false;
          if (lane == 0) {
            S.n_valid++;
            double tentative = dadd(cg, S.cost[idx]); /* gs:107 */
            double gold = S.ng[idx];
            int fl = S.nfl[idx];
            if (tentative < gold) { /* gs:109-141 */
              double f = dadd(tentative, dmul(c.eps, S.nh[idx]));
              NodeHot hn; hn.g = tentative; hn.h = S.nh[idx]; hn.pg = cg; hn.action = (short)idx; hn.pad0 = 0;
              reinterpret_cast<RowHdr *>(rows + (size_t)nid * ROWB)->parent = cn;
              if ((fl & 1) && !(fl & 2)) { /* increase(): f lowered, sift up only (gs:131-133) */
                int pos = S.npos[idx];
                if (pos < 0 || pos >= S.n_heap || (H.node_at(pos) & 0x7fffffff) != nid) pos = hot[nid].heap_pos; /* stale cache */
                hn.flags = (unsigned char)fl; hn.heap_pos = pos;
                hot[nid] = hn;
                H.sift_up(pos, f, tentative, nid, hot);
              } else {
                int tag = nid;
                if (fl & 2) { /* closed node re-pushed (gs:135-141): refresh g copies of its stale entries */
                  for (int q = 0; q < S.n_heap; q++) if ((H.node_at(q) & 0x7fffffff) == nid) H.set_g(q, tentative);
                  tag = nid | 0x80000000;
                }
                fl |= 1;
                hn.flags = (unsigned char)fl; hn.heap_pos = S.n_heap;
                hot[nid] = hn;
                H.sift_up(S.n_heap, f, tentative, tag, hot);
                S.n_heap++;
              }
              changed = true; new_g = tentative; new_pg = cg; new_fl = fl; new_pos = -2; /* position unknown: re-read on use */
            } else if (tentative == gold && cg > S.npg[idx]) { /* recoverTraj tie: larger predecessor g wins (gs:398-403) */
              hot[nid].pg = cg; hot[nid].action = (short)idx;
              reinterpret_cast<RowHdr *>(rows + (size_t)nid * ROWB)->parent = cn;
              changed = true; new_g = gold; new_pg = cg; new_fl = fl; new_pos = S.npos[idx];
            }
          }
          /* keep the cached copies of this node coherent for later successors that map to it */
          changed = __shfl_sync(0xffffffffu, changed, 0);
          if (changed) {
            new_g = __shfl_sync(0xffffffffu, new_g, 0); new_pg = __shfl_sync(0xffffffffu, new_pg, 0);
            new_fl = __shfl_sync(0xffffffffu, new_fl, 0); new_pos = __shfl_sync(0xffffffffu, new_pos, 0);
            for (int j = lane; j < c.nU; j += 32)
              if (S.nid[j] == nid) { S.ng[j] = new_g; S.npg[j] = new_pg; S.nfl[j] = new_fl; S.npos[j] = new_pos; }
          }
          __syncwarp();
        }
        /* ---- termination (gs:146-161) and the next pop (gs:64-68) */
        if (lane == 0) {
          S.n_samples += ns_acc;
          int status = -1;
          if (S.key_bad) status = MPLB_PLAN_KEY_RANGE;
          else if (goal_hit) status = MPLB_PLAN_OK;
          else if (c.max_num > 0 && S.pops >= c.max_num) status = MPLB_PLAN_MAX_EXPAND;
          else if (S.n_heap == 0) status = MPLB_PLAN_QUEUE_EMPTY;
          if (status >= 0) S.status = status;
          else {
            int tagged = S.hn[0];
            double topg = S.hg[0];
            int n = S.n_heap - 1;
            if (n > 0) { double lf, lg; int ln; H.get(n, lf, lg, ln); H.sift_down(n, 0, lf, lg, ln, hot); }
            S.n_heap = n;
            int nx = tagged & 0x7fffffff;
            unsigned long long k0, k1, kh;
            if (nx >= S.n_before) { /* created in this expansion: forward its state from shared memory */
              int j = -1;
              for (int q = 0; q < S.n_created; q++) if (S.nid[S.cr_idx[q]] == nx) { j = S.cr_idx[q]; break; }
              k0 = S.k0[j]; k1 = S.k1[j]; kh = S.kh[j];
#pragma unroll
              for (int f = 0; f < NS; f++) pf_st[f] = S.es[j * NS + f];
            } else if (nx == pf_node) {
              k0 = pf_k0; k1 = pf_k1; kh = pf_kh;
            } else {
              const RowHdr *rh = reinterpret_cast<const RowHdr *>(rows + (size_t)nx * ROWB);
              k0 = rh->k0; k1 = rh->k1; kh = rh->kh;
              const double *rs = reinterpret_cast<const double *>(rows + (size_t)nx * ROWB + sizeof(RowHdr));
#pragma unroll
              for (int f = 0; f < NS; f++) pf_st[f] = rs[f];
            }
#pragma unroll
            for (int f = 0; f < NS; f++) S.cur[f] = pf_st[f];
            unpack_ints<NS>(c, k0, k1, S.cur_ints);
            S.cur_node = nx;
            S.cur_g = topg;
            S.pop_hash = (S.pop_hash ^ kh) * 0x100000001B3ull;
            if (!(tagged & 0x80000000)) { S.n_closed++; S.closed_hash += kh; }
            hot[nx].flags = 3; /* iterationclosed = true (gs:68); a popped node is always opened */
            if (a.want_poplog && S.pops < a.cap) poplog[S.pops] = nx;
            S.pops++;
          }
        }
      }
      __syncthreads();
    }

    /* ---------------- results + recoverTraj (gs:369-455) */
    if (tid == 0) {
      mplb_result r;
      r.status = S.status; r.n_seg = 0; r.cost = kInf;
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
          int p = reinterpret_cast<const RowHdr *>(rows + (size_t)cnode * ROWB)->parent;
          if (p < 0 || n > S.n_nodes) { ok = false; break; }
          n++; cnode = p;
        }
        if (!ok) r.status = MPLB_PLAN_TRACEBACK_FAILED;
        else {
          r.n_seg = n;
          r.cost = hot[S.cur_node].g; /* gs:179 */
          cnode = S.cur_node;
          for (int k = n - 1; k >= 0; k--) {
            int p = reinterpret_cast<const RowHdr *>(rows + (size_t)cnode * ROWB)->parent;
            if (k < a.max_seg) {
              if (acts) acts[k] = hot[cnode].action;
              if (a.seg_states) {
                double *row = a.seg_states + ((size_t)pid * a.max_seg + k) * 13;
                const double *ps = reinterpret_cast<const double *>(rows + (size_t)p * ROWB + sizeof(RowHdr));
                for (int q = 0; q < 13; q++) row[q] = 0.0;
                for (int d = 0; d < ORD; d++)
                  for (int ax = 0; ax < DIM; ax++) row[d * 3 + ax] = ps[d * DIM + ax];
              }
            }
            cnode = p;
          }
        }
      }
      a.results[pid] = r;
      if (r.status == MPLB_INTERNAL_OVERFLOW) a.overflow_list[atomicAdd(a.overflow_count, 1)] = pid;
    }
    if (a.want_poplog) { /* retained plan: expose the shared-memory part of the heap to the host getters */
      __syncthreads();
      for (int i = tid; i < S.n_heap && i < SM::HCAP; i += MPLB_NT) {
        HeapEnt e; e.f = S.hf[i]; e.g = S.hg[i]; e.node = S.hn[i]; e.pad = 0;
        spill[i] = e;
      }
    }
  }
}

/* ---------------------------------------------------------------- get_succ for arbitrary states (parity artefact) */
template <int DIM, int ORD, int MAXU>
__global__ void __launch_bounds__(MPLB_NT) expand_trace_kernel(const DevCfg c, const mplb_waypoint *states, int n_states,
                                                               mplb_prim_trace *rows) {
  constexpr int NS = DIM * ORD;
  constexpr int NW = MPLB_NT / 32;
  using SM = PlanSmem<DIM, ORD, MAXU>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SM &S = *reinterpret_cast<SM *>(smem_raw);
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
    expand_b2_percontrol<DIM, ORD>(c, S, warp, lane, NW);
    __syncthreads();
    for (int i = tid; i < c.nU; i += MPLB_NT) {
      mplb_prim_trace r;
      int v = S.verdict[i];
      if (v == 5) v = (S.first[i] == 0x7fffffff) ? 3 : 2;
      r.verdict = v;
      r.n = (v == 2 || v == 3) ? S.nsamp[i] : 0;
      r.n_tested = (v == 3) ? S.cnt[i] : (v == 2 ? S.first[i] + 1 : 0);
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
