/*
 * poly_oracle.cpp — CPU ORACLE for the trajectory post-processing row: TrajSolver / PolySolver.
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle/mpl_oracle.h): only tests/, smoke() and bench.py load it.
 *
 * A plain-array restatement of
 *   TrajSolver<Dim>::solve            MPL/include/mpl_traj_solver/traj_solver.h:73-109
 *   PolySolver<Dim>::solve            MPL/src/mpl_traj_solver/poly_solver.cpp:23-221
 *   PolyTraj<Dim>::toPrimitives       MPL/src/mpl_traj_solver/poly_traj.cpp:75-92
 * in the reference's operation order: dense A, Q, M of the full (segments*N)^2 size, X = A.partialPivLu().solve(M),
 * R = X^T * Q * X (left to right), Dp = -Rpp.partialPivLu().solve(Rpf * Df), d = M * D, one N x N solve per segment,
 * coeff(k) = p(k, j) * k! reversed into six entries (highest order first).
 *
 * Third-party arithmetic restated (Eigen is absent from the image): `partialPivLu().solve()` is taken as the textbook
 * algorithm Eigen's unblocked kernel implements — column k: first row of maximal |a(r,k)|, r >= k; row swap; the
 * sub-column divided by the pivot; rank-1 update — followed by the permuted forward substitution (unit lower, contributions
 * subtracted in ascending column order) and the back substitution (contributions subtracted in descending column order,
 * then a division by the diagonal) — the order of Eigen's column-major triangular solver inside one panel, i.e. exactly
 * its order for the N x N (N <= 6) per-segment systems; matrix products as plain (i, j, k) loops with k ascending.  Real Eigen blocks and vectorises these loops for sizes > 16, which changes results
 * by rounding only (~1e-15 relative); the tolerance the tests use against this oracle is written there.
 *
 * Parity pin: the reference's own traj_solver.h / poly_solver.cpp / poly_traj.cpp compiled against oracle/shim
 * (ref_harness.cpp: ref_traj_solve) — tests/test_oracle_trajsolver.py compares the two on the MPL/test/test_traj_solver.cpp
 * path, the map_planner_node.cpp:216-227 refinement flow and seeded random inputs.  The reference publishes no numbers
 * for this component (test_traj_solver.cpp draws a picture).
 */
#include <cmath>
#include <cstring>
#include <utility>
#include <vector>

#include "mpl_oracle.h"

namespace {

struct Mat { /* dense column-major, like Eigen's default */
  int r = 0, c = 0;
  std::vector<double> d;
  Mat() {}
  Mat(int r_, int c_) : r(r_), c(c_), d((size_t)r_ * c_, 0.0) {}
  double &operator()(int i, int j) { return d[(size_t)i + (size_t)j * r]; }
  double operator()(int i, int j) const { return d[(size_t)i + (size_t)j * r]; }
};

double power(double t, int n) { /* math.h:197-203 */
  double tn = 1;
  while (n > 0) { tn *= t; n--; }
  return tn;
}

int factorial(int n) { /* math.h:187-194 */
  int nf = 1;
  while (n > 0) { nf *= n; n--; }
  return nf;
}

Mat mul(const Mat &a, const Mat &b) {
  Mat m(a.r, b.c);
  for (int i = 0; i < a.r; i++)
    for (int j = 0; j < b.c; j++) {
      double acc = 0;
      for (int k = 0; k < a.c; k++) acc += a(i, k) * b(k, j);
      m(i, j) = acc;
    }
  return m;
}

Mat transpose(const Mat &a) {
  Mat m(a.c, a.r);
  for (int i = 0; i < a.r; i++)
    for (int j = 0; j < a.c; j++) m(j, i) = a(i, j);
  return m;
}

Mat block(const Mat &a, int r0, int c0, int nr, int nc) {
  Mat m(nr, nc);
  for (int i = 0; i < nr; i++)
    for (int j = 0; j < nc; j++) m(i, j) = a(r0 + i, c0 + j);
  return m;
}

/* PartialPivLU::solve as described in the header comment */
Mat lu_solve(Mat a, Mat b) {
  const int n = a.r;
  std::vector<int> piv(n);
  for (int k = 0; k < n; k++) {
    int p = k;
    double best = std::fabs(a(k, k));
    for (int r = k + 1; r < n; r++)
      if (std::fabs(a(r, k)) > best) { best = std::fabs(a(r, k)); p = r; }
    piv[k] = p;
    if (best != 0) {
      if (p != k)
        for (int c = 0; c < n; c++) std::swap(a(k, c), a(p, c));
      for (int r = k + 1; r < n; r++) a(r, k) /= a(k, k);
    }
    for (int c = k + 1; c < n; c++)
      for (int r = k + 1; r < n; r++) a(r, c) -= a(r, k) * a(k, c);
  }
  for (int k = 0; k < n; k++)
    if (piv[k] != k)
      for (int c = 0; c < b.c; c++) std::swap(b(k, c), b(piv[k], c));
  for (int c = 0; c < b.c; c++) {
    for (int i = 0; i < n; i++) {
      double acc = b(i, c);
      for (int j = 0; j < i; j++) acc -= a(i, j) * b(j, c);
      b(i, c) = acc;
    }
    for (int i = n - 1; i >= 0; i--) {
      double acc = b(i, c);
      for (int j = n - 1; j > i; j--) acc -= a(i, j) * b(j, c); /* column-oriented substitution visits j downwards */
      b(i, c) = acc / a(i, i);
    }
  }
  return b;
}

struct Wp { /* the fields PolySolver reads */
  const double *pos, *vel, *acc, *jrk;
  bool use_pos, use_vel, use_acc, use_jrk;
};

/* PolySolver<Dim>(smooth_derivative_order, minimize_derivative)::solve, poly_solver.cpp:23-221.
 * out: per segment, per axis, six Primitive coefficients highest order first (poly_traj.cpp:80-86). */
bool poly_solve(int Dim, unsigned N_, unsigned R_, const std::vector<Wp> &waypoints, const double *dts, double *out,
                int out_axis_stride, int out_seg_stride) {
  const unsigned num_waypoints = (unsigned)waypoints.size();
  if (num_waypoints < 2) return false; /* ps:31 */
  const unsigned num_segments = num_waypoints - 1;
  const int S = (int)(num_segments * N_);
  Mat A(S, S), Q(S, S);
  for (unsigned i = 0; i < num_segments; i++) { /* ps:40-70 */
    const double seg_time = dts[i];
    for (unsigned n = 0; n < N_; n++) {
      if (n < N_ / 2) {
        int val = 1;
        for (unsigned m = 0; m < n; m++) val *= (n - m);
        A(i * N_ + n, i * N_ + n) = val;
      }
      for (unsigned r = 0; r < N_ / 2; r++) {
        if (r <= n) {
          int val = 1;
          for (unsigned m = 0; m < r; m++) val *= (n - m);
          A(i * N_ + N_ / 2 + r, i * N_ + n) = val * power(seg_time, n - r);
        }
      }
      for (unsigned r = 0; r < N_; r++) {
        if (r >= R_ && n >= R_) {
          int val = 1;
          for (unsigned m = 0; m < R_; m++) val *= (r - m) * (n - m);
          Q(i * N_ + r, i * N_ + n) = val * power(seg_time, r + n - 2 * R_ + 1) / (r + n - 2 * R_ + 1);
        }
      }
    }
  }
  const unsigned smooth = N_ / 2 - 1; /* ps:75 */
  const unsigned num_total = num_waypoints * N_ / 2;
  unsigned num_fixed = 0;
  for (const auto &it : waypoints) { /* ps:78-83 */
    if (it.use_pos) num_fixed++;
    if (it.use_vel && smooth >= 1) num_fixed++;
    if (it.use_acc && smooth >= 2) num_fixed++;
    if (it.use_jrk && smooth >= 3) num_fixed++;
  }
  const unsigned num_free = num_total - num_fixed;
  std::vector<std::pair<unsigned, unsigned>> table; /* ps:90-160 */
  unsigned raw_cnt = 0, fix_cnt = 0, free_cnt = 0, id = 0;
  for (const auto &it : waypoints) {
    const bool interior = id > 0 && id < num_waypoints - 1;
    const bool use[4] = {it.use_pos, it.use_vel, it.use_acc, it.use_jrk};
    for (unsigned k = 0; k < 4; k++) {
      if (smooth < k) continue;
      const unsigned nid = use[k] ? fix_cnt : num_fixed + free_cnt;
      table.push_back(std::make_pair(raw_cnt, nid));
      if (interior) table.push_back(std::make_pair(raw_cnt + N_ / 2, nid));
      raw_cnt++;
      if (use[k]) fix_cnt++; else free_cnt++;
    }
    if (interior) raw_cnt += N_ / 2;
    id++;
  }
  Mat M(S, (int)num_total); /* ps:170-171 */
  for (const auto &it : table) M((int)it.first, (int)it.second) = 1;
  Mat A_inv_M = lu_solve(A, M);                                /* ps:174 */
  Mat R = mul(mul(transpose(A_inv_M), Q), A_inv_M);           /* ps:175 */
  Mat Rpp = block(R, num_fixed, num_fixed, num_free, num_free); /* ps:183-186 */
  Mat Rpf = block(R, num_fixed, 0, num_free, num_fixed);
  Mat D((int)num_total, Dim);
  for (const auto &it : table) { /* ps:189-203 */
    if (it.second < num_fixed) {
      const int wid = (int)std::floor((it.first + N_ / 2) / N_);
      const int derivative = it.first % (N_ / 2);
      const double *src = derivative == 0 ? waypoints[wid].pos : derivative == 1 ? waypoints[wid].vel
                          : derivative == 2 ? waypoints[wid].acc : waypoints[wid].jrk;
      for (int a = 0; a < Dim; a++) D((int)it.second, a) = src[a];
    }
  }
  if (num_waypoints > 2 && num_free > 0) { /* ps:209-213 */
    Mat Df = block(D, 0, 0, num_fixed, Dim);
    Mat Dp = lu_solve(Rpp, mul(Rpf, Df));
    for (unsigned i = 0; i < num_free; i++)
      for (int a = 0; a < Dim; a++) D(num_fixed + i, a) = -Dp(i, a);
  }
  Mat d = mul(M, D); /* ps:215 */
  for (unsigned i = 0; i < num_segments; i++) { /* ps:217-223 + poly_traj.cpp:75-92 */
    Mat p = lu_solve(block(A, i * N_, i * N_, N_, N_), block(d, i * N_, 0, N_, Dim));
    for (int j = 0; j < Dim; j++) {
      double coeff[6] = {0, 0, 0, 0, 0, 0};
      for (unsigned k = 0; k < N_; k++) coeff[k] = p(k, j) * factorial(k);
      double *o = out + (size_t)i * out_seg_stride + (size_t)j * out_axis_stride;
      for (int k = 0; k < 6; k++) o[k] = coeff[5 - k]; /* Vec6f::reverse() */
    }
  }
  return true;
}

bool solver_orders(int control, unsigned *N, unsigned *R) { /* traj_solver.h:21-27 (pos) / :31-36 (yaw) */
  const int c = control & 0xf;
  if (c == 1) { *N = 2; *R = 1; return true; }
  if (c == 3) { *N = 4; *R = 2; return true; }
  if (c == 7) { *N = 6; *R = 3; return true; }
  return false; /* SNP: no solver ("only works up to third order") */
}

}  // namespace

extern "C" {

/* TrajSolver<dim>(control, yaw_control).setWaypoints(wps); setDts(dts); solve()  (traj_solver.h:73-109).
 * coefs: (n_wp - 1) segments x (dim + 1) rows (axes, then yaw) x 6 coefficients, highest order first — what
 * Trajectory::segs[i].prs_[a].coeff() / pr_yaw_.coeff() hold.  Returns the number of segments, 0 when the solver is not
 * initialised (control SNP, yaw control other than VEL/ACC/JRK) or fewer than two waypoints are given. */
int orc_traj_solve(int dim, int control, int yaw_control, int n_wp, const orc_waypoint *wps, const double *dts, double *coefs) {
  unsigned N, R, Ny, Ry;
  if (!solver_orders(control, &N, &R)) return 0;
  if (!(yaw_control == 1 || yaw_control == 3 || yaw_control == 7) || !solver_orders(yaw_control, &Ny, &Ry)) return 0;
  if (n_wp < 2) return 0;
  std::vector<Wp> ws(n_wp), ys(n_wp);
  std::vector<double> yaw_pos(n_wp), zero(1, 0.0);
  for (int i = 0; i < n_wp; i++) {
    const int c = wps[i].control;
    ws[i] = Wp{wps[i].pos, wps[i].vel, wps[i].acc, wps[i].jrk, (c & 1) != 0, (c & 2) != 0, (c & 4) != 0, (c & 8) != 0};
    yaw_pos[i] = wps[i].yaw; /* traj_solver.h:88-95: Waypoint<1>(Control::VEL), pos = yaw, derivatives 0 */
    const int yc = (i == 0 || i == n_wp - 1) ? yaw_control : 1;
    ys[i] = Wp{&yaw_pos[i], zero.data(), zero.data(), zero.data(), (yc & 1) != 0, (yc & 2) != 0, (yc & 4) != 0, (yc & 8) != 0};
  }
  const int seg_stride = (dim + 1) * 6;
  poly_solve(dim, N, R, ws, dts, coefs, 6, seg_stride);
  poly_solve(1, Ny, Ry, ys, dts, coefs + (size_t)dim * 6, 6, seg_stride);
  return n_wp - 1;
}

/* TrajSolver::allocate_time (traj_solver.h:122-131): L-inf distance of consecutive positions over v. Returns n - 1 or 0. */
int orc_traj_allocate_time(int dim, int n, const double *pts3, double v, double *dts) {
  if (n < 2 || v <= 0) return 0;
  for (int i = 1; i < n; i++) {
    double m = 0;
    for (int a = 0; a < dim; a++) {
      const double x = std::fabs(pts3[3 * i + a] - pts3[3 * (i - 1) + a]);
      if (x > m) m = x;
    }
    dts[i - 1] = m / v;
  }
  return n - 1;
}
}
