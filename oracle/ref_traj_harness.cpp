/*
 * ref_traj_harness.cpp — drives the REFERENCE'S OWN trajectory post-processing sources (TEST INFRASTRUCTURE).
 *
 * Compiles, from where they lie under /root/reference, mpl_traj_solver/traj_solver.h, src/mpl_traj_solver/poly_solver.cpp
 * and src/mpl_traj_solver/poly_traj.cpp, unmodified, against the stand-in Eigen of oracle/shim (dense matrices, block views
 * and PartialPivLU restated there).  Every line of TrajSolver / PolySolver / PolyTraj logic executed is the reference's;
 * the linear algebra underneath is the shim's.  Used by tests/test_oracle_trajsolver.py to pin oracle/poly_oracle.cpp.
 */
#include <mpl_traj_solver/traj_solver.h>

#include "mpl_oracle.h"

namespace {
template <int Dim>
int emit(const Trajectory<Dim> &traj, double *coefs) {
  const int nseg = (int)traj.segs.size();
  for (int i = 0; i < nseg; i++) {
    for (int a = 0; a <= Dim; a++) {
      const Vec6f c = a < Dim ? traj.segs[i].pr(a).coeff() : traj.segs[i].pr_yaw().coeff();
      for (int k = 0; k < 6; k++) coefs[((size_t)i * (Dim + 1) + a) * 6 + k] = c(k);
    }
  }
  return nseg;
}

template <int Dim>
int solve_wps(int control, int yaw_control, int n_wp, const orc_waypoint *wps, const double *dts, double *coefs) {
  vec_E<Waypoint<Dim>> ws(n_wp);
  for (int i = 0; i < n_wp; i++) {
    for (int a = 0; a < Dim; a++) { ws[i].pos(a) = wps[i].pos[a]; ws[i].vel(a) = wps[i].vel[a]; ws[i].acc(a) = wps[i].acc[a]; ws[i].jrk(a) = wps[i].jrk[a]; }
    ws[i].yaw = wps[i].yaw; ws[i].t = wps[i].t; ws[i].control = (Control::Control)wps[i].control;
  }
  TrajSolver<Dim> solver((Control::Control)control, (Control::Control)yaw_control); /* map_planner_node.cpp:224-227 */
  solver.setWaypoints(ws);
  solver.setDts(std::vector<decimal_t>(dts, dts + n_wp - 1));
  return emit<Dim>(solver.solve(), coefs);
}

template <int Dim>
int solve_path(int control, int n, const double *pts3, double v, double *coefs, double *dts_out) {
  vec_Vecf<Dim> path(n);
  for (int i = 0; i < n; i++) for (int a = 0; a < Dim; a++) path[i](a) = pts3[3 * i + a];
  TrajSolver<Dim> solver((Control::Control)control); /* MPL/test/test_traj_solver.cpp:29-34 */
  solver.setPath(path);
  solver.setV(v);
  const Trajectory<Dim> traj = solver.solve();
  const std::vector<decimal_t> dts = solver.getDts();
  for (size_t i = 0; i < dts.size(); i++) dts_out[i] = dts[i];
  return emit<Dim>(traj, coefs);
}
}  // namespace

extern "C" {
int ref_traj_solve(int dim, int control, int yaw_control, int n_wp, const orc_waypoint *wps, const double *dts, double *coefs) {
  return dim == 2 ? solve_wps<2>(control, yaw_control, n_wp, wps, dts, coefs) : solve_wps<3>(control, yaw_control, n_wp, wps, dts, coefs);
}
int ref_traj_solve_path(int dim, int control, int n, const double *pts3, double v, double *coefs, double *dts_out) {
  return dim == 2 ? solve_path<2>(control, n, pts3, v, coefs, dts_out) : solve_path<3>(control, n, pts3, v, coefs, dts_out);
}
}
