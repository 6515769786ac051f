/*
 * traj_solver.hpp — header-only C++ host layer: the reference's TrajSolver<Dim> on top of the C ABI
 * (mplb_traj_solve_batch, include/mplb.h), member for member as
 *   motion_primitive_library/include/mpl_traj_solver/traj_solver.h:12-162
 * so that MPL/test/test_traj_solver.cpp and the refinement step of mpl_test_node/src/map_planner_node.cpp:216-227 compile
 * unchanged.  The spline (PolySolver<Dim>::solve, src/mpl_traj_solver/poly_solver.cpp:23-221) is solved on the GPU; time
 * allocation (L-inf distance over v, traj_solver.h:122-131) stays on the host exactly as in the reference.  `solve_batch`
 * is the entry the GPU is there for: many waypoint lists in one launch.
 */
#ifndef MPL_B200_TRAJ_SOLVER_HPP
#define MPL_B200_TRAJ_SOLVER_HPP

#include "map_planner.hpp"

template <int Dim>
class TrajSolver {
 public:
  TrajSolver(Control::Control control, Control::Control yaw_control = Control::VEL, bool debug = false)
      : control_(control), yaw_control_(yaw_control), debug_(debug) {}

  void setWaypoints(const vec_E<Waypoint<Dim>> &ws) { /* traj_solver.h:39-43 */
    path_.resize(ws.size());
    for (size_t i = 0; i < ws.size(); i++) path_[i] = ws[i].pos;
    waypoints_ = ws;
  }
  void setV(decimal_t v) { v_ = v; }
  void setDts(const std::vector<decimal_t> &dts) { dts_ = dts; }
  void setPath(const vec_Vecf<Dim> &path) { /* traj_solver.h:54-70 */
    path_ = path;
    waypoints_.resize(path_.size());
    for (size_t i = 0; i < waypoints_.size(); i++) {
      waypoints_[i].pos = path[i];
      waypoints_[i].vel = Vecf<Dim>::Zero();
      waypoints_[i].acc = Vecf<Dim>::Zero();
      waypoints_[i].jrk = Vecf<Dim>::Zero();
      waypoints_[i].yaw = 0;
      waypoints_[i].control = Control::VEL;
    }
    if (!waypoints_.empty()) {
      waypoints_.front().control = control_;
      waypoints_.back().control = control_;
    }
  }

  /* traj_solver.h:73-109 */
  Trajectory<Dim> solve(bool verbose = false) {
    if (waypoints_.size() != dts_.size() + 1) dts_ = allocate_time(path_, v_);
    if (waypoints_.size() != dts_.size() + 1 || waypoints_.size() < 2) return Trajectory<Dim>();
    std::vector<Trajectory<Dim>> out;
    if (!solve_batch(control_, yaw_control_, {waypoints_}, {dts_}, out, verbose || debug_)) return Trajectory<Dim>();
    return out[0];
  }

  /* Batch form: trajectory i is (waypoints[i], dts[i]); results in `out`.  Returns false when the library call failed
   * (message printed when verbose). */
  static bool solve_batch(Control::Control control, Control::Control yaw_control, const std::vector<vec_E<Waypoint<Dim>>> &waypoints,
                          const std::vector<std::vector<decimal_t>> &dts, std::vector<Trajectory<Dim>> &out, bool verbose = false) {
    const int n = (int)waypoints.size();
    std::vector<int32_t> off(n + 1, 0), nseg(n > 0 ? n : 1, 0);
    std::vector<mplb_waypoint> wps;
    std::vector<double> d;
    for (int i = 0; i < n; i++) {
      off[i + 1] = off[i] + (int)waypoints[i].size();
      for (const auto &w : waypoints[i]) {
        mplb_waypoint r;
        std::memset(&r, 0, sizeof(r));
        for (int k = 0; k < Dim; k++) { r.pos[k] = w.pos(k); r.vel[k] = w.vel(k); r.acc[k] = w.acc(k); r.jrk[k] = w.jrk(k); }
        r.yaw = w.yaw; r.t = w.t; r.control = (int32_t)w.control;
        wps.push_back(r);
      }
      const size_t slots = waypoints[i].empty() ? 0 : waypoints[i].size() - 1;
      if (dts[i].size() != slots) return false;
      d.insert(d.end(), dts[i].begin(), dts[i].end());
    }
    std::vector<double> coefs((d.size() > 0 ? d.size() : 1) * (Dim + 1) * 6, 0.0);
    const int rc = mplb_traj_solve_batch(Dim, (int)control, (int)yaw_control, n, off.data(), wps.data(), d.data(), coefs.data(), nseg.data());
    if (rc != MPLB_OK) {
      if (verbose) std::printf(ANSI_COLOR_RED "TrajSolver: %s\n" ANSI_COLOR_RESET, mplb_last_error());
      return false;
    }
    out.clear();
    size_t so = 0;
    for (int i = 0; i < n; i++) {
      vec_E<Primitive<Dim>> prs;
      for (int s = 0; s < nseg[i]; s++) {
        const double *c = &coefs[(so + s) * (Dim + 1) * 6];
        prs.push_back(Primitive<Dim>(c, c + Dim * 6, dts[i][s], waypoints[i].front().control)); /* poly_traj.cpp:88 */
      }
      out.push_back(prs.empty() ? Trajectory<Dim>() : Trajectory<Dim>(prs));
      so += waypoints[i].empty() ? 0 : waypoints[i].size() - 1;
    }
    return true;
  }

  vec_Vecf<Dim> getPath() const { return path_; }
  vec_E<Waypoint<Dim>> getWaypoints() const { return waypoints_; }
  std::vector<decimal_t> getDts() const { return dts_; }

 private:
  std::vector<decimal_t> allocate_time(const vec_Vecf<Dim> &pts, decimal_t v) { /* traj_solver.h:122-131 */
    if (pts.size() < 2 || v <= 0) return std::vector<decimal_t>();
    std::vector<decimal_t> dts(pts.size() - 1);
    for (unsigned int i = 1; i < pts.size(); i++) {
      decimal_t d = 0;
      for (int k = 0; k < Dim; k++) {
        const decimal_t x = std::fabs(pts[i](k) - pts[i - 1](k));
        if (x > d) d = x;
      }
      dts[i - 1] = d / v;
    }
    return dts;
  }
  vec_Vecf<Dim> path_;
  vec_E<Waypoint<Dim>> waypoints_;
  std::vector<decimal_t> dts_;
  decimal_t v_{1};
  Control::Control control_, yaw_control_;
  bool debug_;
};
typedef TrajSolver<2> TrajSolver2D;
typedef TrajSolver<3> TrajSolver3D;
#endif
