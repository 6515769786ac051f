/*
 * C++ host-side test over include/mpl_b200/map_planner.hpp, shaped like the reference's own
 * motion_primitive_library/test/test_planner_2d.cpp:7-68 (load the corridor map, freeUnknown, ACC-control start/goal,
 * U = {-0.5, 0, 0.5}^2 built with the same floating-point loop, setVmax/setAmax/setDt/setU, plan, print).
 * The map comes from a flat binary written by the pytest wrapper (the reference reads YAML through yaml-cpp,
 * which this image does not have): int32 dim[2], double origin[2], double res, double start[2], double goal[2], int8 data[].
 * Expected output (MPL/README.md:195-202): expanded states 615, T = 35, J(VEL) = 36.75, J(ACC) = 1.5.
 */
#include <cstdio>
#include <cstring>
#include <fstream>
#include <memory>

#include "mpl_b200/map_planner.hpp"

int main(int argc, char **argv) {
  if (argc != 2) { std::printf("Input map required!\n"); return -1; }
  std::ifstream f(argv[1], std::ios::binary);
  if (!f) { std::printf("Cannot find input file [%s]!\n", argv[1]); return -1; }
  int32_t dim[2]; double origin[2], res, st[2], gl[2];
  f.read((char *)dim, sizeof(dim)); f.read((char *)origin, sizeof(origin)); f.read((char *)&res, sizeof(res));
  f.read((char *)st, sizeof(st)); f.read((char *)gl, sizeof(gl));
  MPL::Tmap data((size_t)dim[0] * dim[1]);
  f.read((char *)data.data(), data.size());

  std::shared_ptr<MPL::OccMapUtil> map_util(new MPL::OccMapUtil);
  map_util->setMap(Vec2f(origin[0], origin[1]), Vec2i(dim[0], dim[1]), data, res);
  map_util->freeUnknown();

  Waypoint2D start, goal;
  start.pos = Vec2f(st[0], st[1]);
  start.use_pos = true; start.use_vel = true; start.use_acc = false; start.use_jrk = false; start.use_yaw = false;
  goal.pos = Vec2f(gl[0], gl[1]);
  goal.control = start.control;

  decimal_t u = 0.5, du = u;
  vec_E<VecDf> U;
  for (decimal_t dx = -u; dx <= u; dx += du)
    for (decimal_t dy = -u; dy <= u; dy += du) U.push_back(VecDf{dx, dy});

  std::unique_ptr<MPL::OccMapPlanner> planner(new MPL::OccMapPlanner(true));
  planner->setMapUtil(map_util);
  planner->setVmax(1.0);
  planner->setAmax(1.0);
  planner->setDt(1.0);
  planner->setU(U);

  bool valid = planner->plan(start, goal);
  std::printf("MPL Planner expanded states: %zu\n", planner->getCloseSet().size());
  if (!valid) { std::printf("plan failed\n"); return 1; }
  auto traj = planner->getTraj();
  std::printf("Total time T: %f\n", traj.getTotalTime());
  std::printf("Total J:  J(VEL) = %f, J(ACC) = %f\n", traj.J(Control::VEL), traj.J(Control::ACC));
  std::printf("cost: %f open: %zu expanded: %zu waypoints: %zu\n", planner->getTrajCost(), planner->getOpenSet().size(),
              planner->getExpandedNodes().size(), traj.getWaypoints().size());
  std::printf("edges: %zu valid: %zu cloud: %zu ray: %zu\n", planner->getExpandedEdges().size(), planner->getValidPrimitives().size(),
              map_util->getCloud().size(), map_util->rayTrace(start.pos, goal.pos).size());
  planner->reset();
  std::printf("initialized after reset: %d\n", planner->initialized() ? 1 : 0);
  return 0;
}
