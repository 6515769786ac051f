/*
 * C++ host-side test of the yaw controls through include/mpl_b200/map_planner.hpp, following the reference's
 * motion_primitive_library/test/test_planner_2d_with_yaw.cpp:28-69 (ACC control with use_yaw, start yaw pi/2,
 * U = {-0.5,0,0.5}^2 x {-0.5,0,0.5}, yaw_max 0.7).  Map input: the flat binary of test_planner_2d.cpp.
 * The reference publishes no numbers for this test; the pytest wrapper compares with the oracle's answer.
 */
#include <cmath>
#include <cstdio>
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
  start.yaw = M_PI / 2;
  start.use_pos = true; start.use_vel = true; start.use_acc = false; start.use_jrk = false; start.use_yaw = true;
  goal.pos = Vec2f(gl[0], gl[1]);
  goal.yaw = 0;
  goal.control = start.control;

  decimal_t u_yaw = 0.5, u = 0.5, du = u;
  vec_E<VecDf> U;
  for (decimal_t dx = -u; dx <= u; dx += du)
    for (decimal_t dy = -u; dy <= u; dy += du)
      for (decimal_t dyaw = -u_yaw; dyaw <= u_yaw; dyaw += u_yaw) U.push_back(VecDf{dx, dy, dyaw});

  decimal_t yaw_max = 0.7;
  std::unique_ptr<MPL::OccMapPlanner> planner(new MPL::OccMapPlanner(true));
  planner->setMapUtil(map_util);
  planner->setVmax(1.0);
  planner->setAmax(1.0);
  planner->setYawmax(yaw_max);
  planner->setDt(1.0);
  planner->setU(U);

  bool valid = planner->plan(start, goal);
  std::printf("MPL Planner expanded states: %zu\n", planner->getCloseSet().size());
  if (!valid) { std::printf("plan failed\n"); return 1; }
  auto traj = planner->getTraj();
  const auto ws = traj.getWaypoints();
  std::printf("yaw: cost %.10f pops %d segs %zu first yaw %.6f last yaw %.6f\n", planner->getTrajCost(), planner->getExpandedNum(),
              traj.getPrimitives().size(), ws.front().yaw, ws.back().yaw);
  return 0;
}
