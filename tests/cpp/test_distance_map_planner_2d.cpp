/*
 * C++ host-side test of the cost-shaping members of include/mpl_b200/map_planner.hpp, following the flow of the
 * reference's motion_primitive_library/test/test_distance_map_planner_2d.cpp:46-93: plain plan, tunnel around its
 * trajectory (setSearchRadius/setSearchRegion), potential map (setPotentialRadius/Weight, updatePotentialMap), plan
 * again, then iterativePlan.  Map input: the flat binary of test_planner_2d.cpp.  The reference publishes no numbers
 * for this flow; the pytest wrapper compares the printed values with the oracle's.
 */
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
  start.use_pos = true; start.use_vel = true; start.use_acc = false; start.use_jrk = false; start.use_yaw = false;
  goal.pos = Vec2f(gl[0], gl[1]);
  goal.control = start.control;

  decimal_t u = 0.5, du = u;
  vec_E<VecDf> U;
  for (decimal_t dx = -u; dx <= u; dx += du)
    for (decimal_t dy = -u; dy <= u; dy += du) U.push_back(VecDf{dx, dy});

  auto fresh = [&]() {
    std::unique_ptr<MPL::OccMapPlanner> pl(new MPL::OccMapPlanner(false));
    pl->setMapUtil(map_util);
    pl->setVmax(1.0);
    pl->setAmax(1.0);
    pl->setDt(1.0);
    pl->setU(U);
    return pl;
  };
  auto planner = fresh();
  if (!planner->plan(start, goal)) { std::printf("plain plan failed\n"); return 1; }
  std::printf("MPL Planner expanded states: %zu\n", planner->getCloseSet().size());
  const auto traj = planner->getTraj();

  vec_Vec2f path;
  for (const auto &w : traj.getWaypoints()) path.push_back(w.pos);
  planner = fresh();
  planner->setEpsilon(1.0);
  planner->setSearchRadius(Vec2f(0.5, 0.5));
  planner->setSearchRegion(path);
  planner->setPotentialRadius(Vec2f(1.0, 1.0));
  planner->setPotentialWeight(0.5);
  planner->setGradientWeight(0);
  planner->updatePotentialMap(start.pos);
  if (!planner->plan(start, goal)) { std::printf("distance plan failed\n"); return 1; }
  const auto traj_dist = planner->getTraj();
  std::printf("MPL Distance Planner expanded states: %zu\n", planner->getCloseSet().size());
  std::printf("distance: cost %.10f pops %d segs %zu region %zu\n", planner->getTrajCost(), planner->getExpandedNum(),
              traj_dist.getPrimitives().size(), planner->getSearchRegion().size());

  bool ok = planner->iterativePlan(start, goal, traj_dist, 3);
  std::printf("iterative: ok %d cost %.10f segs %zu\n", ok ? 1 : 0, planner->getTrajCost(), planner->getTraj().getPrimitives().size());
  return 0;
}
