/* The replanning callbacks of mpl_test_node/src/map_replanner_node.cpp:107-241 (replanCallback, addCloudCallback,
 * clearCloudCallback, subtreeCallback) re-typed without ROS against this repo's header: an A* planner and an LPA* planner on
 * one shared MapUtil, an obstacle dropped across the trajectory through getMap / edit / setMap + updateBlockedNodes, a
 * replan, half of it cleared again + updateClearedNodes, a replan, the root moved one step (getSubStateSpace(1)) and a last
 * replan.  Prints what the node prints (costs, open / closed counts) in a fixed format; tests/test_cpp_shim.py compares it
 * with the oracle's answers for the same flow.  argv[1]: corridor.bin (dim, origin, res, start, goal, cells). */
#include <mpl_b200/map_planner.hpp>

#include <cstdio>
#include <fstream>

using namespace MPL;

int main(int argc, char **argv) {
  if (argc < 2) return 2;
  std::ifstream f(argv[1], std::ios::binary);
  int nd[2];
  double ori[2], res, st[2], gl[2];
  f.read((char *)nd, sizeof(nd)); f.read((char *)ori, sizeof(ori)); f.read((char *)&res, sizeof(res));
  f.read((char *)st, sizeof(st)); f.read((char *)gl, sizeof(gl));
  Tmap data((size_t)nd[0] * nd[1]);
  f.read((char *)data.data(), data.size());

  std::shared_ptr<OccMapUtil> map_util(new OccMapUtil);
  map_util->setMap(Vec2f(ori[0], ori[1]), Vec2i(nd[0], nd[1]), data, res);
  map_util->freeUnknown();

  Waypoint2D start, goal;
  start.pos = Vec2f(st[0], st[1]); start.vel = Vec2f::Zero(); start.acc = Vec2f::Zero(); start.jrk = Vec2f::Zero();
  start.use_pos = true; start.use_vel = true; start.use_acc = false; start.use_jrk = false; start.use_yaw = false;
  goal.pos = Vec2f(gl[0], gl[1]); goal.vel = Vec2f::Zero(); goal.acc = Vec2f::Zero(); goal.jrk = Vec2f::Zero();
  goal.control = start.control;

  vec_E<VecDf> U;
  const decimal_t du = 0.5;
  for (decimal_t dx = -0.5; dx <= 0.5; dx += du)
    for (decimal_t dy = -0.5; dy <= 0.5; dy += du) { VecDf u(2); u[0] = dx; u[1] = dy; U.push_back(u); }

  OccMapPlanner planner_(false), replan_planner_(false);
  for (OccMapPlanner *p : {&planner_, &replan_planner_}) {
    p->setMapUtil(map_util); p->setVmax(1.0); p->setAmax(1.0); p->setDt(1.0); p->setU(U);
  }
  replan_planner_.setLPAstar(true);

  auto replan = [&](const char *tag) { /* replanCallback */
    const bool a = planner_.plan(start, goal);
    const bool l = replan_planner_.plan(start, goal);
    std::printf("%s: astar ok %d cost %.10f closed %zu | lpastar ok %d cost %.10f expand %d open %zu closed %zu segs %zu\n", tag, (int)a,
                a ? planner_.getTrajCost() : -1.0, planner_.getCloseSet().size(), (int)l, l ? replan_planner_.getTrajCost() : -1.0,
                replan_planner_.getExpandedNum(), replan_planner_.getOpenSet().size(), replan_planner_.getCloseSet().size(),
                replan_planner_.getTraj().getPrimitives().size());
    std::printf("%s: linked %zu\n", tag, replan_planner_.getLinkedNodes().size()); /* visualizeGraph */
    return l;
  };
  if (!replan("first")) return 1;

  /* addCloudCallback: a 5 x 5 patch on the trajectory becomes occupied */
  auto ws = replan_planner_.getTraj().getWaypoints();
  const Vec2i c = map_util->floatToInt(ws[(int)(ws.size() * 0.45)].pos);
  vec_Vec2i new_obs;
  Tmap m = map_util->getMap();
  for (int nx = -2; nx <= 2; nx++)
    for (int ny = -2; ny <= 2; ny++) {
      const Vec2i pn(c(0) + nx, c(1) + ny);
      if (pn(0) < 0 || pn(0) >= nd[0] || pn(1) < 0 || pn(1) >= nd[1]) continue;
      signed char &v = m[pn(0) + nd[0] * pn(1)];
      if (v >= 0 && v < 100) { v = 100; new_obs.push_back(pn); }
    }
  map_util->setMap(map_util->getOrigin(), map_util->getDim(), m, map_util->getRes());
  if (replan_planner_.initialized()) replan_planner_.updateBlockedNodes(new_obs);
  std::printf("blocked %zu cells\n", new_obs.size());
  if (!replan("blocked")) return 1;

  /* clearCloudCallback: half of them are free again */
  vec_Vec2i new_clear(new_obs.begin(), new_obs.begin() + new_obs.size() / 2);
  for (const auto &pn : new_clear) m[pn(0) + nd[0] * pn(1)] = 0;
  map_util->setMap(map_util->getOrigin(), map_util->getDim(), m, map_util->getRes());
  if (replan_planner_.initialized()) replan_planner_.updateClearedNodes(new_clear);
  if (!replan("cleared")) return 1;

  /* subtreeCallback: the root moves to the next waypoint */
  replan_planner_.getSubStateSpace(1);
  ws = replan_planner_.getTraj().getWaypoints();
  if (ws.size() < 3) return 1;
  start = ws[1];
  replan("subtree");
  return 0;
}
