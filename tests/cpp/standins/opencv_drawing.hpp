/* Stand-in for MPL/test/opencv_drawing.hpp (which needs OpenCV, absent from this image): the drawing calls of the
 * reference's tests become no-ops; what they would have drawn is counted and printed by save()/show() so that the run
 * still says something about the planner's getters.  Used only to compile MPL/test/*.cpp UNMODIFIED. */
#pragma once
#include <cstdio>
#include <memory>
#include <string>

struct StandInColor { int b, g, r; };
static const StandInColor blue{255, 0, 0}, red{0, 0, 255}, green{0, 255, 0}, cyan{255, 255, 0}, magenta{255, 0, 255}, black{0, 0, 0},
    white{255, 255, 255}, grey{100, 100, 100}, yellow{0, 255, 255};

class OpenCVDrawing {
 public:
  OpenCVDrawing(std::shared_ptr<MPL::OccMapUtil> map_util) : map_util_(map_util) {}
  void drawPoints(const vec_Vec2f &pts, StandInColor, int = 1) { n_points_ += pts.size(); }
  void drawCircle(const Vec2f &, StandInColor, int, int = 1) { n_circles_++; }
  void drawTraj(const Trajectory2D &traj, StandInColor, int = 1, int = 200) { n_segs_ += traj.getPrimitives().size(); }
  void drawText(std::string, const Vec2i &, double, StandInColor) {}
  void drawLineStrip(const vec_E<vec_Vec2f> &trias, StandInColor, int = 1) { n_points_ += trias.size(); }
  void drawPotential(StandInColor, StandInColor) {}
  void show(std::string name) { save(name); }
  void save(std::string name) { std::printf("[stand-in drawing] %s: %zu points, %zu circles, %zu trajectory segments\n", name.c_str(), n_points_, n_circles_, n_segs_); }

 private:
  std::shared_ptr<MPL::OccMapUtil> map_util_;
  size_t n_points_ = 0, n_circles_ = 0, n_segs_ = 0;
};
