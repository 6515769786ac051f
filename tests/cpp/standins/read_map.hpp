/* Stand-in for MPL/test/read_map.hpp (which needs yaml-cpp, absent from this image): the same MapReader<Ti, Tf> members,
 * reading the little binary dump the tests write from tests/golden/maps/corridor.npz
 * (int32 dim[2], double origin[2], double res, double start[2], double goal[2], int8 data[]) instead of the YAML file.
 * Used only to compile the reference's own MPL/test/*.cpp UNMODIFIED against include/compat (tests/test_reference_tests.py). */
#pragma once
#include <cstdio>
#include <string>
#include <vector>

template <class Ti, class Tf>
class MapReader {
 public:
  MapReader(const std::string &file, bool verbose = false) {
    (void)verbose;
    FILE *f = std::fopen(file.c_str(), "rb");
    if (!f) return;
    int d[2];
    double o[2], s[2], g[2];
    if (std::fread(d, sizeof(int), 2, f) == 2 && std::fread(o, sizeof(double), 2, f) == 2 && std::fread(&resolution_, sizeof(double), 1, f) == 1 &&
        std::fread(s, sizeof(double), 2, f) == 2 && std::fread(g, sizeof(double), 2, f) == 2) {
      for (int i = 0; i < 2; i++) { dim_(i) = d[i]; origin_(i) = o[i]; start_(i) = s[i]; goal_(i) = g[i]; }
      data_.resize((size_t)d[0] * d[1]);
      exist_ = std::fread(data_.data(), 1, data_.size(), f) == data_.size();
    }
    std::fclose(f);
  }
  bool exist() { return exist_; }
  Tf origin() { return origin_; }
  Ti dim() { return dim_; }
  double start(int i) { return start_(i); }
  double goal(int i) { return goal_(i); }
  double resolution() { return resolution_; }
  std::vector<signed char> data() { return data_; }

 private:
  Tf start_, goal_, origin_;
  Ti dim_;
  double resolution_ = 0;
  std::vector<signed char> data_;
  bool exist_ = false;
};
