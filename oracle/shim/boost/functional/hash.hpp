/* Stand-in for boost::hash / hash_combine (see oracle/shim/Eigen/Core).  waypoint.h:92-135 folds the lattice ints of a
 * Waypoint through hash_combine and treats two waypoints as EQUAL when the 64-bit results are equal, so the mixing
 * function is part of the reference's behaviour: a weak one merges distinct states.  Restated here is the combine that
 * Boost 1.56 - 1.80 select for a 64-bit std::size_t (the era and platforms of the reference: hash_combine_impl for
 * boost::uint64_t, a MurmurHash2-style mix); integers hash to themselves.  (The textbook formula
 * seed ^= v + 0x9e3779b9 + (seed << 6) + (seed >> 2), which those Boost versions use only for other size_t widths,
 * collides on real planner keys: on the levine-256 benchmark map it merged 6 pairs of distinct states in one search.
 * Boost >= 1.81 uses yet another mixer.  The oracle and the CUDA path compare the integer tuples themselves.) */
#ifndef MPLB_SHIM_BOOST_HASH
#define MPLB_SHIM_BOOST_HASH
#include <cstddef>
#include <cstdint>
#include <functional>
namespace boost {
inline std::size_t hash_value(int v) { return (std::size_t)v; }
inline std::size_t hash_value(std::size_t v) { return v; }
inline void hash_combine_impl(std::uint64_t &h, std::uint64_t k) {
  const std::uint64_t m = 0xc6a4a7935bd1e995ull;
  const int r = 47;
  k *= m;
  k ^= k >> r;
  k *= m;
  h ^= k;
  h *= m;
  h += 0xe6546b64; /* "completely arbitrary number, to prevent 0's from hashing to 0" */
}
template <class T>
inline void hash_combine(std::size_t &seed, const T &v) {
  std::uint64_t h = seed;
  hash_combine_impl(h, (std::uint64_t)hash_value(v));
  seed = (std::size_t)h;
}
template <class T>
struct hash {
  std::size_t operator()(const T &v) const { return hash_value(v); } /* ADL finds hash_value(const Waypoint<Dim>&) */
};
}  // namespace boost
#endif
