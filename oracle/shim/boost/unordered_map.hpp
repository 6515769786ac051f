/* Stand-in for boost::unordered_map (TEST INFRASTRUCTURE, see oracle/shim/Eigen/Core).
 *
 * Boost leaves the iteration order of its unordered containers unspecified (it depends on the Boost version, the bucket
 * count history and the hash), and two pieces of the reference's LPA* branch observe it: StateSpace::getSubStateSpace re-pushes
 * the open set in hm_ iteration order (state_space.h:190-199) and MapPlanner::getLinkedNodes fills the voxel -> edge lists in
 * hm_ iteration order (map_planner.cpp:125-158), which decides the order of updateNode calls and thereby the heap layout among
 * exact key ties.  This stand-in iterates in INSERTION order — a legal order for an unordered container, and the one the oracle
 * and the CUDA path define for those two loops.  Everything else (operator[], find, size, clear, copy)
 * behaves like any unordered_map; A* never iterates hm_ on a path that affects results. */
#ifndef MPLB_SHIM_BOOST_UNORDERED_MAP
#define MPLB_SHIM_BOOST_UNORDERED_MAP
#include <functional>
#include <list>
#include <unordered_map>
#include <utility>
namespace boost {
template <class K, class V, class H = std::hash<K>, class E = std::equal_to<K>>
class unordered_map {
  typedef std::list<std::pair<const K, V>> List;

 public:
  typedef typename List::iterator iterator;
  typedef typename List::const_iterator const_iterator;
  typedef std::pair<const K, V> value_type;
  unordered_map() {}
  unordered_map(const unordered_map &o) { *this = o; }
  unordered_map &operator=(const unordered_map &o) {
    if (this == &o) return *this;
    clear();
    for (const auto &kv : o.items_) { items_.push_back(kv); index_[kv.first] = --items_.end(); }
    return *this;
  }
  V &operator[](const K &k) {
    auto it = index_.find(k);
    if (it != index_.end()) return it->second->second;
    items_.push_back(value_type(k, V()));
    index_[k] = --items_.end();
    return items_.back().second;
  }
  iterator find(const K &k) { auto it = index_.find(k); return it == index_.end() ? items_.end() : it->second; }
  const_iterator find(const K &k) const { auto it = index_.find(k); return it == index_.end() ? items_.end() : const_iterator(it->second); }
  iterator begin() { return items_.begin(); }
  iterator end() { return items_.end(); }
  const_iterator begin() const { return items_.begin(); }
  const_iterator end() const { return items_.end(); }
  std::size_t size() const { return items_.size(); }
  bool empty() const { return items_.empty(); }
  void clear() { items_.clear(); index_.clear(); }

 private:
  List items_;
  std::unordered_map<K, iterator, H, E> index_;
};
}  // namespace boost
#endif
