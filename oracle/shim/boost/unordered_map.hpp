/* Stand-in: boost::unordered_map -> std::unordered_map (iteration order is not part of the planner's contract). */
#ifndef MPLB_SHIM_BOOST_UNORDERED_MAP
#define MPLB_SHIM_BOOST_UNORDERED_MAP
#include <unordered_map>
namespace boost {
template <class K, class V, class H = std::hash<K>, class E = std::equal_to<K>>
using unordered_map = std::unordered_map<K, V, H, E>;
}
#endif
