/*
 * Stand-in for boost::heap::d_ary_heap<T, mutable_<true>, arity<D>, compare<Cmp>> (see oracle/shim/Eigen/Core), restated
 * from Boost.Heap's published algorithm (boost/heap/d_ary_heap.hpp, 1.6x-1.8x):
 *   push      append, then sift up while cmp(parent, child)            (parent has strictly lower priority)
 *   pop       move the last element to the root, then sift down: pick the best child with std::max_element under
 *             cmp (the FIRST of equally good children) and move down unless cmp(child, element)
 *   increase  sift up                      erase   sift up unconditionally to the root, then pop
 *   iteration in internal array order      handles stay valid across moves (the mutable wrapper stores list nodes)
 * The same rules are restated a second time inside oracle/mpl_oracle.cpp; this copy exists so that the reference's own
 * graph_search.h can be compiled and run.
 */
#ifndef MPLB_SHIM_BOOST_D_ARY_HEAP
#define MPLB_SHIM_BOOST_D_ARY_HEAP
#include <algorithm>
#include <cstddef>
#include <list>
#include <vector>
namespace boost {
namespace heap {
template <bool B> struct mutable_ {};
template <unsigned D> struct arity { static const unsigned value = D; };
template <class C> struct compare { typedef C type; };

template <class T, class M, class A, class CmpOpt>
class d_ary_heap {
  struct Node { T value; std::size_t index; };
  typedef std::list<Node> List;
  typedef typename CmpOpt::type Cmp;
  static const unsigned D = A::value;

 public:
  typedef T value_type;
  class handle_type {
   public:
    handle_type() {}
    T &operator*() const { return it_->value; }
   private:
    friend class d_ary_heap;
    explicit handle_type(typename List::iterator it) : it_(it) {}
    typename List::iterator it_;
  };
  class const_iterator {
   public:
    const_iterator(const std::vector<typename List::iterator> *q, std::size_t i) : q_(q), i_(i) {}
    const T &operator*() const { return (*q_)[i_]->value; }
    const_iterator &operator++() { ++i_; return *this; }
    bool operator!=(const const_iterator &o) const { return i_ != o.i_; }
   private:
    const std::vector<typename List::iterator> *q_;
    std::size_t i_;
  };
  typedef const_iterator iterator;

  d_ary_heap() {}
  d_ary_heap(const d_ary_heap &o) { *this = o; }
  d_ary_heap &operator=(const d_ary_heap &o) {
    if (this == &o) return *this;
    clear();
    for (std::size_t i = 0; i < o.q_.size(); i++) { nodes_.push_back(Node{o.q_[i]->value, i}); q_.push_back(--nodes_.end()); }
    return *this;
  }
  bool empty() const { return q_.empty(); }
  std::size_t size() const { return q_.size(); }
  void clear() { q_.clear(); nodes_.clear(); }
  const T &top() const { return q_.front()->value; }
  handle_type push(const T &v) {
    nodes_.push_back(Node{v, q_.size()});
    typename List::iterator it = --nodes_.end();
    q_.push_back(it);
    siftup(q_.size() - 1, false);
    return handle_type(it);
  }
  void pop() {
    typename List::iterator victim = q_.front();
    swap_pos(0, q_.size() - 1);
    q_.pop_back();
    nodes_.erase(victim);
    if (!q_.empty()) siftdown(0);
  }
  void increase(handle_type h) { siftup(h.it_->index, false); }
  void decrease(handle_type h) { siftdown(h.it_->index); }
  void update(handle_type h) {
    std::size_t i = h.it_->index;
    if (i != 0 && cmp_(q_[(i - 1) / D]->value, q_[i]->value)) siftup(i, false);
    else siftdown(i);
  }
  void erase(handle_type h) { siftup(h.it_->index, true); pop(); }
  const_iterator begin() const { return const_iterator(&q_, 0); }
  const_iterator end() const { return const_iterator(&q_, q_.size()); }

 private:
  void swap_pos(std::size_t a, std::size_t b) {
    std::swap(q_[a], q_[b]);
    q_[a]->index = a;
    q_[b]->index = b;
  }
  void siftup(std::size_t index, bool force) {
    while (index != 0) {
      std::size_t parent = (index - 1) / D;
      if (force || cmp_(q_[parent]->value, q_[index]->value)) { swap_pos(parent, index); index = parent; }
      else return;
    }
  }
  void siftdown(std::size_t index) {
    while (index * D + 1 < q_.size()) { /* not a leaf */
      const std::size_t first = index * D + 1, last = std::min(first + D, q_.size());
      std::size_t best = first;
      for (std::size_t c = first + 1; c < last; c++)
        if (cmp_(q_[best]->value, q_[c]->value)) best = c; /* std::max_element: first of the equally good children */
      if (!cmp_(q_[best]->value, q_[index]->value)) { swap_pos(index, best); index = best; }
      else return;
    }
  }
  List nodes_;
  std::vector<typename List::iterator> q_;
  Cmp cmp_;
};
}  // namespace heap
}  // namespace boost
#endif
