// Minimal stand-in for <boost/dynamic_bitset.hpp> (Boost is not installed in this image).
// TEST INFRASTRUCTURE ONLY: lets the UNMODIFIED reference sources under /root/reference/src compile
// into oracle/_ref/. Written from the documented Boost interface; implements only the members
// the reference's pairsnp.hpp touches (ctor(size), size, operator[], &, |=, count, flip,
// find_first, find_next, npos). Pure bit semantics, so results are identical to real Boost.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>
namespace boost {
template <typename Block = unsigned long, typename Alloc = void>
class dynamic_bitset {
 public:
  typedef std::size_t size_type;
  static const size_type npos = static_cast<size_type>(-1);
  class reference {
   public:
    reference(uint64_t &w, unsigned b) : w_(w), b_(b) {}
    reference &operator=(int v) {
      if (v) w_ |= (uint64_t(1) << b_); else w_ &= ~(uint64_t(1) << b_);
      return *this;
    }
    operator bool() const { return (w_ >> b_) & 1u; }
   private:
    uint64_t &w_;
    unsigned b_;
  };
  dynamic_bitset() : n_(0) {}
  explicit dynamic_bitset(size_type n) : n_(n), w_((n + 63) / 64, 0) {}
  size_type size() const { return n_; }
  reference operator[](size_type i) { return reference(w_[i >> 6], unsigned(i & 63)); }
  bool operator[](size_type i) const { return (w_[i >> 6] >> (i & 63)) & 1u; }
  dynamic_bitset &operator|=(const dynamic_bitset &o) {
    for (size_type k = 0; k < w_.size(); ++k) w_[k] |= o.w_[k];
    return *this;
  }
  dynamic_bitset &operator&=(const dynamic_bitset &o) {
    for (size_type k = 0; k < w_.size(); ++k) w_[k] &= o.w_[k];
    return *this;
  }
  size_type count() const {
    size_type c = 0;
    for (size_type k = 0; k < w_.size(); ++k) c += size_type(__builtin_popcountll(w_[k]));
    return c;
  }
  dynamic_bitset &flip() {
    for (size_type k = 0; k < w_.size(); ++k) w_[k] = ~w_[k];
    trim();
    return *this;
  }
  size_type find_first() const { return scan(0); }
  size_type find_next(size_type pos) const {
    if (pos == npos || pos + 1 >= n_) return npos;
    return scan(pos + 1);
  }
 private:
  void trim() {
    if (n_ & 63) w_.back() &= (uint64_t(1) << (n_ & 63)) - 1;
  }
  size_type scan(size_type from) const {
    if (from >= n_) return npos;
    size_type k = from >> 6;
    uint64_t cur = w_[k] & (~uint64_t(0) << (from & 63));
    for (;;) {
      if (cur) {
        size_type p = (k << 6) + size_type(__builtin_ctzll(cur));
        return p < n_ ? p : npos;
      }
      if (++k >= w_.size()) return npos;
      cur = w_[k];
    }
  }
  size_type n_;
  std::vector<uint64_t> w_;
};
template <typename B, typename A>
inline dynamic_bitset<B, A> operator&(const dynamic_bitset<B, A> &a, const dynamic_bitset<B, A> &b) {
  dynamic_bitset<B, A> r(a);
  r &= b;
  return r;
}
}  // namespace boost
