// Minimal stand-in for <boost/math/distributions/binomial.hpp> (Boost is not installed here).
// TEST INFRASTRUCTURE ONLY. cdf(binomial_distribution(n,p), k) = P(X <= k), by direct
// log-space summation of the pmf. Only reached on the reference's filter=True sub-path
// (pairsnp.hpp:41-58), whose parity is documented as UNPINNED (real Boost uses ibetac).
#pragma once
#include <cmath>
namespace boost { namespace math {
template <typename Real = double>
class binomial_distribution {
 public:
  binomial_distribution(Real n, Real p) : n_(n), p_(p) {}
  Real trials() const { return n_; }
  Real success_fraction() const { return p_; }
 private:
  Real n_, p_;
};
template <typename Real>
inline Real cdf(const binomial_distribution<Real> &d, Real k) {
  const double n = d.trials(), p = d.success_fraction();
  if (k >= n) return 1.0;
  if (k < 0) return 0.0;
  if (p <= 0) return 1.0;
  if (p >= 1) return 0.0;
  const double lp = std::log(p), lq = std::log1p(-p);
  double s = 0.0;
  for (double i = 0; i <= std::floor(k); i += 1.0)
    s += std::exp(std::lgamma(n + 1) - std::lgamma(i + 1) - std::lgamma(n - i + 1) + i * lp + (n - i) * lq);
  return s > 1.0 ? 1.0 : s;
}
template <typename Real, typename K>
inline Real cdf(const binomial_distribution<Real> &d, K k) { return cdf(d, static_cast<Real>(k)); }
}}  // namespace boost::math
