"""Independent checks of the binomial CDF used by the recombination filter (TEST INFRASTRUCTURE ONLY).

The reference evaluates 1 - boost::math::cdf(binomial_distribution(n, p), k) (src/pairsnp.hpp:41-58, :251-318); Boost
is absent from this image, so oracle/_ref, the C oracle and the CUDA kernel all sum the pmf term by term
(oracle/standin/boost/math/distributions/binomial.hpp). Boost computes the same quantity through the regularised
incomplete beta function: cdf(n, p, k) = ibetac(k + 1, n - k, p) = I_{1-p}(n - k, k + 1). This file restates
  * that formula with a continued fraction written from the published algorithm (DLMF 8.17.22, modified Lentz), and
  * the filter's window enumeration (which (count, span) pairs it ever evaluates),
so that tests/test_filter_pin.py can compare every keep / drop decision under three CDF implementations (pmf sum,
this continued fraction, scipy.stats.binom -- itself a regularised incomplete beta) and report the smallest margin
to the decision threshold."""
import math


def _betacf(a, b, x, eps=3e-16, max_iter=10000):
    """Continued fraction of the incomplete beta function (DLMF 8.17.22), modified Lentz evaluation."""
    tiny = 1e-300
    qab, qap, qam = a + b, a + 1.0, a - 1.0
    c, d = 1.0, 1.0 - qab * x / qap
    d = 1.0 / (d if abs(d) > tiny else tiny)
    h = d
    for m in range(1, max_iter + 1):
        m2 = 2 * m
        aa = m * (b - m) * x / ((qam + m2) * (a + m2))
        d = 1.0 + aa * d
        d = 1.0 / (d if abs(d) > tiny else tiny)
        c = 1.0 + aa / c
        c = c if abs(c) > tiny else tiny
        h *= d * c
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2))
        d = 1.0 + aa * d
        d = 1.0 / (d if abs(d) > tiny else tiny)
        c = 1.0 + aa / c
        c = c if abs(c) > tiny else tiny
        delta = d * c
        h *= delta
        if abs(delta - 1.0) < eps:
            return h
    raise ArithmeticError("incomplete beta continued fraction did not converge")


def ibeta(a, b, x):
    """Regularised incomplete beta function I_x(a, b)."""
    if x <= 0.0:
        return 0.0
    if x >= 1.0:
        return 1.0
    lbt = math.lgamma(a + b) - math.lgamma(a) - math.lgamma(b) + a * math.log(x) + b * math.log1p(-x)
    if x < (a + 1.0) / (a + b + 2.0):
        return math.exp(lbt) * _betacf(a, b, x) / a
    return 1.0 - math.exp(lbt) * _betacf(b, a, 1.0 - x) / b


def binom_cdf_ibeta(n, p, k):
    """P(X <= k), X ~ Binomial(n, p), through the incomplete beta function: I_{1-p}(n - k, k + 1)."""
    if k >= n:
        return 1.0
    if k < 0:
        return 0.0
    return ibeta(float(n - k), float(k + 1), 1.0 - p)


def binom_cdf_pmf_sum(n, p, k):
    """The stand-in's direct summation (oracle/standin/.../binomial.hpp), term for term."""
    if k >= n:
        return 1.0
    if k < 0:
        return 0.0
    if p <= 0:
        return 1.0
    if p >= 1:
        return 0.0
    lp, lq, s = math.log(p), math.log1p(-p), 0.0
    for i in range(int(math.floor(k)) + 1):
        s += math.exp(math.lgamma(n + 1) - math.lgamma(i + 1) - math.lgamma(n - i + 1) + i * lp + (n - i) * lq)
    return min(s, 1.0)


def filter_windows(snp, L):
    """The windows filter_recomb evaluates for one pair (src/pairsnp.hpp:251-318, range_count :223-248): for every SNP
    position, (count, span) of the pair's SNPs inside [pos - h, pos + h + 1) clipped to the alignment. Yields
    (count, span, p, threshold) per SNP; count <= 1 means "kept without a test"."""
    d = len(snp)
    if d <= 1:
        return
    p = d / float(int(L))
    thr = 0.05 / d
    h = int(1.0 / p / 2.0 + 1)
    h = max(50, min(5000, h))
    for i in snp:
        left, right = max(0, i - h), min(int(L), i + h + 1)
        inside = [t for t in snp if left <= t < right]
        yield len(inside), (inside[-1] - inside[0] + 1) if inside else 0, p, thr


def filtered_distance(snp, L, cdf):
    """filt for one pair under the CDF implementation `cdf(n, p, k)`; also returns the smallest relative margin
    |(1 - cdf) - thr| / thr over the windows that were tested."""
    d = len(snp)
    if d <= 1:
        return d, float("inf")
    kept, margin = 0, float("inf")
    for cnt, span, p, thr in filter_windows(snp, L):
        if cnt > 1:
            pv = 1.0 - cdf(span, p, cnt)
            margin = min(margin, abs(pv - thr) / thr)
            kept += pv >= thr
        else:
            kept += 1
    return kept, margin
