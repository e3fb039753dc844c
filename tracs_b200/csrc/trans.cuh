// Device-side TransCluster series (shared by trans.cu and the fused path in sweep.cu).
// Restates reference src/transcluster.hpp:62-75,131-238 -- see trans.cu for the derivation.
#pragma once
#include <math.h>
#include <stdint.h>

namespace tracs {

__device__ __forceinline__ double lae(double x, double y) {
  // transcluster.hpp:62-75
  const double t = x - y;
  if (x == y) return x + 0.69314718055994530942;
  if (t > 0) return x + log1p(exp(-t));
  else if (t <= 0) return y + log1p(exp(t));
  return t;
}

// one (N, delta) key -> log P(k = 0 | N, delta) and E[K]
__device__ __forceinline__ void trans_eval(int64_t N, double delta, const double *__restrict__ lg, double lamb, double beta,
                                           double thr, double *p0_out, double *eK_out) {
  const double ln_l = log(lamb), ln_b = log(beta), ln_lb = log(lamb + beta);
  if (!(delta > 0)) {
    // transcluster.hpp:163-167 with k = 0 ; E[K]: negative-binomial mean (see trans.cu header)
    *p0_out = (double)(N + 1) * ln_l + lg[N + 1] - lg[N + 1] - lg[1] - (double)(N + 1) * ln_lb;
    *eK_out = (double)(N + 1) * beta / lamb;
    return;
  }
  const double ln_ld = log(lamb * delta);
  const double ln_d = log(delta);
  double pois = -INFINITY;
  for (int64_t i = 0; i <= N; ++i) pois = lae((double)i * ln_ld - lg[i + 1], pois);
  const double lx = ln_d + ln_lb;
  double G = -INFINITY;  // G_M = LSE_{j<=M}( j*ln(delta*(lamb+beta)) - lg[j+1] ), here M = N
  for (int64_t j = 0; j <= N; ++j) G = lae((double)j * lx - lg[j + 1], G);
  const double common = (double)(N + 1) * ln_l - lg[N + 1] - delta * beta - pois;
  {
    const double lhs = common + lg[N + 1] - lg[1];
    *p0_out = lhs + (G - (double)(N + 1) * ln_lb);
  }
  const double ub = exp(ln_b + delta * lamb + log((double)(N + 1)) - (ln_l + pois));
  double lprob = -INFINITY, elprob = -INFINITY, diff = thr + 1.0;
  int64_t k = 1;
  while (diff > thr && k < 10000) {  // a NaN bound ends the loop after one term, like the reference's `diff > threshold_Ek`
    const int64_t M = N + k;
    G = lae((double)M * lx - lg[M + 1], G);
    const double lhs = common + (double)k * ln_b + lg[M + 1] - lg[k + 1];
    const double lp = lhs + (G - (double)(M + 1) * ln_lb);
    const double lk = log((double)k);
    lprob = lae(lprob, lp + lk);
    elprob = lae(elprob, lhs + lk + delta * (lamb + beta) - (double)(M + 1) * ln_lb);
    diff = ub - exp(elprob);
    ++k;
  }
  *eK_out = exp(lprob);
}

}  // namespace tracs
