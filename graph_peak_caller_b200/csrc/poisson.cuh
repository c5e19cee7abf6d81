// -log10 of the Poisson survival function, as the reference evaluates it:
//   p = scipy.stats.poisson.logsf(k, lam) / -ln(10)       (sparsepvalues.py:21-22)
// with scipy's logsf(k, mu) = log(pdtrc(floor(k), mu)), pdtrc(k, mu) =
// igam(k + 1, mu) (cephes / xsf).  The reference maps -inf (the survival
// function underflowing to 0 in double) to 1000 (sparsepvalues.py:24); in
// cephes that happens in igam_series when  a*log(x) - x - lgam(a) < -MAXLOG,
// which is replicated here.  Everything else only has to be accurate (the
// bar is 1e-6 relative): the tail sums are evaluated in log space.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define GPC_HD __host__ __device__ __forceinline__
#else
#define GPC_HD inline
#endif

namespace gpc {

constexpr double GPC_LN10 = 2.302585092994046;         // np.log(10)
constexpr double GPC_MAXLOG = 7.09782712893383996843e2;  // cephes MAXLOG

// counts, lambdas -> the reference's clean_p_values for one element.
GPC_HD double neglog10_poisson_sf(double count, double lam) {
    if (count == 0.0) return 0.0;                 // p_values[counts == 0] = 0
    double kf = floor(count);
    if (kf < 0.0) return 0.0;                     // sf == 1 -> log == 0 -> -0.0
    if (!(lam > 0.0)) return 1000.0;              // sf(k >= 0, 0) == 0 -> -inf -> 1000
    const double a = kf + 1.0, x = lam;
    double logsf;
    if (x < a) {
        // upper tail: sf = x^a e^-x / Gamma(a+1) * sum_{n>=0} x^n / ((a+1)...(a+n))
        double ax = a * log(x) - x - lgamma(a);
        bool far = fabs(a - x) > 0.4 * fabs(a);   // igam_fac's non-Lanczos branch
        if (far && ax < -GPC_MAXLOG) return 1000.0;
        double term = 1.0, sum = 1.0, d = a;
        for (int n = 0; n < 100000; ++n) {
            d += 1.0;
            term *= x / d;
            sum += term;
            if (term <= sum * 1e-17) break;
        }
        logsf = ax - log(a) + log(sum);
        if (logsf > 0.0) logsf = 0.0;
    } else {
        // lower tail: cdf = pmf(k) * sum_{j=0..k} k(k-1)...(k-j+1) / x^j ; sf = 1 - cdf
        double lp = kf * log(x) - x - lgamma(kf + 1.0);
        double term = 1.0, sum = 1.0, d = kf;
        while (d > 0.0) {
            term *= d / x;
            sum += term;
            d -= 1.0;
            if (term <= sum * 1e-17) break;
        }
        double cdf = exp(lp) * sum;
        double sf = 1.0 - cdf;                    // double rounding, as 1 - igamc
        if (sf <= 0.0) return 1000.0;
        logsf = log(sf);
    }
    double p = logsf / -GPC_LN10;
    if (isinf(p)) return 1000.0;
    return p == 0.0 ? 0.0 : p;                    // -0.0 -> 0.0 (same dict key)
}

}  // namespace gpc
