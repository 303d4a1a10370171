// estimators.cuh — fp64 arithmetic of the selection path, written once for device
// (and host-callable for the launcher code that needs the same formulas).
//
// Everything here must be compiled with -fmad=false: the sequence of IEEE-754
// double operations is then exactly the one written below (the reference is
// g++ code whose contraction is compiler-dependent; SURVEY.md §7 "Hard parts" 1).
//
// Reference (paths relative to the reference tree):
//   ertl_mle      sketch/include/sketch/hll.h:628-688 (ertl_ml_estimate), relerr=1e-2 (hll.h:212)
//   sigma_p       include/criteria_sketch.hpp:7-20
//   crit_cb       include/criteria_sketch.hpp:45-49, called at src/selection.cpp:282
//   crit_hll_a    include/criteria_sketch.hpp:60-64 + 36-43 (kota_mas)
//   crit_hll_an   include/criteria_sketch.hpp:52-58 + 22-34 (cota_n)
//   jaccard       src/selection.cpp:287
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define SELB_HD __host__ __device__ __forceinline__
#else
#define SELB_HD inline
#endif

namespace selb {

SELB_HD int imax(int a, int b) { return a > b ? a : b; }
SELB_HD int imin(int a, int b) { return a < b ? a : b; }

struct NeverStop {
    SELB_HD bool operator()(double) const { return false; }
};

// Ertl maximum-likelihood cardinality from a register histogram c[0..q+1], q = 64-p.
// CountT is uint32_t (global/shared histogram rows).  `stride` lets the caller keep
// its histogram column-interleaved in shared memory.
//
// `stop(t_lb)` is an optional early exit.  The secant iteration never decreases x: deltaX is
// either 0 or deltaX*(g-m')/(gprev-g) with g<=m' and gprev<g, i.e. a product of non-negative
// factors, and x += deltaX — in IEEE arithmetic too, since every operation involved is monotone.
// So x*m at any point is a LOWER BOUND of the value the full iteration returns.  A caller whose
// decision is monotone non-increasing in the estimate (Jaccard >= tau, hll_a, hll_an) may
// therefore stop as soon as the decision already fails at the bound: the outcome is identical
// to running the reference iteration to its end.  When stop() fires, *stopped is set and the
// bound is returned; otherwise the result is bit-for-bit the reference's sequence of operations.
// The estimator proper, for a histogram whose smallest and largest non-empty bins are already known (kMin <= kMax <= q+1).
template <typename CountT, typename Stop = NeverStop>
SELB_HD double ertl_mle_range(const CountT* c, int p, int stride, int kMin, int kMax, Stop stop = Stop(), bool* stopped = nullptr);

template <typename CountT, typename Stop = NeverStop>
SELB_HD double ertl_mle(const CountT* c, int p, int stride = 1, Stop stop = Stop(), bool* stopped = nullptr) {
    const int q = 64 - p;
    const unsigned long long m = 1ull << p;
    if ((unsigned long long)c[(q + 1) * stride] == m) return __builtin_huge_val();
    int kMin, kMax;
    for (kMin = 0; kMin <= q + 1 && c[kMin * stride] == 0; ++kMin) {}
    if (kMin > q + 1) return 0.;   // empty histogram: cannot happen for a sketch (counts sum to m); guards the scan
    for (kMax = q + 1; kMax && c[kMax * stride] == 0; --kMax) {}
    return ertl_mle_range(c, p, stride, kMin, kMax, stop, stopped);
}

template <typename CountT, typename Stop>
SELB_HD double ertl_mle_range(const CountT* c, int p, int stride, int kMin, int kMax, Stop stop, bool* stopped) {
    const int q = 64 - p;
    const unsigned long long m = 1ull << p;
    if ((unsigned long long)c[(q + 1) * stride] == m) return __builtin_huge_val();
    const int kMinP = imax(1, kMin);
    const int kMaxP = imin(q, kMax);
    double z = 0.;
    for (int k = kMaxP; k >= kMinP; --k) z = 0.5 * z + (double)c[k * stride];
    z = ldexp(z, -kMinP);
    unsigned cP = (unsigned)c[(q + 1) * stride];
    if (q) cP += (unsigned)c[kMaxP * stride];
    const double a = z + (double)c[0];
    const int mP = (int)(m - (unsigned long long)c[0]);
    double gprev = z + ldexp((double)c[(q + 1) * stride], -q);
    double x = gprev <= 1.5 * a ? (double)mP / (0.5 * gprev + a) : ((double)mP / gprev) * log1p(gprev / a);
    gprev = 0.;
    double dx = x;
    const double relerr = 1e-2 / sqrt((double)m);
    while (dx > x * relerr) {
        if (stop(x * (double)m)) {
            if (stopped) *stopped = true;
            return x * (double)m;
        }
        int kappaM1;
        frexp(x, &kappaM1);
        double xp = ldexp(x, -imax(kMaxP + 1, kappaM1 + 2));
        const double xp2 = xp * xp;
        double h = xp - xp2 / 3 + (xp2 * xp2) * (1. / 45. - xp2 / 472.5);
        for (int k = kappaM1; k >= kMaxP; --k) {
            const double hp = 1. - h;
            h = (xp + h * hp) / (xp + hp);
            xp += xp;
        }
        double g = (double)cP * h;
        for (int k = kMaxP - 1; k >= kMinP; --k) {
            const double hp = 1. - h;
            h = (xp + h * hp) / (xp + hp);
            xp += xp;
            g += (double)c[k * stride] * h;
        }
        g += x * a;
        if (gprev < g && g <= (double)mP) dx *= (g - (double)mP) / (gprev - g);
        else dx = 0.;
        x += dx;
        gprev = g;
    }
    return x * (double)m;
}

SELB_HD float sigma_p(int p) {
    const double s = sqrt((double)(1 << p));
    double c = 1.039;
    if (p == 4) c = 1.106;
    else if (p == 5) c = 1.07;
    else if (p == 6) c = 1.054;
    else if (p == 7) c = 1.046;
    return (float)(c / s);
}

// tau is (double)(float)threshold; cardinalities are the size_t truncations.
SELB_HD bool crit_cb(double tau, unsigned long long e1, unsigned long long e2) {
    const double gamma = (double)e1 / (double)e2;
    return gamma >= tau;
}

// zs = Z * sigma_p as a FLOAT product (criteria_sketch.hpp:29,32,40).
SELB_HD bool crit_hll_a(double tau, unsigned long long e1, unsigned long long e2, double t_union, float zs) {
    const unsigned long long t_trunc = (unsigned long long)t_union;   // size_t t_hat (:61)
    const double t_hat = (double)t_trunc;
    const double gamma = (double)e1 / (double)e2;
    const double t_mas = t_hat / (1.0 + (double)zs);
    const double k_mas = ((1.0 + gamma) * (double)e2 - t_mas) / t_mas;
    return k_mas >= tau;
}

SELB_HD bool crit_hll_an(double tau, unsigned long long e1, unsigned long long e2, double t_hat, float zs,
                         int order_n) {
    const double j_hat = ((double)(e1 + e2) - t_hat) / t_hat;          // size_t sum (:55)
    const double gamma = (double)e1 / (double)e2;
    double S = 0., num = 1.;
    for (int k = 1; k < order_n + 1; ++k) {
        num *= (double)zs;
        S += num;
    }
    const double lim = (1.0 + (double)zs) * (double)e2 / t_hat;
    const double minimo = lim < 1.0 ? lim : 1.0;                        // std::min(1.0, lim)
    const double C = minimo * (1 + gamma) * S;
    return (j_hat + C) >= tau;
}

// Sufficient (never necessary) fp32 test that hll_a (an = 0) / hll_an (an = 1) FAIL for a pair whose auxiliary union has
// harmonic sum (over its non-empty registers) at most z_ub and at most c0_ub empty registers, m = 2^p_aux registers, none
// of them at q+1.  ertl_mle starts from x0 = m'/(0.5 g + a), g = z, a = z + c[0], m' = m - c[0] (its first branch: g <= 1.5 a
// always holds when c[q+1] = 0) and never decreases x, so t_lb = m x0 is a lower bound of the estimate; x0 decreases in z
// and c[0], so upper bounds of those give a lower bound of x0.  Both criteria are non-increasing in the estimate
// (Z sigma >= 0), so failing at the lower bound means failing.  Everything is biased towards "not sure": z is inflated
// and t_lb deflated by 2e-5 (fp32 rounding of the whole chain stays below 4e-6), hll_a's size_t truncation is charged a
// full unit, and the criterion must miss tau by 1e-4.  A pair that is not rejected here is decided exactly elsewhere.
SELB_HD bool hll_surely_fails(int an, float tau, float zs, int order_n, float m, float e1, float e2, float z_ub, float c0_ub) {
    const float den = 1.5f * (z_ub * 1.00002f) + c0_ub;
    if (!(den > 0.f) || !(zs >= 0.f)) return false;
    const float t_lb = m * (m - c0_ub) / den * 0.99998f;
    if (!(t_lb > 2.f)) return false;
    const float s = e1 + e2;
    if (!an) {
        const float t_mas = (t_lb - 1.f) / (1.f + zs);
        return (s - t_mas) / t_mas < tau - 1e-4f;
    }
    float S = 0.f, num = 1.f;
    for (int k = 1; k < order_n + 1; ++k) {
        num *= zs;
        S += num;
    }
    const float lim = (1.f + zs) * e2 / t_lb;
    const float C = (lim < 1.f ? lim : 1.f) * (s / e2) * S * 1.00002f;
    return (s - t_lb) / t_lb + C < tau - 1e-4f;
}

SELB_HD double jaccard(unsigned long long e1, unsigned long long e2, double t) {
    return ((double)e1 + (double)e2 - t) / t;
}

}  // namespace selb
