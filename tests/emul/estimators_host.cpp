// estimators_host.cpp — C entry points over the PRODUCT's fp64 arithmetic (csrc/estimators.cuh) compiled for the
// host with -ffp-contract=off (the device build uses -fmad=false), so that tests/test_estimators_host.py can hold
// it bit-for-bit against the oracle without a GPU.  Test infrastructure: nothing in the package loads this.
#include <cmath>
#include <cstdint>

#include "../../cuda_selection_criteria_b200/csrc/estimators.cuh"

namespace {
struct StopJ {   // the functor of k_estimate_emit (csrc/kernels/estimate_sort.inl)
    double tau, slack;
    unsigned long long e1, e2;
    bool operator()(double t_lb) const { return selb::jaccard(e1, e2, t_lb) < tau - slack; }
};
}  // namespace

extern "C" {
double est_ertl_mle(const uint32_t* c, int p) { return selb::ertl_mle<uint32_t>(c, p); }
double est_ertl_mle_stride(const uint32_t* c, int p, int stride) { return selb::ertl_mle<uint32_t>(c, p, stride); }
double est_ertl_mle_stopj(const uint32_t* c, int p, double tau, uint64_t e1, uint64_t e2, int* stopped) {
    bool st = false;
    const double t = selb::ertl_mle(c, p, 1, StopJ{tau, 1e-6 * std::fabs(tau), e1, e2}, &st);
    *stopped = st;
    return t;
}
float est_sigma_p(int p) { return selb::sigma_p(p); }
int est_cb(double tau, uint64_t e1, uint64_t e2) { return selb::crit_cb(tau, e1, e2); }
int est_hll_a(double tau, uint64_t e1, uint64_t e2, double t, float zs) { return selb::crit_hll_a(tau, e1, e2, t, zs); }
int est_hll_an(double tau, uint64_t e1, uint64_t e2, double t, float zs, int n) { return selb::crit_hll_an(tau, e1, e2, t, zs, n); }
double est_jaccard(uint64_t e1, uint64_t e2, double t) { return selb::jaccard(e1, e2, t); }
// pass A of the hll plane filter: the fp32 sufficient test, and the exact decision it must never contradict
int est_hll_surely_fails(int an, float tau, float zs, int order_n, float m, float e1, float e2, float z_ub, float c0_ub) {
    return selb::hll_surely_fails(an, tau, zs, order_n, m, e1, e2, z_ub, c0_ub);
}
int est_hll_exact(int an, const uint32_t* hist, int p_aux, double tau, uint64_t e1, uint64_t e2, float zs, int order_n) {
    const double t = selb::ertl_mle<uint32_t>(hist, p_aux);
    return an ? selb::crit_hll_an(tau, e1, e2, t, zs, order_n) : selb::crit_hll_a(tau, e1, e2, t, zs);
}
}
