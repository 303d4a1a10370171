// emul_union.cpp — runs the bit-plane forms of the union pass on the CPU through cuda_emul.h, from the same .inl
// sources the GPU build compiles:
//   kernels/union_planes.inl : k_planes_from_bytes + k_pair_hist_planes<EpiSubsets<..>>   (the default form)
//                              k_pair_hist_planes<EpiWriteHist>: one-hot counting       (SELB200_UNION=planes)
// and compares every pair's 64-bin histogram with the byte-wise definition
//   hist[max(a[j], b[j])]++   (sketch/include/sketch/hll.h:1191-1206 of the reference),
// and the pairs each form hands to the byte kernel ("wide") with the rule it documents.
// Test infrastructure (tests/test_emul_union.py builds and runs it); exit code 0 = all cases identical.

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>

#define SELB_EMUL 1
#include "cuda_emul.h"

struct SrcPairs {
    const uint2* pairs;
    const int32_t* order;
    long long n;
    const unsigned long long* n_dev;
    long long count() const { return n_dev ? (long long)min((unsigned long long)n, *n_dev) : n; }
    uint2 rows(long long pi, uint2& id) const {
        id = pairs[pi];
        return order ? make_uint2((uint32_t)order[id.x], (uint32_t)order[id.y]) : id;
    }
    long long slot(long long pi) const { return pi; }
};
struct EpiWriteHist {
    uint32_t* out;
    void operator()(long long pi, uint2, uint32_t s0, uint32_t s1, uint32_t lane) const {
        out[pi * 64 + lane] = s0;
        out[pi * 64 + 32 + lane] = s1;
    }
};

#include "../../cuda_selection_criteria_b200/csrc/kernels/union_planes.inl"

namespace {

struct Case {
    const char* name;
    int p;
    std::vector<double> load;      // n/m of each genome (register law P(v <= k) = exp(-load * 2^-k))
    double shared;                 // fraction of a genome's load that comes from a common core
    int clamp_hi;                  // registers are clamped to [0, clamp_hi]
    unsigned grid;
    int outlier_every = 0;         // > 0: every such genome gets one register of value 50 (a range wider than 32)
};

int run_case(const Case& cs, uint64_t seed) {
    const size_t m = (size_t)1 << cs.p;
    const int n = (int)cs.load.size();
    const int chunk_regs = (int)std::min<size_t>(m, PL_CHUNK_REGS);
    std::mt19937_64 rng(seed);
    std::uniform_real_distribution<double> U(1e-12, 1.0);
    auto draw = [&](double load) {
        if (load <= 0) return 0;
        const double k = std::ceil(std::log2(load / -std::log(U(rng))));
        return (int)std::max(0.0, std::min((double)cs.clamp_hi, k));
    };
    // registers: max(core part, private part) — similar genomes share their high registers
    std::vector<uint8_t> core(m), regs((size_t)n * m);
    for (size_t j = 0; j < m; ++j) core[j] = (uint8_t)draw(cs.load[0] * cs.shared);
    for (int g = 0; g < n; ++g)
        for (size_t j = 0; j < m; ++j) {
            const int c = cs.shared > 0 ? (int)std::min<int>(core[j], cs.clamp_hi) : 0;
            regs[(size_t)g * m + j] = (uint8_t)std::max(c, draw(cs.load[g] * (1.0 - cs.shared)));
        }
    if (cs.outlier_every > 0)
        for (int g = 0; g < n; g += cs.outlier_every) regs[(size_t)g * m + 7] = 50;
    // per-genome histogram + range, as k_pair_hist(SrcSelf) / k_genome_cards leave them
    std::vector<uint32_t> ghist((size_t)n * 64, 0);
    std::vector<uint16_t> grange(n);
    for (int g = 0; g < n; ++g) {
        int vmin = 63, vmax = 0;
        for (size_t j = 0; j < m; ++j) ghist[(size_t)g * 64 + regs[(size_t)g * m + j]]++;
        for (int b = 0; b < 64; ++b)
            if (ghist[(size_t)g * 64 + b]) { vmin = std::min(vmin, b); vmax = std::max(vmax, b); }
        grange[g] = (uint16_t)(std::min(vmin, vmax) | (vmax << 8));
    }
    // all pairs
    std::vector<uint2> pairs;
    for (int a = 0; a < n; ++a)
        for (int b = a + 1; b < n; ++b) pairs.push_back(a & 1 ? make_uint2(b, a) : make_uint2(a, b));
    const long long np = (long long)pairs.size();
    unsigned long long np_dev = (unsigned long long)np;
    SrcPairs src{pairs.data(), nullptr, np, &np_dev};
    int bad = 0;

    // runs one form, then checks its wide set against `expect_wide` and every other pair against the definition
    auto check = [&](const char* form, const std::function<void(EpiWriteHist, uint32_t*, unsigned long long*)>& run,
                     const std::function<bool(uint32_t, uint32_t)>& expect_wide) {
        std::vector<uint32_t> hist((size_t)np * 64, 0xDEADBEEFu), wide(np, 0);
        unsigned long long counters[3] = {0, 0, 0};           // wide count, batch counter, error word
        std::memset(pl_smem, 0x5A, sizeof pl_smem);
        emul_tma_bytes = emul_tma_expected = 0;
        emul::n_lop3 = emul::n_popc = 0;
        run(EpiWriteHist{hist.data()}, wide.data(), counters);
        const uint64_t ops_lop3 = emul::n_lop3, ops_popc = emul::n_popc;
        int fbad = 0;
        if (counters[2]) { printf("  %s/%s: kernel error word %llx\n", cs.name, form, counters[2]); ++fbad; }
        if (emul_tma_bytes != emul_tma_expected) { printf("  %s/%s: expect_tx %llu != copied %llu\n", cs.name, form, (unsigned long long)emul_tma_expected, (unsigned long long)emul_tma_bytes); ++fbad; }
        std::vector<char> is_wide(np, 0);
        for (unsigned long long w = 0; w < counters[0]; ++w) is_wide[wide[w]] = 1;
        long long n_done = 0;
        for (long long pi = 0; pi < np; ++pi) {
            const uint32_t a = pairs[pi].x, b = pairs[pi].y;
            const bool ew = expect_wide(a, b);
            if (ew != (bool)is_wide[pi]) { printf("  %s/%s: pair %lld wide=%d expected %d\n", cs.name, form, pi, is_wide[pi], ew); ++fbad; continue; }
            if (ew) continue;
            uint32_t want[64] = {0};
            for (size_t j = 0; j < m; ++j) want[std::max(regs[(size_t)a * m + j], regs[(size_t)b * m + j])]++;
            ++n_done;
            if (std::memcmp(want, &hist[(size_t)pi * 64], sizeof want)) {
                if (fbad < 5) {
                    printf("  %s/%s: pair %lld (%u,%u) differs:", cs.name, form, pi, a, b);
                    for (int v = 0; v < 64; ++v)
                        if (want[v] != hist[(size_t)pi * 64 + v]) printf(" [%d] %u!=%u", v, hist[(size_t)pi * 64 + v], want[v]);
                    printf("\n");
                }
                ++fbad;
            }
        }
        // per counted pair and lane: lop3<> and __popc calls as the source writes them (the butterfly's 32 __popc
        // included).  The one-hot form writes its 16 mask ANDs per group and step as plain `&` (LOP3 in SASS, not
        // counted here): tests/emul/union_lop3_model.py and tools/sass_loops.py give the comparable numbers.
        printf("%-26s %-7s p=%d n=%d pairs=%lld counted=%lld wide=%llu lop3<>/pair=%.0f popc/pair=%.0f  %s\n", cs.name, form, cs.p, n,
               np, n_done, counters[0], n_done ? (double)ops_lop3 / 32.0 / (double)n_done : 0.0,
               n_done ? (double)ops_popc / 32.0 / (double)n_done : 0.0, fbad ? "FAIL" : "ok");
        bad += fbad;
    };

    // ---- plane form, one-hot counting (SELB200_UNION=planes) ----
    std::vector<uint32_t> planes((size_t)n * 6 * (m >> 5), 0xA5A5A5A5u);
    emul::launch(2, [&] { k_planes_from_bytes(regs.data(), n, m, chunk_regs, planes.data()); });
    check("planes",
          [&](EpiWriteHist epi, uint32_t* wide, unsigned long long* counters) {
              emul::launch(cs.grid, [&] {
                  k_pair_hist_planes<EpiWriteHist>(planes.data(), m, chunk_regs, grange.data(), src, epi, wide, counters, counters + 1);
              });
          },
          [&](uint32_t a, uint32_t b) {   // the pair's values must fit 32 consecutive values starting at a multiple of 8
              const int lo = std::max(grange[a] & 0xff, grange[b] & 0xff), hi = std::max(grange[a] >> 8, grange[b] >> 8);
              return (hi >> 3) > std::min(lo >> 3, 4) + 3;
          });
    // ---- plane form with subset counting on groups of four values (the default): same wide rule ----
    check("subsets",
          [&](EpiWriteHist epi, uint32_t* wide, unsigned long long* counters) {
              emul::launch(cs.grid, [&] {
                  k_pair_hist_planes<EpiSubsets<EpiWriteHist>>(planes.data(), m, chunk_regs, grange.data(), src, EpiSubsets<EpiWriteHist>{epi}, wide, counters, counters + 1);
              });
          },
          [&](uint32_t a, uint32_t b) {
              const int lo = std::max(grange[a] & 0xff, grange[b] & 0xff), hi = std::max(grange[a] >> 8, grange[b] >> 8);
              return (hi >> 3) > std::min(lo >> 3, 4) + 3;
          });
    // ---- subset counting with the per-step group limit from the per-eighth maxima (m >= 16384: the library's default) ----
    if (m >= 16384) {
        std::vector<uint32_t> gtop((size_t)n * 8, 0u), planes2((size_t)n * 6 * (m >> 5), 0x5A5A5A5Au);
        emul::launch(2, [&] { k_planes_from_bytes(regs.data(), n, m, chunk_regs, planes2.data(), gtop.data()); });
        if (planes2 != planes) { printf("  %s: planes differ when the tops are written too\n", cs.name); ++bad; }
        for (int g = 0; g < n; ++g)
            for (int j = 0; j < 8; ++j) {
                uint32_t mx = 0;
                for (size_t r = (size_t)j * (m / 8); r < (size_t)(j + 1) * (m / 8); ++r) mx = std::max<uint32_t>(mx, regs[(size_t)g * m + r]);
                if (gtop[(size_t)g * 8 + j] != mx) { printf("  %s: gtop[%d][%d] = %u, want %u\n", cs.name, g, j, gtop[(size_t)g * 8 + j], mx); ++bad; }
            }
        check("subtops",
              [&](EpiWriteHist epi, uint32_t* wide, unsigned long long* counters) {
                  emul::launch(cs.grid, [&] {
                      k_pair_hist_planes<EpiSubsets<EpiWriteHist>>(planes.data(), m, chunk_regs, grange.data(), src, EpiSubsets<EpiWriteHist>{epi}, wide, counters, counters + 1,
                                                                   nullptr, 0u, gtop.data());
                  });
              },
              [&](uint32_t a, uint32_t b) {
                  const int lo = std::max(grange[a] & 0xff, grange[b] & 0xff), hi = std::max(grange[a] >> 8, grange[b] >> 8);
                  return (hi >> 3) > std::min(lo >> 3, 4) + 3;
              });
    }
    return bad;
}

}  // namespace

int main(int argc, char** argv) {
    int bad = 0;
    if (argc > 4 && std::string(argv[2]) == "fuzz") {        // emul_union LIB fuzz SEED COUNT: random cases
        std::mt19937_64 rng(strtoull(argv[3], nullptr, 10));
        const int count = atoi(argv[4]);
        auto uni = [&](double a, double b) { return std::uniform_real_distribution<double>(a, b)(rng); };
        for (int it = 0; it < count; ++it) {
            Case cs;
            cs.name = "fuzz";
            cs.p = 9 + (int)(rng() % 7);
            const int n = 2 + (int)(rng() % 9);
            const double centre = std::exp(uni(std::log(0.05), std::log(2e5)));
            const double spread = uni(0.0, 3.0);
            for (int g = 0; g < n; ++g) cs.load.push_back(rng() % 17 == 0 ? 0.0 : centre * std::exp(uni(-spread, spread)));
            cs.shared = (rng() % 3 == 0) ? 0.0 : uni(0.1, 0.99);
            cs.clamp_hi = 64 - cs.p + 1;
            cs.grid = 1 + (unsigned)(rng() % 3);
            cs.outlier_every = (rng() % 3 == 0) ? 2 + (int)(rng() % 3) : 0;
            const int b = run_case(cs, rng());
            if (b) printf("FUZZ FAILURE: p=%d n=%d centre=%g spread=%g shared=%g grid=%u outliers=%d\n", cs.p, n, centre, spread,
                          cs.shared, cs.grid, cs.outlier_every);
            bad += b;
        }
        printf(bad ? "FAILED (%d)\n" : "all identical\n", bad);
        return bad ? 1 : 0;
    }
    const std::vector<Case> cases = {
        {"bacterial p14", 14, {61, 80, 122, 200, 305, 488, 480, 300}, 0.85, 51, 3},
        {"identical-ish p14", 14, {400, 400, 400, 401}, 0.98, 51, 2},
        {"disjoint p14", 14, {150, 160, 170}, 0.0, 51, 1},
        {"large genomes base 8 p14", 14, {3000, 3300, 2800, 3100}, 0.8, 51, 2},
        {"mixed bases p14", 14, {800, 900, 1400, 1800, 2500}, 0.5, 51, 2},
        {"long lists p14", 14, {900, 950, 1000, 1190, 40000}, 0.7, 51, 2},   // ~450-500 entries fit, ~600 overflow the slot
        {"tiny sketches p9", 9, {100, 120, 140, 90, 300}, 0.7, 56, 2},
        {"p12 one chunk", 12, {200, 220, 180}, 0.8, 53, 1},
        {"p16 sixteen chunks", 16, {250, 260}, 0.9, 49, 1},
        {"p15 eight chunks", 15, {120, 260, 700}, 0.6, 50, 2},      // an eighth of the sketch is one chunk, two steps
        {"empty and sparse p14", 14, {0, 0.01, 0.5, 3}, 0.0, 51, 2},
        {"saturated values p14", 14, {1e12, 2e12, 200}, 0.5, 51, 1},
        // genome 0 has another base: its 25 pairs come first in the list, so with batches of 16 a whole batch is wide
        {"whole batches wide p10", 10, {9000, 100, 101, 102, 103, 104, 105, 106, 107, 108, 109, 110, 111, 112, 113, 114, 115, 116, 117, 118,
                                        119, 120, 121, 122, 123, 124}, 0.0, 55, 3},
        // the plane kernel hands every pair of genomes 0, 3, 6, ... to the byte kernel (25 + 22 + ... consecutive pairs);
        // for the split kernel the outlier is one more list entry
        {"outliers p10", 10, {100, 101, 102, 103, 104, 105, 106, 107, 108, 109, 110, 111, 112, 113, 114, 115, 116, 117, 118, 119,
                              120, 121, 122, 123, 124, 125}, 0.5, 54, 3, 3},
        {"many pairs few warps p10", 10, {50, 55, 60, 65, 70, 75, 80, 85, 90, 95, 100, 105, 110, 115, 120, 125, 130, 135}, 0.8, 55, 2},
    };
    uint64_t seed = 12345;
    const bool first_fail = getenv("EMUL_UNION_FIRST_FAIL") != nullptr;      // mutation tests: one failing case is enough
    for (const Case& cs : cases) {
        bad += run_case(cs, seed++);
        if (bad && first_fail) break;
    }
    printf(bad ? "FAILED (%d)\n" : "all identical\n", bad);
    return bad ? 1 : 0;
}
