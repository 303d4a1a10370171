// emul_split.cpp — runs kernels/union_split.inl (k_split_build + k_pair_hist_split) on the CPU through cuda_emul.h
// and compares every pair's 64-bin histogram with the byte-wise definition
//   hist[max(a[j], b[j])]++   (sketch/include/sketch/hll.h:1191-1206 of the reference).
// Test infrastructure (tests/test_emul_split.py builds and runs it); exit code 0 = all cases identical.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>

#include "cuda_emul.h"

constexpr int PL_STAGES = 2;
constexpr int PL_CHUNK_REGS = 4096;
constexpr int PL_NQ = PL_CHUNK_REGS / 64;

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

#include "../../cuda_selection_criteria_b200/csrc/kernels/union_split.inl"

namespace {

struct Case {
    const char* name;
    int p;
    std::vector<double> load;      // n/m of each genome (register law P(v <= k) = exp(-load * 2^-k))
    double shared;                 // fraction of a genome's load that comes from a common core
    int clamp_hi;                  // registers are clamped to [0, clamp_hi]
    unsigned grid;
};

int run_case(const Case& cs, uint64_t seed) {
    const size_t m = (size_t)1 << cs.p;
    const int n = (int)cs.load.size();
    const int chunk_regs = (int)std::min<size_t>(m, PL_CHUNK_REGS);
    const int nchunks = (int)(m / chunk_regs);
    const uint32_t cb = split_chunk_bytes(chunk_regs);
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
    const size_t rec_bytes = (size_t)(nchunks + 1) * cb;
    uint8_t* rec = (uint8_t*)aligned_alloc(128, ((size_t)n * rec_bytes + 127) / 128 * 128);
    std::memset(rec, 0xA5, (size_t)n * rec_bytes);       // stale bytes behind short lists must not matter
    std::vector<uint32_t> gmeta(n, 0);
    emul::launch(std::min<unsigned>(n, 3), [&] {
        k_split_build(regs.data(), n, m, chunk_regs, grange.data(), ghist.data(), rec, gmeta.data());
    });
    // all pairs
    std::vector<uint2> pairs;
    for (int a = 0; a < n; ++a)
        for (int b = a + 1; b < n; ++b) pairs.push_back(a & 1 ? make_uint2(b, a) : make_uint2(a, b));
    const long long np = (long long)pairs.size();
    std::vector<uint32_t> hist((size_t)np * 64, 0xDEADBEEFu), wide(np, 0);
    unsigned long long counters[3] = {0, 0, 0};           // wide count, batch counter, error word
    unsigned long long np_dev = (unsigned long long)np;
    SrcPairs src{pairs.data(), nullptr, np, &np_dev};
    EpiWriteHist epi{hist.data()};
    std::memset(pl_smem, 0x5A, sizeof pl_smem);
    emul_tma_bytes = emul_tma_expected = 0;
    emul::launch(cs.grid, [&] {
        k_pair_hist_split<EpiWriteHist>(rec, m, chunk_regs, gmeta.data(), src, epi, wide.data(), counters, counters + 1);
    });
    int bad = 0;
    if (counters[2]) { printf("  %s: kernel error word %llx\n", cs.name, counters[2]); ++bad; }
    if (emul_tma_bytes != emul_tma_expected) { printf("  %s: expect_tx %llu != copied %llu\n", cs.name, (unsigned long long)emul_tma_expected, (unsigned long long)emul_tma_bytes); ++bad; }
    std::vector<char> is_wide(np, 0);
    for (unsigned long long w = 0; w < counters[0]; ++w) is_wide[wide[w]] = 1;
    long long n_dense = 0, n_high = 0;
    for (long long pi = 0; pi < np; ++pi) {
        const uint32_t a = pairs[pi].x, b = pairs[pi].y;
        const bool expect_wide = (gmeta[a] & 0xff) != (gmeta[b] & 0xff) || (gmeta[a] >> 8) == SPLIT_LEN_OVERFLOW ||
                                 (gmeta[b] >> 8) == SPLIT_LEN_OVERFLOW;
        if (expect_wide != (bool)is_wide[pi]) { printf("  %s: pair %lld wide=%d expected %d\n", cs.name, pi, is_wide[pi], expect_wide); ++bad; continue; }
        if (expect_wide) continue;
        uint32_t want[64] = {0};
        for (size_t j = 0; j < m; ++j) want[std::max(regs[(size_t)a * m + j], regs[(size_t)b * m + j])]++;
        ++n_dense;
        n_high += (gmeta[a] >> 8) + (gmeta[b] >> 8);
        if (std::memcmp(want, &hist[(size_t)pi * 64], sizeof want)) {
            if (bad < 5) {
                printf("  %s: pair %lld (%u,%u) base %u lens %u %u differs:", cs.name, pi, a, b, gmeta[a] & 0xff, gmeta[a] >> 8, gmeta[b] >> 8);
                for (int v = 0; v < 64; ++v)
                    if (want[v] != hist[(size_t)pi * 64 + v]) printf(" [%d] %u!=%u", v, hist[(size_t)pi * 64 + v], want[v]);
                printf("\n");
            }
            ++bad;
        }
    }
    printf("%-28s p=%d n=%d pairs=%lld split=%lld wide=%llu avg list=%.1f  %s\n", cs.name, cs.p, n, np, n_dense, counters[0],
           n_dense ? (double)n_high / (2.0 * n_dense) : 0.0, bad ? "FAIL" : "ok");
    free(rec);
    return bad;
}

}  // namespace

int main() {
    int bad = 0;
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
        {"empty and sparse p14", 14, {0, 0.01, 0.5, 3}, 0.0, 51, 2},
        {"saturated values p14", 14, {1e12, 2e12, 200}, 0.5, 51, 1},
        // genome 0 has another base: its 39 pairs are consecutive in the list, so whole batches are wide
        {"whole batches wide p10", 10, {9000, 100, 101, 102, 103, 104, 105, 106, 107, 108, 109, 110, 111, 112, 113, 114, 115, 116, 117, 118,
                                        119, 120, 121, 122, 123, 124, 125, 126, 127, 128, 129, 130, 131, 132, 133, 134, 135, 136, 137, 138}, 0.0, 55, 2},
        {"many pairs few warps p10", 10, {50, 55, 60, 65, 70, 75, 80, 85, 90, 95, 100, 105, 110, 115, 120, 125, 130, 135}, 0.8, 55, 2},
    };
    uint64_t seed = 12345;
    for (const Case& cs : cases) bad += run_case(cs, seed++);
    printf(bad ? "FAILED (%d)\n" : "all identical\n", bad);
    return bad ? 1 : 0;
}
