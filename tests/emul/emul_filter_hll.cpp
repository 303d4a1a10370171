// emul_filter_hll.cpp — runs the CB + hll_a / hll_an chain of a selection run on the CPU through cuda_emul.h, from the
// same .inl sources the GPU build compiles:
//   k_cb_bounds -> k_rowblock_span -> exclusive scan -> k_tile_table                 (kernels/tiles.inl)
//   k_aux_planes_quad, k_aux_range (load time) -> k_tile_filter_hll_bound<AN> -> k_hll_verify<AN>   (kernels/filter_hll.inl:
//                                            the two-pass plane filter), or the single pass k_tile_filter_hll_planes<AN>
//                                            (SELB200_HLLFILTER=onepass),
//   or the byte form k_tile_filter_hll<AN> over the transposed registers
// Input (file): sorted truncated cardinalities, auxiliary HLL registers in sorted order, tau, Z*sigma, criterion.
// Output (file): P_cb and the surviving pair list — tests/test_emul_filter.py holds them against the oracle's
// union_size + hll_a / hll_an decision of every pair inside the CB band.  Test infrastructure.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <random>
#include <string>
#include <type_traits>

#define SELB_EMUL 1
#include "cuda_emul.h"
#include "../../cuda_selection_criteria_b200/csrc/estimators.cuh"

constexpr int TILE = 128;          // as in csrc/selb200.cu

#include "../../cuda_selection_criteria_b200/csrc/kernels/helpers.inl"
#include "../../cuda_selection_criteria_b200/csrc/kernels/tiles.inl"
#include "../../cuda_selection_criteria_b200/csrc/kernels/filter_hll.inl"

template <class T> static void rd(FILE* f, T* p, size_t n) { if (fread(p, sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } }
template <class T> static void wr(FILE* f, const T* p, size_t n) { if (fwrite(p, sizeof(T), n, f) != n) { fprintf(stderr, "short write\n"); exit(2); } }

// emul_filter_hll hist-check SEED COUNT: the per-thread histogram step of the plane filter (aux_plane_hist, every
// window) against the byte-wise definition hist[max(a[j], b[j])]++ on random sketch pairs
static int hist_check(uint64_t seed, int count) {
    std::mt19937_64 rng(seed);
    int bad = 0, done = 0, bound_done = 0;
    for (int it = 0; it < count; ++it) {
        const int p_aux = 6 + (int)(rng() % 7);                       // 6 .. 12
        const int m = 1 << p_aux, nw = m >> 5, nbins = 64 - p_aux + 2;
        const int g0 = (int)(rng() % 5);                              // window 8*g0 .. 8*g0 + 31
        int vlo = 8 * g0 + (int)(rng() % 12), vhi = vlo + (int)(rng() % 24);
        vhi = std::min(std::min(vhi, 8 * g0 + 31), nbins - 1);
        vlo = std::min(vlo, vhi);
        if (std::min(vlo >> 3, 4) != g0) continue;                    // the kernel derives the window from vlo
        std::vector<uint8_t> a(m), b(m);
        for (int j = 0; j < m; ++j) {
            a[j] = (uint8_t)(vlo + (int)(rng() % (unsigned)(vhi - vlo + 1)));
            b[j] = (uint8_t)(vlo + (int)(rng() % (unsigned)(vhi - vlo + 1)));
            if (rng() % 5 == 0) b[j] = a[j];
        }
        a[rng() % m] = b[rng() % m] = (uint8_t)vhi;                   // the range is attained
        a[rng() % m] = (uint8_t)vlo; b[rng() % m] = (uint8_t)vlo;
        const long long npad = 2;                                     // two "genomes", quad layout (k_aux_planes_quad)
        std::vector<uint32_t> Q((size_t)6 * nw * npad, 0u);
        for (int g = 0; g < 2; ++g)
            for (int j = 0; j < m; ++j)
                for (int pl = 0; pl < 6; ++pl)
                    if (((g ? b[j] : a[j]) >> pl) & 1) {
                        const int w = j >> 5, slot = (w & 1) * 6 + pl;
                        Q[(((size_t)(w >> 1) * 3 + (slot >> 2)) * npad + g) * 4 + (slot & 3)] |= 1u << (j & 31);
                    }
        uint32_t want[64] = {0};
        int rlo = 63, rhi = 0;
        for (int j = 0; j < m; ++j) { const int v = std::max(a[j], b[j]); want[v]++; rlo = std::min(rlo, v); rhi = std::max(rhi, v); }
        // the kernel's range: max of the minima .. max of the maxima (a superset of the union's own range)
        const int klo = std::max((int)*std::min_element(a.begin(), a.end()), (int)*std::min_element(b.begin(), b.end()));
        const int khi = std::max((int)*std::max_element(a.begin(), a.end()), (int)*std::max_element(b.begin(), b.end()));
        ++done;
        {
            uint32_t gmask = 0;
            for (int t = 0; t < 8; ++t) if (2 * g0 + t >= (klo >> 2) && 2 * g0 + t <= (khi >> 2)) gmask |= 1u << t;
            std::vector<uint32_t> hcol((size_t)64 * 64, 0xDEADBEEFu);
            aux_plane_hist_g(g0, Q.data(), Q.data() + 4, npad, nw, gmask, hcol.data(), nbins);
            // pass A's sums from its twelve-value state: an upper bound of the union's harmonic sum (every register it does
            // not count charged the first uncounted value), the empty count exact
            const int t0 = (klo >> 2) - 2 * g0;
            if (t0 <= 1) {
                uint32_t S[HLLB_NV] = {0}, C2[HLLB_NV] = {0};
                float zb = -1.f, cb = -1.f;
                const uint4* q4 = reinterpret_cast<const uint4*>(Q.data());
                switch (2 * g0 + t0) {
                    case 0: aux_bound_segment<0, 0>(q4, q4 + 1, (uint32_t)npad, 0, nw >> 1, gmask, S, C2, zb, cb); break;
                    case 1: aux_bound_segment<0, 1>(q4, q4 + 1, (uint32_t)npad, 0, nw >> 1, gmask, S, C2, zb, cb); break;
                    case 2: aux_bound_segment<1, 0>(q4, q4 + 1, (uint32_t)npad, 0, nw >> 1, gmask, S, C2, zb, cb); break;
                    case 3: aux_bound_segment<1, 1>(q4, q4 + 1, (uint32_t)npad, 0, nw >> 1, gmask, S, C2, zb, cb); break;
                    case 4: aux_bound_segment<2, 0>(q4, q4 + 1, (uint32_t)npad, 0, nw >> 1, gmask, S, C2, zb, cb); break;
                    case 5: aux_bound_segment<2, 1>(q4, q4 + 1, (uint32_t)npad, 0, nw >> 1, gmask, S, C2, zb, cb); break;
                    case 6: aux_bound_segment<3, 0>(q4, q4 + 1, (uint32_t)npad, 0, nw >> 1, gmask, S, C2, zb, cb); break;
                    case 7: aux_bound_segment<3, 1>(q4, q4 + 1, (uint32_t)npad, 0, nw >> 1, gmask, S, C2, zb, cb); break;
                    case 8: aux_bound_segment<4, 0>(q4, q4 + 1, (uint32_t)npad, 0, nw >> 1, gmask, S, C2, zb, cb); break;
                    default: aux_bound_segment<4, 1>(q4, q4 + 1, (uint32_t)npad, 0, nw >> 1, gmask, S, C2, zb, cb); break;
                }
                double z_true = 0., z_cap = 0.;
                const int vcap = 8 * g0 + 4 * t0 + HLLB_NV;
                for (int v = 1; v < 64; ++v) {
                    z_true += std::ldexp((double)want[v], -v);
                    z_cap += std::ldexp((double)want[v], -std::min(v, vcap));
                }
                ++bound_done;
                if (cb != (float)want[0] || !(zb >= z_true * (1 - 1e-6)) || !(zb <= z_cap * (1 + 1e-5))) {
                    if (bad < 10) printf("bound-check %d p_aux=%d g0=%d t0=%d range %d..%d: z %g (true %g, capped %g) empty %g (true %u)\n", it, p_aux, g0, t0, klo, khi, zb, z_true, z_cap, cb, want[0]);
                    ++bad;
                }
            }
            for (int v = 0; v < nbins; ++v)
                if (hcol[(size_t)v * 64] != want[v]) {
                    if (bad < 10) printf("hist-check %d p_aux=%d g0=%d range %d..%d: bin %d is %u, want %u\n", it, p_aux, g0, klo, khi, v, hcol[(size_t)v * 64], want[v]);
                    ++bad;
                }
        }
    }
    if (bad) printf("hist-check FAILED (%d bins of %d pairs)\n", bad, done);
    else printf("hist-check: identical to the definition on %d pairs, bound sums valid on %d\n", done, bound_done);
    return bad ? 1 : 0;
}

int main(int argc, char** argv) {
    if (argc == 4 && std::string(argv[1]) == "hist-check") return hist_check(strtoull(argv[2], nullptr, 10), atoi(argv[3]));
    if (argc < 5) { fprintf(stderr, "usage: emul_filter_hll in.bin out.bin twopass|onepass|bytes n_shards [grid]\n"); return 2; }
    const bool twopass = std::string(argv[3]) == "twopass";      // the default plane filter: bound, then exact decision
    const bool planes = std::string(argv[3]) == "onepass" || twopass;
    const int n_shards = atoi(argv[4]);
    const unsigned fgrid = argc > 5 ? (unsigned)atoi(argv[5]) : 3u;
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    int32_t hdr[5];
    double tau;
    float zs;
    rd(f, hdr, 5);
    rd(f, &tau, 1);
    rd(f, &zs, 1);
    const int n = hdr[0], p_aux = hdr[1], an = hdr[2], order_n = hdr[3], zeros = hdr[4];
    const size_t m_aux = (size_t)1 << p_aux;
    std::vector<unsigned long long> e((size_t)n);
    std::vector<uint8_t> aux((size_t)n * m_aux + 32);      // k_aux_planes_quad reads 32 bytes per word
    rd(f, e.data(), e.size());
    rd(f, aux.data(), (size_t)n * m_aux);
    fclose(f);

    const long long npad = ((long long)n + TILE - 1) / TILE * TILE;
    const int nrb = (n + TILE - 1) / TILE;
    std::vector<int32_t> lo(n), hi(n), tile_nt(nrb + 1), tile_prefix(nrb + 1), tile_cb0(nrb), order(n);
    std::iota(order.begin(), order.end(), 0);              // the input is in sorted order already
    std::vector<unsigned long long> rb_pairs(nrb), meta(M_WORDS, 0);
    emul::launch((unsigned)((n + 255) / 256), 256, [&] { k_cb_bounds(e.data(), n, zeros, tau, lo.data(), hi.data()); });
    emul::launch((unsigned)((nrb + 1 + 3) / 4), 128, [&] {
        k_rowblock_span(lo.data(), hi.data(), n, nrb, tile_nt.data(), tile_cb0.data(), rb_pairs.data(), meta.data());
    });
    std::exclusive_scan(tile_nt.begin(), tile_nt.end(), tile_prefix.begin(), 0);
    const long long tile_cap = std::max<long long>(1, (long long)nrb * (nrb + 1) / 2);
    std::vector<int2> tile_rc((size_t)tile_cap);
    emul::launch((unsigned)((nrb + 3) / 4), 128, [&] {
        k_tile_table(tile_prefix.data(), tile_cb0.data(), nrb, tile_cap, tile_rc.data(), meta.data());
    });
    // load-time layouts: auxT[word][genome] (k_aux_transpose, restated here), bit planes and ranges by the kernels
    const int row_words = (int)(m_aux / 4);
    std::vector<uint32_t> auxT((size_t)row_words * npad, 0u);
    for (int g = 0; g < n; ++g)
        for (int j = 0; j < row_words; ++j) std::memcpy(&auxT[(size_t)j * npad + g], &aux[(size_t)g * m_aux + 4 * (size_t)j], 4);
    const int nw = (int)(m_aux >> 5);
    std::vector<uint32_t> auxP(planes ? (size_t)6 * nw * npad : 1, 0u);
    std::vector<uint16_t> agrange((size_t)npad, 0);
    std::vector<AuxTail> atail((size_t)npad, AuxTail{});
    if (planes) {
        emul::launch(2, 256, [&] { k_aux_planes_quad(aux.data(), order.data(), n, npad, p_aux, auxP.data()); });
        emul::launch((unsigned)(((long long)n * 32 + 255) / 256), 256,
                     [&] { k_aux_range(aux.data(), order.data(), n, p_aux, agrange.data(), atail.data()); });
    }
    const unsigned long long cap = 1ull << 22;
    std::vector<uint2> pairs((size_t)cap), cand(twopass ? (size_t)cap : 1), all_pairs;
    unsigned long long cand_total = 0;
    for (int shard = 0; shard < n_shards; ++shard) {
        meta[M_PAIRS] = meta[M_UNIT] = meta[M_CAND] = 0;
        const TileWalk tw{tile_rc.data(), meta.data(), tile_cap, shard, n_shards, 0, INT32_MAX};
        if (twopass) {
            emul::launch(fgrid, 64, [&] {
                if (an) k_tile_filter_hll_bound<1>(auxP.data(), agrange.data(), atail.data(), npad, p_aux, tw, lo.data(), hi.data(), n, e.data(), (float)tau, zs, order_n, cand.data(), meta.data() + M_CAND, cap, meta.data() + M_UNIT);
                else k_tile_filter_hll_bound<0>(auxP.data(), agrange.data(), atail.data(), npad, p_aux, tw, lo.data(), hi.data(), n, e.data(), (float)tau, zs, order_n, cand.data(), meta.data() + M_CAND, cap, meta.data() + M_UNIT);
            });
            if (meta[M_CAND] > cap) { fprintf(stderr, "list overflow\n"); return 3; }
            cand_total += meta[M_CAND];
            emul::launch(fgrid, 64, [&] {
                if (an) k_hll_verify<1>(auxP.data(), agrange.data(), auxT.data(), npad, p_aux, cand.data(), meta.data() + M_CAND, cap, n, e.data(), tau, zs, order_n, pairs.data(), meta.data() + M_PAIRS, cap);
                else k_hll_verify<0>(auxP.data(), agrange.data(), auxT.data(), npad, p_aux, cand.data(), meta.data() + M_CAND, cap, n, e.data(), tau, zs, order_n, pairs.data(), meta.data() + M_PAIRS, cap);
            });
        } else {
            emul::launch(fgrid, 64, [&] {
                if (planes) {
                    if (an) k_tile_filter_hll_planes<1>(auxP.data(), agrange.data(), auxT.data(), npad, p_aux, tw, lo.data(), hi.data(), n, e.data(), tau, zs, order_n, pairs.data(), meta.data() + M_PAIRS, cap, meta.data() + M_UNIT);
                    else k_tile_filter_hll_planes<0>(auxP.data(), agrange.data(), auxT.data(), npad, p_aux, tw, lo.data(), hi.data(), n, e.data(), tau, zs, order_n, pairs.data(), meta.data() + M_PAIRS, cap, meta.data() + M_UNIT);
                } else {
                    if (an) k_tile_filter_hll<1>(auxT.data(), npad, p_aux, tw, lo.data(), hi.data(), n, e.data(), tau, zs, order_n, pairs.data(), meta.data() + M_PAIRS, cap, meta.data() + M_UNIT);
                    else k_tile_filter_hll<0>(auxT.data(), npad, p_aux, tw, lo.data(), hi.data(), n, e.data(), tau, zs, order_n, pairs.data(), meta.data() + M_PAIRS, cap, meta.data() + M_UNIT);
                }
            });
        }
        if (meta[M_PAIRS] > cap) { fprintf(stderr, "list overflow\n"); return 3; }
        all_pairs.insert(all_pairs.end(), pairs.begin(), pairs.begin() + (long long)meta[M_PAIRS]);
    }
    std::sort(all_pairs.begin(), all_pairs.end(), [](uint2 a, uint2 b) { return a.x != b.x ? a.x < b.x : a.y < b.y; });
    f = fopen(argv[2], "wb");
    if (!f) { perror(argv[2]); return 2; }
    const long long out_hdr[2] = {(long long)meta[M_PAIRS_CB], (long long)all_pairs.size()};
    wr(f, out_hdr, 2);
    wr(f, all_pairs.data(), all_pairs.size());
    fclose(f);
    printf("n=%d p_aux=%d %s %s P_cb=%llu candidates=%llu pairs=%zu\n", n, p_aux, twopass ? "twopass" : planes ? "onepass" : "bytes", an ? "hll_an" : "hll_a", meta[M_PAIRS_CB], cand_total, all_pairs.size());
    return 0;
}
