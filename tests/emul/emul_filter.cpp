// emul_filter.cpp — runs the CB + smh_a chain of a selection run on the CPU through cuda_emul.h, from the same .inl
// sources the GPU build compiles, launch by launch as selb200_run queues them:
//   k_cb_bounds -> k_rowblock_span -> exclusive scan -> k_tile_table        (kernels/tiles.inl)
//   k_smh_signatures -> k_tile_filter_smh -> k_smh_verify                   (kernels/filter_smh.inl)
// Input (file): sorted truncated cardinalities e[], SuperMinHash sketches in sorted order, band shape, tau.
// Output (file): lo/hi of every row, the number of pairs inside the CB band, the candidates' count and the
// surviving pair list — tests/test_emul_filter.py holds them against the oracle's CB and smh_a decisions.
// Test infrastructure; exit code 0 = ran to completion.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <numeric>

#define SELB_EMUL 1
#include "cuda_emul.h"
#include "../../cuda_selection_criteria_b200/csrc/estimators.cuh"

constexpr int TILE = 128;          // as in csrc/selb200.cu
constexpr int SIG_CHUNK = 8;

#include "../../cuda_selection_criteria_b200/csrc/kernels/helpers.inl"
#include "../../cuda_selection_criteria_b200/csrc/kernels/tiles.inl"
#include "../../cuda_selection_criteria_b200/csrc/kernels/filter_smh.inl"

template <class T> static void rd(FILE* f, T* p, size_t n) { if (fread(p, sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } }
template <class T> static void wr(FILE* f, const T* p, size_t n) { if (fwrite(p, sizeof(T), n, f) != n) { fprintf(stderr, "short write\n"); exit(2); } }

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: emul_filter in.bin out.bin n_shards [filter_grid]\n"); return 2; }
    const int n_shards = atoi(argv[3]);
    const unsigned fgrid = argc > 4 ? (unsigned)atoi(argv[4]) : 3u;
    const bool join = argc > 5 && std::string(argv[5]) == "join";      // equality join instead of tile filter + verify
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    int32_t hdr[5];
    double tau;
    rd(f, hdr, 5);
    rd(f, &tau, 1);
    const int n = hdr[0], m_aux = hdr[1], n_rows = hdr[2], n_bands = hdr[3], zeros = hdr[4];
    std::vector<unsigned long long> e((size_t)n);
    std::vector<uint64_t> aux((size_t)n * m_aux);
    rd(f, e.data(), e.size());
    rd(f, aux.data(), aux.size());
    fclose(f);

    const long long npad = ((long long)n + TILE - 1) / TILE * TILE;
    const int nrb = (n + TILE - 1) / TILE;
    const int n_words = (n_bands + 1) / 2;
    std::vector<int32_t> lo(n), hi(n), tile_nt(nrb + 1), tile_prefix(nrb + 1), tile_cb0(nrb);
    std::vector<unsigned long long> rb_pairs(nrb), meta(M_WORDS, 0);
    emul::launch((unsigned)((n + 255) / 256), 256, [&] { k_cb_bounds(e.data(), n, zeros, tau, lo.data(), hi.data()); });
    emul::launch((unsigned)((nrb + 1 + 3) / 4), 128, [&] {
        k_rowblock_span(lo.data(), hi.data(), n, nrb, tile_nt.data(), tile_cb0.data(), rb_pairs.data(), meta.data());
    });
    std::exclusive_scan(tile_nt.begin(), tile_nt.end(), tile_prefix.begin(), 0);      // cub::DeviceScan::ExclusiveSum
    const long long tile_cap = std::max<long long>(1, (long long)nrb * (nrb + 1) / 2);
    std::vector<int2> tile_rc((size_t)tile_cap);
    emul::launch((unsigned)((nrb + 3) / 4), 128, [&] {
        k_tile_table(tile_prefix.data(), tile_cb0.data(), nrb, tile_cap, tile_rc.data(), meta.data());
    });
    std::vector<uint32_t> sigR((size_t)n_words * npad, 0xDEADBEEFu), sigC((size_t)n_words * npad, 0xDEADBEEFu);
    emul::launch(2, 256, [&] {
        k_smh_signatures(aux.data(), n, npad, m_aux, n_rows, n_bands, sigR.data(), sigC.data());
    });
    const unsigned long long cap = 1ull << 22;
    std::vector<uint2> cand((size_t)cap), pairs((size_t)cap);
    std::vector<uint2> all_pairs;
    unsigned long long cand_total = 0;
    // equality join (the default of the library below four shards): keys + bucket counts, exclusive scan (cub on the
    // device), scatter; then one expansion + walk per shard.  sbits < 16 (coarser buckets, as for more than 256 bands) when
    // the sixth argument says so
    const long long nk = (long long)n * n_bands;
    const int sbits = argc > 6 ? atoi(argv[6]) : 16;
    const long long nbk = (long long)n_bands << sbits;
    std::vector<uint32_t> jkeys, members, sigG, boff;
    if (join) {
        std::vector<uint32_t> rank((size_t)nk, 0xDEADBEEFu), bcnt((size_t)nbk + 1, 0u);
        jkeys.assign((size_t)nk, 0xDEADBEEFu);
        members.assign((size_t)nk, 0xDEADBEEFu);
        sigG.assign((size_t)n * n_words, 0xDEADBEEFu);
        boff.resize((size_t)nbk + 1);
        emul::launch(2, 256, [&] {
            k_smh_sigkeys(aux.data(), n, m_aux, n_rows, n_bands, n_words, sbits, jkeys.data(), rank.data(), bcnt.data(), sigG.data());
        });
        std::exclusive_scan(bcnt.begin(), bcnt.end(), boff.begin(), 0u);
        emul::launch(2, 256, [&] { k_smh_scatter(jkeys.data(), rank.data(), boff.data(), nk, n_bands, members.data()); });
    }
    for (int shard = 0; shard < n_shards; ++shard) {       // every shard, one after the other
        meta[M_CAND] = meta[M_PAIRS] = 0;
        if (join) {
            const unsigned long long item_cap = 1ull << 22;
            std::vector<uint4> items((size_t)item_cap);
            meta[M_ITEMS] = 0;
            emul::launch(2, 256, [&] {
                k_smh_join_expand(jkeys.data(), boff.data(), members.data(), 0, nk, n_bands, sbits, lo.data(), hi.data(), items.data(),
                                  meta.data() + M_ITEMS, item_cap, shard, n_shards, (long long)n);
            });
            if (meta[M_ITEMS] > item_cap) { fprintf(stderr, "item list overflow\n"); return 3; }
            emul::launch(fgrid, 256, [&] {
                k_smh_join(items.data(), meta.data() + M_ITEMS, item_cap, sigG.data(), n_words, sbits, aux.data(), m_aux, n_rows, n_bands,
                           pairs.data(), meta.data() + M_PAIRS, cap, meta.data() + M_CAND, meta.data() + M_ITEMS_MAX);
            });
        } else {
        const TileWalk tw{tile_rc.data(), meta.data(), tile_cap, shard, n_shards, 0, INT32_MAX};
        emul::launch(fgrid, 256, [&] {
            k_tile_filter_smh(sigR.data(), sigC.data(), npad, n_words, tw, lo.data(), hi.data(), n, cand.data(),
                              meta.data() + M_CAND, cap);
        });
        emul::launch(2, 256, [&] {
            k_smh_verify(aux.data(), sigR.data(), npad, m_aux, n_rows, n_bands, cand.data(), meta.data() + M_CAND, cap,
                         pairs.data(), meta.data() + M_PAIRS, cap);
        });
        }
        if (meta[M_CAND] > cap || meta[M_PAIRS] > cap) { fprintf(stderr, "list overflow\n"); return 3; }
        cand_total += meta[M_CAND];
        all_pairs.insert(all_pairs.end(), pairs.begin(), pairs.begin() + (long long)meta[M_PAIRS]);
    }
    // shards must not overlap: a pair emitted twice would show up as a duplicate below
    std::sort(all_pairs.begin(), all_pairs.end(), [](uint2 a, uint2 b) { return a.x != b.x ? a.x < b.x : a.y < b.y; });

    f = fopen(argv[2], "wb");
    if (!f) { perror(argv[2]); return 2; }
    const long long out_hdr[4] = {(long long)meta[M_PAIRS_CB], (long long)meta[M_TILES], (long long)cand_total, (long long)all_pairs.size()};
    wr(f, out_hdr, 4);
    wr(f, lo.data(), lo.size());
    wr(f, hi.data(), hi.size());
    wr(f, all_pairs.data(), all_pairs.size());
    fclose(f);
    printf("n=%d bands=%dx%d tiles=%llu P_cb=%llu candidates=%llu pairs=%zu\n", n, n_bands, n_rows, meta[M_TILES], meta[M_PAIRS_CB],
           cand_total, all_pairs.size());
    return 0;
}
