// emul_run.cpp — a whole smh_a selection on the CPU through cuda_emul.h, launch by launch as the library queues
// them (csrc/selb200.cu: load_chunk, load_end, selb200_run), from the same .inl sources the GPU build compiles:
//   load : k_max_byte, k_pair_hist<SrcSelf> (per-genome histograms), k_genome_cards (Ertl MLE + value range),
//          k_planes_from_bytes, [host: sort by cardinality], k_sorted_prep, k_gather_rows
//   run  : k_cb_bounds, k_rowblock_span, scan, k_tile_table, k_smh_sigkeys, scan, k_smh_scatter, k_smh_join_expand, k_smh_join
//          (or, with a fourth argument
//          "tiles": k_smh_signatures, k_tile_filter_smh, k_smh_verify),
//          k_pair_hist_planes (+ k_pair_hist<SrcWide> for wide pairs), k_estimate_screen, k_estimate_emit,
//          k_rowsort_count, scan, k_rowsort_scatter, k_rowsort_rank
// Input (file): registers and SuperMinHash sketches in FILE-LIST order, tau, band shape.
// Output (file): cardinalities, sorted order, stage counts, the final (i, k, J) list in print order, near-tau count.
// tests/test_emul_run.py holds all of it against the oracle (oracle/oracle.cpp).  Test infrastructure.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <numeric>

#include <string>

#define SELB_EMUL 1
#include "cuda_emul.h"
#include "../../cuda_selection_criteria_b200/csrc/estimators.cuh"
#include "../../cuda_selection_criteria_b200/csrc/hostpack.cpp"     // host packer of the packed upload (plain C++)

constexpr int TILE = 128;          // as in csrc/selb200.cu
constexpr int SIG_CHUNK = 8;

#include "../../cuda_selection_criteria_b200/csrc/kernels/helpers.inl"
#include "../../cuda_selection_criteria_b200/csrc/kernels/union_bytes.inl"
#include "../../cuda_selection_criteria_b200/csrc/kernels/union_planes.inl"
#include "../../cuda_selection_criteria_b200/csrc/kernels/load_kernels.inl"
#include "../../cuda_selection_criteria_b200/csrc/kernels/tiles.inl"
#include "../../cuda_selection_criteria_b200/csrc/kernels/filter_smh.inl"
#include "../../cuda_selection_criteria_b200/csrc/kernels/estimate_sort.inl"

template <class T> static void rd(FILE* f, T* p, size_t n) { if (fread(p, sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } }
template <class T> static void wr(FILE* f, const T* p, size_t n) { if (fwrite(p, sizeof(T), n, f) != n) { fprintf(stderr, "short write\n"); exit(2); } }

#include <chrono>
static void lap(const char* what) {
    static auto t0 = std::chrono::steady_clock::now();
    const auto t1 = std::chrono::steady_clock::now();
    if (getenv("EMUL_TIMING")) fprintf(stderr, "%-28s %.2f s\n", what, std::chrono::duration<double>(t1 - t0).count());
    t0 = t1;
}

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: emul_run in.bin out.bin [subsets|planes [tiles]]\n"); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    int32_t hdr[5];
    double tau;
    rd(f, hdr, 5);
    rd(f, &tau, 1);
    const int n = hdr[0], p = hdr[1], m_aux = hdr[2], n_rows = hdr[3], n_bands = hdr[4];
    const size_t m = (size_t)1 << p;
    std::vector<uint8_t> regs((size_t)n * m);
    std::vector<uint64_t> smh((size_t)n * m_aux);
    rd(f, regs.data(), regs.size());
    rd(f, smh.data(), smh.size());
    fclose(f);

    // ------------------------------------------------------------------ packed upload (load_host_packed / unpack_pieces):
    // the registers go through the host packer and the three unpack kernels, pieces of 50 rows (a short last one), and
    // everything below works on what came out
    {
        const long long piece_rows = 50;
        const selb::Nib4Piece P = selb::nib4_piece(piece_rows, m);
        const unsigned n_pieces = (unsigned)((n + piece_rows - 1) / piece_rows);
        std::vector<uint8_t> pieces((size_t)n_pieces * P.bytes, 0xCD), back((size_t)n * m, 0xEE);
        for (unsigned pi = 0; pi < n_pieces; ++pi) {
            const long long rows = std::min<long long>(piece_rows, n - (long long)pi * piece_rows);
            selb::nib4_pack_piece(regs.data() + (size_t)pi * piece_rows * m, rows, m, pieces.data() + (size_t)pi * P.bytes, 1);
        }
        const Nib4Pieces a{pieces.data(), P.bytes, piece_rows, (long long)n, p};
        emul::launch2(2, n_pieces, 256, [&] { k_unpack_nib4(a, back.data()); });
        emul::launch2(1, n_pieces, 256, [&] { k_apply_nib4_exc(a, back.data()); });
        emul::launch2(selb::NIB4_RAW_CAP, n_pieces, 256, [&] { k_apply_nib4_raw(a, back.data()); });
        if (back != regs) { fprintf(stderr, "packed upload does not reproduce the register bytes\n"); return 5; }
        regs.swap(back);
    }
    lap("packed upload");
    // ------------------------------------------------------------------ load (load_chunk / load_end)
    uint32_t flags[4] = {0, 0, 0, 0};          // max primary register, max aux register, tie flag
    emul::launch(2, 256, [&] { k_max_byte(reinterpret_cast<const uint4*>(regs.data()), (size_t)n * m / 16, flags); });
    if (flags[0] > (uint32_t)(64 - p + 1)) { fprintf(stderr, "register value %u too large\n", flags[0]); return 4; }
    std::vector<uint32_t> ghist((size_t)n * 64, 0xDEADBEEFu);
    {
        SrcSelf src{0, n, flags, (uint32_t)(64 - p + 1)};
        EpiWriteHist epi{ghist.data()};
        emul::launch(3, 64, [&] { k_pair_hist<52, SrcSelf, EpiWriteHist>(regs.data(), m, m, src, epi); });
    }
    lap("max_byte + genome hist");
    std::vector<double> cards(n);
    std::vector<uint16_t> grange(n);
    emul::launch((unsigned)((n + 127) / 128), 128, [&] {
        k_genome_cards(ghist.data(), nullptr, n, p, cards.data(), flags, (uint32_t)(64 - p + 1), grange.data());
    });
    const int chunk_regs = (int)std::min<size_t>(m, PL_CHUNK_REGS);
    std::vector<uint32_t> planes((size_t)n * 6 * (m >> 5), 0xA5A5A5A5u);
    std::vector<uint32_t> gtop((size_t)n * 8, 0u);          // per-eighth maxima, as the library's load step writes them
    emul::launch(2, 256, [&] { k_planes_from_bytes(regs.data(), n, m, chunk_regs, planes.data(), m >= 4096 ? gtop.data() : nullptr); });
    lap("cards + planes");
    // device radix sort of (cardinality, index): distinct keys have one order; ties take the host path in the library
    std::vector<int32_t> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cards[a] < cards[b]; });
    std::vector<double> cards_sorted(n);
    for (int i = 0; i < n; ++i) cards_sorted[i] = cards[order[i]];
    std::vector<unsigned long long> e(n);
    emul::launch((unsigned)((n + 255) / 256), 256, [&] { k_sorted_prep(cards_sorted.data(), n, e.data(), flags + 2); });
    std::vector<uint64_t> aux_sorted((size_t)n * m_aux);
    emul::launch(2, 256, [&] {
        k_gather_rows(reinterpret_cast<const uint32_t*>(smh.data()), order.data(), n, m_aux * 2, reinterpret_cast<uint32_t*>(aux_sorted.data()));
    });

    lap("sort + prep + gather");
    // ------------------------------------------------------------------ run (selb200_run, smh_a)
    int zeros = 0;
    while (zeros < n && e[zeros] == 0) ++zeros;
    const long long npad = ((long long)n + TILE - 1) / TILE * TILE;
    const int nrb = (n + TILE - 1) / TILE;
    const int n_words = (n_bands + 1) / 2;
    std::vector<int32_t> lo(n), hi(n), tile_nt(nrb + 1), tile_prefix(nrb + 1), tile_cb0(nrb);
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
    lap("bounds + tiles");
    const unsigned long long cap = 1ull << 20;
    std::vector<uint2> cand((size_t)cap), pairs((size_t)cap);
    const bool tiles_filter = argc > 4 && std::string(argv[4]) == "tiles";       // SELB200_SMHFILTER=tiles
    if (!tiles_filter) {
        // the default: equality join — k_smh_sigkeys (keys, ranks, bucket counts) -> exclusive scan (cub on the device) ->
        // k_smh_scatter -> k_smh_join_expand -> k_smh_join
        const long long nk = (long long)n * n_bands;
        const int sbits = 16;
        const long long nbk = (long long)n_bands << sbits;
        std::vector<uint32_t> keys((size_t)nk, 0xDEADBEEFu), rank((size_t)nk, 0xDEADBEEFu), members((size_t)nk, 0xDEADBEEFu),
            sigG((size_t)n * n_words, 0xDEADBEEFu), bcnt((size_t)nbk + 1, 0u), boff((size_t)nbk + 1);
        emul::launch(2, 256, [&] {
            k_smh_sigkeys(aux_sorted.data(), n, m_aux, n_rows, n_bands, n_words, sbits, keys.data(), rank.data(), bcnt.data(), sigG.data());
        });
        std::exclusive_scan(bcnt.begin(), bcnt.end(), boff.begin(), 0u);
        emul::launch(2, 256, [&] { k_smh_scatter(keys.data(), rank.data(), boff.data(), nk, n_bands, members.data()); });
        // two passes over two parts of the elements, as after a split pass: expand the bucket members into items, one thread per item
        const long long mid = nk / 3;
        const unsigned long long item_cap = 1ull << 22;
        std::vector<uint4> items((size_t)item_cap);
        for (int part = 0; part < 2; ++part) {
            const long long s0 = part ? mid : 0, s1 = part ? nk : mid;
            meta[M_ITEMS] = 0;
            emul::launch(2, 256, [&] {
                k_smh_join_expand(keys.data(), boff.data(), members.data(), s0, s1, n_bands, sbits, lo.data(), hi.data(), items.data(),
                                  meta.data() + M_ITEMS, item_cap);
            });
            if (meta[M_ITEMS] > item_cap) { fprintf(stderr, "item list overflow\n"); return 3; }
            emul::launch(part ? 3 : 2, 256, [&] {
                k_smh_join(items.data(), meta.data() + M_ITEMS, item_cap, sigG.data(), n_words, sbits, aux_sorted.data(), m_aux, n_rows, n_bands,
                           pairs.data(), meta.data() + M_PAIRS, cap, meta.data() + M_CAND, meta.data() + M_ITEMS_MAX);
            });
        }
    } else {
    std::vector<uint32_t> sigR((size_t)n_words * npad), sigC((size_t)n_words * npad);
    emul::launch(2, 256, [&] { k_smh_signatures(aux_sorted.data(), n, npad, m_aux, n_rows, n_bands, sigR.data(), sigC.data()); });
    const TileWalk tw{tile_rc.data(), meta.data(), tile_cap, 0, 1, 0, INT32_MAX};
    emul::launch(3, 256, [&] {
        k_tile_filter_smh(sigR.data(), sigC.data(), npad, n_words, tw, lo.data(), hi.data(), n, cand.data(), meta.data() + M_CAND, cap);
    });
    emul::launch(2, 256, [&] {
        k_smh_verify(aux_sorted.data(), sigR.data(), npad, m_aux, n_rows, n_bands, cand.data(), meta.data() + M_CAND, cap,
                     pairs.data(), meta.data() + M_PAIRS, cap);
    });
    }
    if (meta[M_CAND] > cap || meta[M_PAIRS] > cap) { fprintf(stderr, "list overflow\n"); return 3; }
    lap("signatures + filter + verify");
    const long long np = (long long)meta[M_PAIRS];
    // K5: plane union over the pair list (rows through `order`), wide pairs through the byte kernel
    std::vector<uint32_t> hist((size_t)std::max<long long>(np, 1) * 64, 0xDEADBEEFu), wide((size_t)std::max<long long>(np, 1));
    // the plane kernel stamps the pairs it hands to the wide list; the library runs the byte kernel and the estimate of
    // those pairs on a second stream, next to the estimate of everything else, which skips stamped pairs.  Here: the main
    // estimate FIRST, while the wide pairs' rows still hold poison, then the byte kernel, then the wide pairs' estimate
    std::vector<uint32_t> wflag((size_t)std::max<long long>(np, 1), 6u);      // stamps of an older pass
    const uint32_t wepoch = 7u;
    SrcWide wsrc{pairs.data(), order.data(), wide.data(), meta.data() + M_WIDE};
    EpiWriteHist epi_w{hist.data()};
    {
        SrcPairs src{pairs.data(), order.data(), (long long)cap, meta.data() + M_PAIRS};
        EpiWriteHist epi{hist.data()};
        const bool subsets = argc > 3 && std::string(argv[3]) == "subsets";     // SELB200_UNION=subsets
        emul::launch(3, 32, [&] {
            if (subsets)
                k_pair_hist_planes<EpiSubsets<EpiWriteHist>>(planes.data(), m, chunk_regs, grange.data(), src, EpiSubsets<EpiWriteHist>{epi},
                                                             wide.data(), meta.data() + M_WIDE, meta.data() + M_BATCH, wflag.data(), wepoch,
                                                             m >= 16384 ? gtop.data() : nullptr);
            else
                k_pair_hist_planes<EpiWriteHist>(planes.data(), m, chunk_regs, grange.data(), src, epi, wide.data(), meta.data() + M_WIDE,
                                                 meta.data() + M_BATCH, wflag.data(), wepoch);
        });
    }
    if (meta[M_KERR]) { fprintf(stderr, "union kernel error word %llx\n", meta[M_KERR]); return 5; }
    lap("union");
    // K6
    const unsigned long long out_cap = 1ull << 20, near_cap = 1ull << 16;
    std::vector<uint64_t> out_keys((size_t)out_cap), near_keys((size_t)near_cap);
    std::vector<double> out_j((size_t)out_cap), near_j((size_t)near_cap);
    // the filtered list goes through k_estimate_emit directly; with "screen" as fifth argument through the two-step form the
    // library uses for unfiltered (criterion cb) lists
    if (argc > 5 && std::string(argv[5]) == "screen") {
        std::vector<uint32_t> surv((size_t)cap, 0xDEADBEEFu);
        emul::launch(3, 128, [&] {
            k_estimate_screen(hist.data(), pairs.data(), meta.data() + M_PAIRS, cap, e.data(), p, tau, surv.data(), meta.data() + M_SURV,
                              wflag.data(), wepoch);
        });
        emul::launch(2, 128, [&] {
            k_estimate_emit(hist.data(), pairs.data(), surv.data(), meta.data() + M_SURV, cap, e.data(), p, tau, out_keys.data(), out_j.data(),
                            meta.data() + M_OUT, out_cap, near_keys.data(), near_j.data(), meta.data() + M_NEAR, near_cap);
        });
        if (meta[M_SURV] > meta[M_PAIRS] || (np > 100 && meta[M_SURV] == meta[M_PAIRS])) {
            fprintf(stderr, "estimate screen kept %llu of %llu pairs\n", meta[M_SURV], meta[M_PAIRS]);
            return 6;
        }
    } else {
        emul::launch(3, 128, [&] {
            k_estimate_emit(hist.data(), pairs.data(), nullptr, meta.data() + M_PAIRS, cap, e.data(), p, tau, out_keys.data(), out_j.data(),
                            meta.data() + M_OUT, out_cap, near_keys.data(), near_j.data(), meta.data() + M_NEAR, near_cap, wflag.data(), wepoch);
        });
    }
    // the wide pairs: byte kernel, then their estimate from the wide list
    emul::launch(2, 64, [&] { k_pair_hist<52, SrcWide, EpiWriteHist>(regs.data(), m, m, wsrc, epi_w); });
    emul::launch(1, 128, [&] {
        k_estimate_emit(hist.data(), pairs.data(), wide.data(), meta.data() + M_WIDE, cap, e.data(), p, tau, out_keys.data(), out_j.data(),
                        meta.data() + M_OUT, out_cap, near_keys.data(), near_j.data(), meta.data() + M_NEAR, near_cap);
    });
    lap("estimate");
    const long long cnt = (long long)meta[M_OUT];
    // K7: sparse-output print order
    std::vector<uint64_t> fin_keys((size_t)std::max<long long>(cnt, 1));
    std::vector<double> fin_j((size_t)std::max<long long>(cnt, 1));
    if (cnt > 1) {
        std::vector<int32_t> row_cnt(n + 1, 0), row_off(n + 1, 0);
        std::vector<uint64_t> tkeys((size_t)cnt);
        std::vector<double> tj((size_t)cnt);
        const unsigned grid = (unsigned)((cnt + 255) / 256);
        emul::launch(grid, 256, [&] { k_rowsort_count(out_keys.data(), cnt, row_cnt.data()); });
        std::exclusive_scan(row_cnt.begin(), row_cnt.end(), row_off.begin(), 0);
        emul::launch(grid, 256, [&] { k_rowsort_scatter(out_keys.data(), out_j.data(), cnt, row_cnt.data(), row_off.data(), tkeys.data(), tj.data()); });
        emul::launch(grid, 256, [&] { k_rowsort_rank(tkeys.data(), tj.data(), cnt, row_off.data(), fin_keys.data(), fin_j.data()); });
    } else if (cnt == 1) {
        fin_keys[0] = out_keys[0];
        fin_j[0] = out_j[0];
    }

    lap("sort");
    f = fopen(argv[2], "wb");
    if (!f) { perror(argv[2]); return 2; }
    const long long out_hdr[8] = {(long long)meta[M_PAIRS_CB], (long long)meta[M_CAND], np, cnt, (long long)meta[M_NEAR],
                                  (long long)meta[M_WIDE], (long long)flags[2], 0};
    wr(f, out_hdr, 8);
    wr(f, cards.data(), cards.size());
    wr(f, order.data(), order.size());
    wr(f, fin_keys.data(), (size_t)cnt);
    wr(f, fin_j.data(), (size_t)cnt);
    fclose(f);
    printf("n=%d P_cb=%llu candidates=%llu P_aux=%lld wide=%llu P_out=%lld near=%llu\n", n, meta[M_PAIRS_CB], meta[M_CAND], np,
           meta[M_WIDE], cnt, meta[M_NEAR]);
    return 0;
}
