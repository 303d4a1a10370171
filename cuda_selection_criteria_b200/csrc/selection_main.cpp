// selection_main.cpp — C++ host driver over the C-ABI (include/selb200.h).
//
// Drop-in for the reference CLIs (same flag letters, same defaults, same stdout lines):
//   ./selection       -l list -t threads -a aux_bytes -h tau -c {smh_a|hll_a|hll_an|cb} [-b block]
//       mirrors src/selection.cpp:70-303: file list (:36-63), sketches `P.hll` +
//       `P.smh<a/8>` | `P.hll_<ctz(a)>` (:125,138-139,231,245-246), output lines
//       `P_i P_k std::to_string(J)` in (sorted row, k) order (:288,297-300), invalid -c
//       message on stdout and exit 0 (:292-294).
//   ./selection_cuda  -l list -b block -a aux_bytes -h tau   (built with -DSELB_CUDA_DRIVER)
//       mirrors src/selection_cuda.cpp:59-189: criterion fixed to smh_a (-c accepted and
//       ignored, :68-88), band search that keeps (1,1) when nothing qualifies (:119-128),
//       similarity printed through operator<<(float) (:184-186).
// Extra, not in the reference: `-c cb` (CB only), `-g` prints run statistics to stderr.
// The device work is entirely behind selb200_load_host / selb200_run; this file only parses
// flags, gunzips sketches (zlib, OpenMP over files) and formats lines.
#include <getopt.h>
#include <chrono>
#include <omp.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/selb200.h"

namespace {

// sketch/include/sketch/hll.h:1126-1143 (read): u32[4] header, u32 np, f64 value, 2^np registers
struct HllFile {
    uint32_t hdr[4];
    uint32_t np;
    double value;
    std::vector<uint8_t> core;
};

void gz_read_exact(gzFile fp, void* dst, size_t len, const std::string& path) {
    if (static_cast<uint64_t>(gzread(fp, dst, (unsigned)len)) != len) {
        gzclose(fp);
        throw std::runtime_error("Error reading from file " + path);
    }
}

HllFile read_hll(const std::string& path) {
    gzFile fp = gzopen(path.c_str(), "rb");
    if (fp == nullptr) throw std::runtime_error(std::string("Could not open file at '") + path + "' for reading");
    HllFile h;
    gz_read_exact(fp, h.hdr, sizeof h.hdr, path);
    gz_read_exact(fp, &h.np, sizeof h.np, path);
    gz_read_exact(fp, &h.value, sizeof h.value, path);
    if (h.np > 30) { gzclose(fp); throw std::runtime_error("Error reading from file " + path); }
    h.core.resize((size_t)1 << h.np);
    gz_read_exact(fp, h.core.data(), h.core.size(), path);
    gzclose(fp);
    return h;
}

// same reader, registers decoded straight into `dst` (a pinned staging slot)
void read_hll_into(const std::string& path, uint32_t expect_np, uint8_t* dst, double* value) {
    gzFile fp = gzopen(path.c_str(), "rb");
    if (fp == nullptr) throw std::runtime_error(std::string("Could not open file at '") + path + "' for reading");
    uint32_t hdr[4], np = 0;
    gz_read_exact(fp, hdr, sizeof hdr, path);
    gz_read_exact(fp, &np, sizeof np, path);
    gz_read_exact(fp, value, sizeof *value, path);
    if (np != expect_np) { gzclose(fp); throw std::runtime_error(path + ": HLL precision differs from the expected one"); }
    if (hdr[1] != 2 || hdr[2] != 2) {
        gzclose(fp);
        throw std::runtime_error(path + ": stored estimator is not ERTL_MLE (hll.h:825-827 default)");
    }
    gz_read_exact(fp, dst, (size_t)1 << np, path);
    gzclose(fp);
}

// src/selection.cpp:12-33 (read_smh): u32 count, count x u64
void read_smh_into(const std::string& path, uint32_t expect_m, uint64_t* dst) {
    gzFile fp = gzopen(path.c_str(), "rb");
    if (fp == nullptr) throw std::runtime_error(std::string("Could not open file at '") + path + "' for reading");
    uint32_t n = 0;
    gz_read_exact(fp, &n, sizeof n, path);
    if (n != expect_m) { gzclose(fp); throw std::runtime_error(path + ": unexpected bucket count"); }
    if (n) gz_read_exact(fp, dst, (size_t)n * 8, path);
    gzclose(fp);
}

// src/selection.cpp:36-63
void load_file_list(std::vector<std::string>& files, const std::string& list_file) {
    if (list_file.empty()) { std::cerr << "No input file provided\n"; exit(-1); }
    std::ifstream file(list_file);
    if (!file.is_open()) { std::cerr << "No valid input file provided\n"; exit(-1); }
    std::string line;
    while (getline(file, line)) {
        line.erase(0, line.find_first_not_of(" \t\r\n"));
        line.erase(line.find_last_not_of(" \t\r\n") + 1);
        if (!line.empty()) files.push_back(line);
    }
}

[[noreturn]] void die(const char* what) {
    std::cerr << "selb200: " << what << ": " << selb200_last_error() << "\n";
    exit(2);
}

}  // namespace

int main(int argc, char* argv[]) {
    std::string list_file, criterion;
    unsigned threads = 8;
    unsigned aux_bytes = 256;
    float threshold = 0.9f;
    float z_score = 1.96f;
    int order_n = 1;
    int block_size = 256;
    bool verbose = false;
    (void)block_size;
#ifdef SELB_CUDA_DRIVER
    criterion = "smh_a";
#endif
    int c;
    while ((c = getopt(argc, argv, "xl:t:a:h:c:b:g")) != -1) {
        switch (c) {
            case 'x': std::cout << "Usage: -l -t -a -h -c\n"; return 0;
            case 'l': list_file = optarg; break;
            case 't': threads = (unsigned)std::stoi(optarg); break;
            case 'a': aux_bytes = (unsigned)std::stoi(optarg); break;
            case 'h': threshold = std::stof(optarg); break;
#ifndef SELB_CUDA_DRIVER
            case 'c': criterion = optarg; break;
#else
            case 'c': break;   // selection_cuda.cpp:68-88: in the optstring, no case
#endif
            case 'b': block_size = std::stoi(optarg); break;
            case 'g': verbose = true; break;
            default: break;
        }
    }
    omp_set_num_threads((int)threads);
    // The CUDA context comes up on its own thread (0.8 - 1.7 s on a B200 box) while this one reads the list and starts
    // decoding sketch files into pageable staging chunks; once the context is there the chunks move into the library's
    // pinned slots and the rest of the files is decoded straight into those.
    std::atomic<bool> warm_done{false};
    std::thread warm([&warm_done] { selb200_warmup(0); warm_done.store(true, std::memory_order_release); });
    struct Joiner { std::thread& t; ~Joiner() { if (t.joinable()) t.join(); } } joiner{warm};
    std::vector<std::string> files;
    load_file_list(files, list_file);

    int crit = -1, aux_kind = SELB200_AUX_NONE;
    if (criterion == "smh_a") { crit = SELB200_CRIT_SMH_A; aux_kind = SELB200_AUX_SMH; }
    else if (criterion == "hll_a") { crit = SELB200_CRIT_HLL_A; aux_kind = SELB200_AUX_HLL; }
    else if (criterion == "hll_an") { crit = SELB200_CRIT_HLL_AN; aux_kind = SELB200_AUX_HLL; }
    else if (criterion == "cb") { crit = SELB200_CRIT_CB; }
    if (crit < 0) {
        std::cout << "Option -c invalid. The accepted criteria are hll_a, hll_an and smh_a.\n";
        return 0;
    }
    const size_t n = files.size();
    const unsigned p_aux = aux_bytes ? (unsigned)__builtin_ctz(aux_bytes) : 0;   // selection.cpp:125
    const unsigned m_aux = aux_bytes / 8;                                      // selection.cpp:231
    const std::string aux_suffix = aux_kind == SELB200_AUX_SMH ? ".smh" + std::to_string(m_aux)
                                 : aux_kind == SELB200_AUX_HLL ? ".hll_" + std::to_string(p_aux) : "";

    const auto t_start = std::chrono::steady_clock::now();
    auto since = [](std::chrono::steady_clock::time_point t0) {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    };
    double ms_gunzip = 0.;

    // ---- load (OpenMP over the files of a chunk, like selection.cpp:241-249); the device copy,
    // validation, histogram and cardinality of chunk c run while chunk c+1 is being gunzipped ----
    int p = 14;
    if (n) p = (int)read_hll(files[0] + ".hll").np;
    const int aux_len = aux_kind == SELB200_AUX_SMH ? (int)m_aux : aux_kind == SELB200_AUX_HLL ? (int)p_aux : 0;
    const size_t m = (size_t)1 << p;
    const size_t aux_row = aux_kind == SELB200_AUX_SMH ? (size_t)m_aux * 8
                         : aux_kind == SELB200_AUX_HLL ? (size_t)1 << p_aux : 0;
    // rows [g0, g0 + count) of the file list -> register rows R, stored cardinalities S, auxiliary rows A
    auto decode = [&](int64_t g0, int64_t count, uint8_t* R, double* S, void* A) {
        std::string load_error;
        const auto t_gz = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic)
        for (int64_t i = 0; i < count; ++i) {
            try {
                const std::string& f = files[(size_t)(g0 + i)];
                read_hll_into(f + ".hll", (uint32_t)p, R + (size_t)i * m, S + i);
                if (aux_kind == SELB200_AUX_SMH) {
                    read_smh_into(f + aux_suffix, m_aux, reinterpret_cast<uint64_t*>((uint8_t*)A + (size_t)i * aux_row));
                } else if (aux_kind == SELB200_AUX_HLL) {
                    double unused;
                    read_hll_into(f + aux_suffix, p_aux, (uint8_t*)A + (size_t)i * aux_row, &unused);
                }
            } catch (const std::exception& e) {
#pragma omp critical
                if (load_error.empty()) load_error = e.what();
            }
        }
        ms_gunzip += since(t_gz);
        if (!load_error.empty()) throw std::runtime_error(load_error);   // uncaught, like the reference
    };
    // ---- while the context comes up: pageable staging chunks (half a pinned slot each, so that every one fits a slot) ----
    // (plain new[]: no value-initialisation, the pages are first touched by the decoding threads)
    struct Staged { int64_t g0 = 0, count = 0; std::unique_ptr<uint8_t[]> regs, aux; std::unique_ptr<double[]> stored; };
    std::vector<Staged> staged;
    const int64_t pre_rows = std::max<int64_t>(1, (int64_t)(32u << 20) / (int64_t)m);
    int64_t g_next = 0;
    while (g_next < (int64_t)n && !warm_done.load(std::memory_order_acquire)) {
        Staged st;
        st.g0 = g_next;
        st.count = std::min<int64_t>(pre_rows, (int64_t)n - g_next);
        st.regs.reset(new uint8_t[(size_t)st.count * m]);
        st.stored.reset(new double[(size_t)st.count]);
        st.aux.reset(new uint8_t[std::max<size_t>((size_t)st.count * aux_row, 1)]);
        decode(st.g0, st.count, st.regs.get(), st.stored.get(), st.aux.get());
        g_next += st.count;
        staged.push_back(std::move(st));
    }
    const double ms_staged = since(t_start);
    const int64_t rows_staged = g_next;

    // ---- device context; the sketches not yet decoded go straight into its pinned staging slots ----
    selb200_ctx* ctx = nullptr;
    if (selb200_create(0, nullptr, &ctx) != SELB200_OK) die("create");
    const double ms_create = since(t_start);
    const auto t_load = std::chrono::steady_clock::now();

    // ---- load (OpenMP over the files of a chunk, like selection.cpp:241-249); the device copy,
    // validation, histogram and cardinality of chunk c run while chunk c+1 is being gunzipped ----
    int64_t rows_per_chunk = 0;
    if (selb200_load_begin(ctx, (int64_t)n, p, aux_kind, aux_len, &rows_per_chunk) != SELB200_OK) die("load");
    // staged chunks first (two to a slot where they fit), then the rest of the list
    for (size_t si = 0; si < staged.size();) {
        int64_t count = 0;
        size_t sj = si;
        while (sj < staged.size() && count + staged[sj].count <= rows_per_chunk) count += staged[sj++].count;
        if (sj == si) { count = staged[si].count; sj = si + 1; }          // cannot happen: pre_rows <= rows_per_chunk
        uint8_t* R = nullptr;
        double* S = nullptr;
        void* A = nullptr;
        if (selb200_load_acquire(ctx, staged[si].g0, count, &R, &S, &A) != SELB200_OK) die("load");
        int64_t off = 0;
        for (size_t k = si; k < sj; ++k) {
            const Staged& st = staged[k];
            const int64_t rows = st.count;
#pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < rows; ++i) {
                std::memcpy(R + (size_t)(off + i) * m, st.regs.get() + (size_t)i * m, m);
                if (aux_row) std::memcpy((uint8_t*)A + (size_t)(off + i) * aux_row, st.aux.get() + (size_t)i * aux_row, aux_row);
                S[off + i] = st.stored[(size_t)i];
            }
            off += rows;
            staged[k] = Staged{};                                          // free the chunk
        }
        if (selb200_load_commit(ctx) != SELB200_OK) die("load");
        si = sj;
    }
    for (int64_t g0 = g_next; g0 < (int64_t)n; g0 += rows_per_chunk) {
        const int64_t count = std::min<int64_t>(rows_per_chunk, (int64_t)n - g0);
        uint8_t* R = nullptr;
        double* S = nullptr;
        void* A = nullptr;
        if (selb200_load_acquire(ctx, g0, count, &R, &S, &A) != SELB200_OK) die("load");
        decode(g0, count, R, S, A);
        if (selb200_load_commit(ctx) != SELB200_OK) die("load");
    }
    if (selb200_load_end(ctx) != SELB200_OK) die("load");
    const double ms_load = since(t_load);
    const auto t_run = std::chrono::steady_clock::now();
    selb200_params prm;
    selb200_default_params(&prm);
    prm.tau = threshold;
    prm.criterion = crit;
    prm.z_score = z_score;
    prm.order_n = order_n;
    if (crit == SELB200_CRIT_SMH_A) {
#ifdef SELB_CUDA_DRIVER
        selb200_band_params((int)m_aux, threshold, 0, &prm.n_bands, &prm.n_rows);
#else
        selb200_band_params((int)m_aux, threshold, 1, &prm.n_bands, &prm.n_rows);
#endif
    }
    selb200_stats st;
    if (selb200_run(ctx, &prm, &st) != SELB200_OK) die("run");
    const int64_t cnt = selb200_result_count(ctx);
    std::vector<int32_t> ri((size_t)cnt), rk((size_t)cnt), order(n);
    std::vector<double> rj((size_t)cnt);
    if (selb200_copy_results(ctx, cnt, ri.data(), rk.data(), rj.data()) != SELB200_OK) die("results");
    if (selb200_get_order(ctx, nullptr, order.data()) != SELB200_OK) die("order");
    const double ms_run = since(t_run);
    const auto t_print = std::chrono::steady_clock::now();

    std::string out;
    out.reserve((size_t)cnt * 96);
    for (int64_t t = 0; t < cnt; ++t) {
        const std::string& a = files[(size_t)order[(size_t)ri[(size_t)t]]];
        const std::string& b = files[(size_t)order[(size_t)rk[(size_t)t]]];
#ifdef SELB_CUDA_DRIVER
        char buf[64];
        snprintf(buf, sizeof buf, "%g", (double)(float)rj[(size_t)t]);   // operator<<(float), 6 significant digits
        out += a + " " + b + " " + buf + "\n";
#else
        out += a + " " + b + " " + std::to_string(rj[(size_t)t]) + "\n";
#endif
    }
    std::cout << out;
    std::cout.flush();
    if (verbose)
        fprintf(stderr, "selb200: host ms: context ready at %.0f (%lld of %lld rows decoded into staging by %.0f), load after that %.0f, "
                        "gunzip in all %.0f (%u threads), run+fetch %.1f, print %.1f, total %.0f\n", ms_create, (long long)rows_staged,
                (long long)n, ms_staged, ms_load, ms_gunzip, threads, ms_run, since(t_print), since(t_start));
    if (verbose)
        fprintf(stderr, "selb200: n=%lld pairs=%lld P_cb=%lld P_aux=%lld P_out=%lld near=%lld | bands x rows %dx%d | "
                        "device ms: bounds %.3f filter %.3f verify %.3f union %.3f estimate %.3f sort %.3f total %.3f\n",
                (long long)st.n, (long long)st.pairs_total, (long long)st.pairs_cb, (long long)st.pairs_aux,
                (long long)st.pairs_out, (long long)st.pairs_near, st.n_bands, st.n_rows, st.ms_bounds, st.ms_filter,
                st.ms_verify, st.ms_union, st.ms_estimate, st.ms_sort, st.ms_total);
    selb200_destroy(ctx);
    return 0;
}
