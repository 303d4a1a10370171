// time_smh_cuda_main.cpp — drop-in for experiments/src/time_smh_cuda.cpp (flags -l -h -m -b, -x usage).
//
// Same three stdout lines, `list;build_smh;tau;secs`, `list;smh_a;tau;secs`, `list;CB+smh_a;tau;secs`
// (experiments/src/time_smh_cuda.cpp:228-230,279-299), with two differences that are the point of the
// replacement: the GPU work is timed to completion (the reference brackets an asynchronous launch with
// no synchronisation, so it prints launch overhead), and the compare phases run the tiled path of
// selb200_run instead of a host-materialised pair list.
//   build_smh : load every `P.hll`, re-sketch the SuperMinHash of every FASTA `P` with
//               SuperMinHash<>(M-1) semantics (time_smh_cuda.cpp:36-38,196-207) on the GPU, load + sort
//   smh_a     : selection without the cardinality bound (experiments/src/time_smh.cpp:229-257)
//   CB+smh_a  : selection with it (time_smh.cpp:261-292)
// Band shape: the drivers' search that keeps (1,1) when nothing qualifies (time_smh_cuda.cpp:231-240).
//
// Built twice (csrc/Makefile): bin/time_smh_cuda, and with -DSELB_TIME_SMH bin/time_smh — the flag set and line
// format of the reference's CPU harness (experiments/src/time_smh.cpp:139 "xl:t:h:m:R:", ";m:M" / ";r:R_b:B"
// suffixes, -R repetitions of the two compare phases, :226-292) over the same GPU path, so run_time_experiment.sh's
// CPU leg has a binary too.  The reference prints each time AFTER the next label (TIMERSTOP has no newline,
// include/metrictime2.hpp:14-18); here every line carries its own seconds in field 4, which is what the script's
// awk reads.  Both builds also report, on stderr, the CUDA-event time of each compare phase
// (`selb200: device ms: ...`): the printed wall-clock seconds bracket a synchronised run and must not be below it.
#include <getopt.h>
#include <omp.h>
#include <zlib.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/selb200.h"

namespace {

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

bool is_iupac(unsigned char c) {
    switch (c | 0x20) {
        case 'a': case 'c': case 'g': case 't': case 'u': case 'r': case 'y': case 's': case 'w':
        case 'k': case 'm': case 'b': case 'd': case 'h': case 'v': case 'n': return true;
    }
    return false;
}
bool is_space(unsigned char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; }

bool slurp_gz(const std::string& path, std::string& raw) {
    gzFile fp = gzopen(path.c_str(), "rb");
    if (fp == nullptr) return false;
    gzbuffer(fp, 1 << 20);
    std::vector<char> buf(1 << 20);
    for (;;) {
        const int got = gzread(fp, buf.data(), (unsigned)buf.size());
        if (got <= 0) break;
        raw.append(buf.data(), (size_t)got);
    }
    gzclose(fp);
    return true;
}

// record sequences joined by 'N' (SeqAn rules, see build_sketch_main.cpp); tries P then P.gz
// like time_smh_cuda.cpp:40-49
void read_fasta_clean(const std::string& path, std::string& out) {
    out.clear();
    std::string raw;
    if (!slurp_gz(path, raw) && !slurp_gz(path + ".gz", raw)) {
        std::cerr << "ERROR: Could not open the file " << path << " or " << path << ".gz.\n";
        return;
    }
    size_t i = 0;
    const size_t n = raw.size();
    bool first = true;
    while (i < n) {
        while (i < n && raw[i] != '>') ++i;
        if (i >= n) break;
        while (i < n && raw[i] != '\n') ++i;
        const size_t mark = out.size();
        if (!first) out.push_back('N');
        bool bad = false;
        while (i < n && raw[i] != '>') {
            const unsigned char c = (unsigned char)raw[i++];
            if (is_space(c)) continue;
            if (!is_iupac(c)) { bad = true; break; }
            out.push_back((char)c);
        }
        if (bad) { out.resize(mark); break; }
        first = false;
    }
}

void read_hll_into(const std::string& path, uint32_t expect_np, uint8_t* dst, double* value) {
    gzFile fp = gzopen(path.c_str(), "rb");
    if (fp == nullptr) throw std::runtime_error(std::string("Could not open file at '") + path + "' for reading");
    uint32_t hdr[5];
    bool ok = gzread(fp, hdr, sizeof hdr) == (int)sizeof hdr && gzread(fp, value, 8) == 8;
    ok = ok && hdr[4] == expect_np && gzread(fp, dst, 1u << expect_np) == (int)(1u << expect_np);
    gzclose(fp);
    if (!ok) throw std::runtime_error("Error reading from file " + path);
}

double seconds_since(std::chrono::high_resolution_clock::time_point t0) {
    return std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
}

[[noreturn]] void die(const char* what, const char* msg) {
    std::cerr << "selb200: " << what << ": " << msg << "\n";
    exit(2);
}

}  // namespace

int main(int argc, char* argv[]) {
    std::string list_file;
    unsigned threads = 8;                 // time_smh_cuda.cpp:145 (fixed there; -t in time_smh.cpp:149)
    float threshold = 0.9f;
    int mh_size = 8, block_size = 256, total_rep = 1;
    (void)block_size;
    int c;
#ifdef SELB_TIME_SMH
    const char* optstring = "xl:t:h:m:R:";      // experiments/src/time_smh.cpp:139
#else
    const char* optstring = "xl:h:m:b:";        // experiments/src/time_smh_cuda.cpp:154
#endif
    while ((c = getopt(argc, argv, optstring)) != -1) {
        switch (c) {
            case 'x': std::cout << "Usage: -l -t -h -m\n"; return 0;
            case 'l': list_file = optarg; break;
            case 'h': threshold = std::stof(optarg); break;
            case 'm': mh_size = std::stoi(optarg); break;
            case 'b': block_size = std::stoi(optarg); break;
            case 't': threads = (unsigned)std::stoi(optarg); break;
            case 'R': total_rep = std::stoi(optarg); break;
            default: break;
        }
    }
    omp_set_num_threads((int)threads);
    std::thread warm([] { selb200_warmup(0); });
    struct Joiner { std::thread& t; ~Joiner() { if (t.joinable()) t.join(); } } joiner{warm};
    std::vector<std::string> files;
    load_file_list(files, list_file);
    const int64_t n = (int64_t)files.size();
    const int p = 14;
    const size_t m_hll = (size_t)1 << p;
    if (mh_size < 1) die("flags", "-m must be positive");
    const int m_build = selb200_smh_size(mh_size - 1 > 0 ? mh_size - 1 : 1);    // SuperMinHash<>(M-1)
    if (m_build < mh_size) die("flags", "-m: SuperMinHash<>(M-1) holds fewer than M buckets (the reference reads past its end)");

    // ---- build_smh ----------------------------------------------------------------------------
    auto t0 = std::chrono::high_resolution_clock::now();
    std::vector<uint8_t> regs((size_t)n * m_hll);
    std::vector<double> stored((size_t)n, -1.0);
    std::vector<std::string> seqs((size_t)n);
    std::string load_error;
#pragma omp parallel for schedule(dynamic)
    for (int64_t i = 0; i < n; ++i) {
        try {
            read_hll_into(files[(size_t)i] + ".hll", (uint32_t)p, regs.data() + (size_t)i * m_hll, &stored[(size_t)i]);
            read_fasta_clean(files[(size_t)i], seqs[(size_t)i]);
        } catch (const std::exception& e) {
#pragma omp critical
            if (load_error.empty()) load_error = e.what();
        }
    }
    if (!load_error.empty()) throw std::runtime_error(load_error);
    std::vector<int64_t> offsets((size_t)n + 1, 0);
    for (int64_t i = 0; i < n; ++i) offsets[(size_t)i + 1] = offsets[(size_t)i] + (int64_t)seqs[(size_t)i].size();
    std::vector<uint8_t> blob((size_t)offsets[(size_t)n] + 16);
    for (int64_t i = 0; i < n; ++i) std::memcpy(blob.data() + offsets[(size_t)i], seqs[(size_t)i].data(), seqs[(size_t)i].size());
    seqs.clear();
    std::vector<uint8_t> scratch_hll((size_t)n * m_hll);
    std::vector<uint64_t> smh_full((size_t)n * (size_t)m_build), smh((size_t)n * (size_t)mh_size);
    if (joiner.t.joinable()) joiner.t.join();
    if (selb200_sketch_host(0, n, blob.data(), offsets.data(), p, SELB200_AUX_SMH, m_build, scratch_hll.data(),
                            smh_full.data()) != SELB200_OK)
        die("sketch", selb200_sketch_last_error());
    for (int64_t i = 0; i < n; ++i)       // mh_vector[i] = smh.h_[i] for i < M (time_smh_cuda.cpp:95-97)
        std::memcpy(smh.data() + (size_t)i * mh_size, smh_full.data() + (size_t)i * m_build, (size_t)mh_size * 8);
    selb200_ctx* ctx = nullptr;
    if (selb200_create(0, nullptr, &ctx) != SELB200_OK) die("create", selb200_last_error());
    if (selb200_load_host(ctx, n, p, regs.data(), stored.data(), SELB200_AUX_SMH, mh_size, smh.data()) != SELB200_OK)
        die("load", selb200_last_error());
#ifdef SELB_TIME_SMH
    std::cout << list_file << ";build_smh;" << threshold << ";" << seconds_since(t0) << ";m:" << mh_size << std::endl;
#else
    std::cout << list_file << ";build_smh;" << threshold << ";" << seconds_since(t0) << std::endl;
#endif

    selb200_params prm;
    selb200_default_params(&prm);
    prm.tau = threshold;
    prm.criterion = SELB200_CRIT_SMH_A;
    selb200_band_params(mh_size, threshold, 0, &prm.n_bands, &prm.n_rows);
    selb200_stats st;
    long long out_smh = 0;
    for (int rep = 0; rep < total_rep; ++rep) {
        // ---- smh_a (no cardinality bound) and CB+smh_a; selb200_run returns after the device finished ----
        const char* label[2] = {"smh_a", "CB+smh_a"};
        float dev_ms[2] = {0.f, 0.f};
        double secs[2] = {0., 0.};
        for (int phase = 0; phase < 2; ++phase) {
            prm.no_cb = phase == 0;
            t0 = std::chrono::high_resolution_clock::now();
            if (selb200_run(ctx, &prm, &st) != SELB200_OK) die("run", selb200_last_error());
            secs[phase] = seconds_since(t0);
            dev_ms[phase] = st.ms_total;
            std::cout << list_file << ";" << label[phase] << ";" << threshold << ";" << secs[phase];
#ifdef SELB_TIME_SMH
            std::cout << ";r:" << prm.n_rows << "_b:" << prm.n_bands;
#endif
            std::cout << std::endl;
            if (phase == 0) out_smh = (long long)st.pairs_out;
        }
        fprintf(stderr, "selb200: device ms: smh_a %.4f, CB+smh_a %.4f | wall ms: %.4f, %.4f\n", dev_ms[0], dev_ms[1],
                secs[0] * 1e3, secs[1] * 1e3);
    }
    fprintf(stderr, "selb200: n=%lld bands x rows %dx%d | pairs out: smh_a %lld, CB+smh_a %lld\n", (long long)n,
            prm.n_bands, prm.n_rows, out_smh, (long long)st.pairs_out);
    selb200_destroy(ctx);
    return 0;
}
