// build_sketch_main.cpp — C++ host driver of the CUDA sketch builder (include/selb200.h).
//
// Drop-in for the reference's `build_sketch -l list -t threads -a aux_bytes -c {smh_a|hll_a|hll_an}`
// (src/build_sketch.cpp:186-295): for every path P of the list it writes `P.hll` (p = 14) and
// `P.smh<a/8>` or `P.hll_<ctz(a)>` in the reference's gzip formats (sketch/include/sketch/hll.h:1103-1124,
// src/build_sketch.cpp:9-20); an invalid -c still writes the primary sketches and then prints the
// reference's message (:290-292).  FASTA records are read with SeqAn's rules
// (seqan/seq_io/fasta_fastq.h:262-282): skip to '>', id = rest of the line, sequence = everything up to
// the next '>' with whitespace dropped; a character outside the IUPAC alphabet is a ParseError, at
// which the reference stops reading the file (src/build_sketch.cpp:55-58).  The host only inflates
// and de-lines the FASTA (OpenMP over files); k-mers, hashing and both sketches are computed on
// the GPU by selb200_sketch_host.
#include <getopt.h>
#include <omp.h>
#include <zlib.h>

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

void load_file_list(std::vector<std::string>& files, const std::string& list_file) {   // build_sketch.cpp:152-180
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

// whole (optionally gzip) file -> record sequences joined by 'N'
bool read_fasta_clean(const std::string& path, std::string& out) {
    out.clear();
    gzFile fp = gzopen(path.c_str(), "rb");
    if (fp == nullptr) {
        std::cerr << "ERROR: Could not open the file " << path << ".\n";      // build_sketch.cpp:44-48
        return false;
    }
    gzbuffer(fp, 1 << 20);
    std::string raw;
    std::vector<char> buf(1 << 20);
    for (;;) {
        const int got = gzread(fp, buf.data(), (unsigned)buf.size());
        if (got <= 0) break;
        raw.append(buf.data(), (size_t)got);
    }
    gzclose(fp);
    size_t i = 0;
    const size_t n = raw.size();
    bool first = true;
    while (i < n) {
        while (i < n && raw[i] != '>') ++i;          // skipUntil('>')
        if (i >= n) break;
        while (i < n && raw[i] != '\n') ++i;         // id line
        const size_t mark = out.size();
        if (!first) out.push_back('N');
        bool bad = false;
        while (i < n && raw[i] != '>') {
            const unsigned char c = (unsigned char)raw[i++];
            if (is_space(c)) continue;
            if (!is_iupac(c)) { bad = true; break; }
            out.push_back((char)c);
        }
        if (bad) { out.resize(mark); break; }        // ParseError: this record and the rest are dropped
        first = false;
    }
    return true;
}

void gz_write_all(gzFile fp, const void* src, size_t len) {
    if (len && gzwrite(fp, src, (unsigned)len) == 0) throw std::runtime_error("Error writing to file.");
}

void write_hll(const std::string& path, const uint8_t* regs, uint32_t np) {            // hll.h:1103-1124
    gzFile fp = gzopen(path.c_str(), "wb");
    if (!fp) throw std::runtime_error(std::string("Could not open file at '") + path + "' for writing");
    const uint32_t hdr[4] = {0u, 2u, 2u, 1u};      // is_calculated, estim = ERTL_MLE, jestim = ERTL_MLE, 1
    const double value = -1.0;
    gz_write_all(fp, hdr, sizeof hdr);
    gz_write_all(fp, &np, sizeof np);
    gz_write_all(fp, &value, sizeof value);
    gz_write_all(fp, regs, (size_t)1 << np);
    gzclose(fp);
}

void write_smh(const std::string& path, const uint64_t* h, uint32_t m) {              // build_sketch.cpp:9-20
    gzFile fp = gzopen(path.c_str(), "wb");
    if (!fp) throw std::runtime_error(std::string("Could not open file at '") + path + "' for writing");
    gz_write_all(fp, &m, sizeof m);
    gz_write_all(fp, h, (size_t)m * 8);
    gzclose(fp);
}

}  // namespace

int main(int argc, char* argv[]) {
    std::string list_file, criterion;
    unsigned threads = 8, aux_bytes = 256;
    bool verbose = false;
    int c;
    while ((c = getopt(argc, argv, "l:t:a:c:g")) != -1) {
        switch (c) {
            case 'l': list_file = optarg; break;
            case 't': threads = (unsigned)std::stoi(optarg); break;
            case 'a': aux_bytes = (unsigned)std::stoi(optarg); break;
            case 'c': criterion = optarg; break;
            case 'g': verbose = true; break;
            default: break;
        }
    }
    omp_set_num_threads((int)threads);
    std::thread warm([] { selb200_warmup(0); });   // CUDA context comes up while the files are read
    struct Joiner { std::thread& t; ~Joiner() { if (t.joinable()) t.join(); } } joiner{warm};
    std::vector<std::string> files;
    load_file_list(files, list_file);

    int aux_kind = SELB200_AUX_NONE, aux_len = 0;
    std::string aux_suffix;
    if (criterion == "smh_a") {
        aux_kind = SELB200_AUX_SMH;
        aux_len = (int)(aux_bytes / 8);                                   // build_sketch.cpp:274
        aux_suffix = ".smh" + std::to_string(aux_bytes / 8);               // :288
    } else if (criterion == "hll_a" || criterion == "hll_an") {
        aux_kind = SELB200_AUX_HLL;
        aux_len = aux_bytes ? __builtin_ctz(aux_bytes) : 0;                // :243,259
        aux_suffix = ".hll_" + std::to_string(aux_len);
    }
    const int p = 14;
    const size_t m_hll = (size_t)1 << p;
    const size_t aux_elems = aux_kind == SELB200_AUX_SMH ? (size_t)selb200_smh_size(aux_len)
                           : aux_kind == SELB200_AUX_HLL ? (size_t)1 << aux_len : 0;
    const size_t aux_row = aux_kind == SELB200_AUX_SMH ? aux_elems * 8 : aux_elems;
    const size_t batch_bytes = (size_t)1 << 30;    // sequence characters per device batch
    size_t done = 0, total_bases = 0;
    while (done < files.size()) {
        // ---- inflate + de-line a batch of genomes (OpenMP over files) -------------------------------
        std::vector<std::string> seqs;
        size_t take = 0, bytes = 0;
        const size_t group = std::max<size_t>(threads * 4, 16);
        while (done + take < files.size() && bytes < batch_bytes) {
            const size_t g1 = std::min(files.size(), done + take + group);
            const size_t base = seqs.size();
            seqs.resize(base + (g1 - (done + take)));
#pragma omp parallel for schedule(dynamic)
            for (size_t i = done + take; i < g1; ++i) read_fasta_clean(files[i], seqs[base + (i - (done + take))]);
            for (size_t i = base; i < seqs.size(); ++i) bytes += seqs[i].size();
            take = g1 - done;
        }
        std::vector<int64_t> offsets(take + 1, 0);
        for (size_t i = 0; i < take; ++i) offsets[i + 1] = offsets[i] + (int64_t)seqs[i].size();
        std::vector<uint8_t> blob((size_t)offsets[take] + 16);
        for (size_t i = 0; i < take; ++i) std::memcpy(blob.data() + offsets[i], seqs[i].data(), seqs[i].size());
        seqs.clear();
        std::vector<uint8_t> hll(take * m_hll), aux(take * aux_row + 8);
        if (selb200_sketch_host(0, (int64_t)take, blob.data(), offsets.data(), p, aux_kind, aux_len, hll.data(),
                                aux_kind == SELB200_AUX_NONE ? nullptr : aux.data()) != SELB200_OK) {
            std::cerr << "selb200: sketch: " << selb200_sketch_last_error() << "\n";
            return 2;
        }
        total_bases += (size_t)offsets[take];
        // ---- write the files (OpenMP over files) -----------------------------------------------------
#pragma omp parallel for schedule(dynamic)
        for (size_t i = 0; i < take; ++i) {
            const std::string& f = files[done + i];
            write_hll(f + ".hll", hll.data() + i * m_hll, (uint32_t)p);                  // build_sketch.cpp:237
            if (aux_kind == SELB200_AUX_SMH)
                write_smh(f + aux_suffix, reinterpret_cast<const uint64_t*>(aux.data() + i * aux_row), (uint32_t)aux_elems);
            else if (aux_kind == SELB200_AUX_HLL)
                write_hll(f + aux_suffix, aux.data() + i * aux_row, (uint32_t)aux_len);
        }
        done += take;
    }
    if (aux_kind == SELB200_AUX_NONE)
        printf("Option -c invalid. The accepted criteria are hll_a, hll_an and smh_a.\n");    // :290-292
    if (verbose) fprintf(stderr, "selb200: sketched %zu genomes, %zu sequence characters\n", files.size(), total_bases);
    return 0;
}
