// emul_sketch.cpp — runs the sketch builder's kernel (csrc/kernels/sketch_kernels.inl: rolling canonical 31-mers,
// WangHash, HLL registers, SuperMinHash buckets; one 512-thread CTA per genome) on the CPU through cuda_emul.h.
// Input (file): cleaned sequences (records joined by one 'N') with offsets, p, auxiliary kind and length.
// Output (file): primary HLL registers and the auxiliary sketch of every genome — tests/test_emul_sketch.py holds
// them against the sketch files the REFERENCE's build_sketch wrote (tests/golden/influenza).  Test infrastructure.
#include <cstdio>
#include <cstdlib>

#define SELB_EMUL 1
#include "cuda_emul.h"
#include "../../include/selb200.h"

#include "../../cuda_selection_criteria_b200/csrc/kernels/sketch_kernels.inl"

template <class T> static void rd(FILE* f, T* p, size_t n) { if (fread(p, sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } }
template <class T> static void wr(FILE* f, const T* p, size_t n) { if (fwrite(p, sizeof(T), n, f) != n) { fprintf(stderr, "short write\n"); exit(2); } }

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: emul_sketch in.bin out.bin\n"); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    int32_t hdr[4];
    rd(f, hdr, 4);
    const int n = hdr[0], p = hdr[1], aux_kind = hdr[2], aux_len = hdr[3];     // aux_len: buckets (smh) or precision (hll)
    std::vector<long long> offsets((size_t)n + 1);
    rd(f, offsets.data(), offsets.size());
    std::vector<uint8_t> seq((size_t)offsets[n] + 1);
    rd(f, seq.data(), (size_t)offsets[n]);
    fclose(f);
    const size_t m_hll = (size_t)1 << p;
    const size_t m_aux = aux_kind == SELB200_AUX_HLL ? (size_t)1 << aux_len : 0;
    const size_t m_smh = aux_kind == SELB200_AUX_SMH ? (size_t)aux_len : 0;
    if ((m_hll + m_aux) * 4 + m_smh * 8 > sizeof smem_raw) { fprintf(stderr, "sketch too large for the emulated shared memory\n"); return 3; }
    std::vector<uint8_t> hll((size_t)n * m_hll), auxh((size_t)n * m_aux + 1);
    std::vector<unsigned long long> smh((size_t)n * m_smh + 1);
    if (m_smh > 256) {
        std::vector<uint16_t> perm((size_t)n * SK_THREADS * 2 * m_smh);
        emul::launch((unsigned)n, SK_THREADS, [&] {
            k_sketch_build<uint16_t>(seq.data(), offsets.data(), p, aux_kind, aux_len, hll.data(), auxh.data(), smh.data(), perm.data());
        });
    } else {
        std::vector<uint8_t> perm((size_t)n * SK_THREADS * 2 * m_smh + 1);
        emul::launch((unsigned)n, SK_THREADS, [&] {
            k_sketch_build<uint8_t>(seq.data(), offsets.data(), p, aux_kind, aux_len, hll.data(), auxh.data(), smh.data(), perm.data());
        });
    }
    f = fopen(argv[2], "wb");
    if (!f) { perror(argv[2]); return 2; }
    wr(f, hll.data(), (size_t)n * m_hll);
    wr(f, auxh.data(), (size_t)n * m_aux);
    wr(f, smh.data(), (size_t)n * m_smh);
    fclose(f);
    printf("n=%d p=%d aux_kind=%d aux_len=%d bases=%lld\n", n, p, aux_kind, aux_len, offsets[n]);
    return 0;
}
