// hostpack.h — host side of the packed upload of HLL register matrices (selb200_load_host).
//
// An HLL register is a byte holding a value <= 64-p+1, and the values of one sketch sit in a narrow band above the
// smallest one (P(value >= k) falls off as 2^-k), so a genome travels over PCIe as
//   base            u8           its smallest register value
//   nibbles         m/2 bytes    register j -> min(value - base, 15) in nibble j&1 of byte j>>1
//   exceptions      <= 32 x u32  (position << 8 | value) of the registers with value - base >= 15, ascending position
// i.e. 4.06 bits per register instead of 8.  A genome with more than 32 such registers (not an HLL of a real set) is
// flagged raw and its bytes are copied as they are.  The device side (k_unpack_nib4, kernels/load_kernels.inl) rebuilds
// the byte matrix, which is what every later kernel of the load reads: results cannot depend on the transport.
#pragma once
#include <cstddef>
#include <cstdint>

namespace selb {
constexpr int NIB4_EXC_CAP = 32;
struct Nib4Hdr {            // one per genome
    uint8_t base;
    uint8_t raw;            // 1: more than NIB4_EXC_CAP exceptions — nibbles unusable, the genome follows as raw bytes
    uint16_t n_exc;
};
// packs genomes [0, rows) of `regs` (rows x m bytes, m a multiple of 64); nib: rows x m/2 bytes, exc: rows x NIB4_EXC_CAP,
// hdr: rows.  Runs on `threads` OpenMP threads (<= 0: the OpenMP default).  Returns the number of raw genomes.
int64_t nib4_pack(const uint8_t* regs, int64_t rows, size_t m, uint8_t* nib, uint32_t* exc, Nib4Hdr* hdr, int threads);
// "avx2" or "scalar": the code path nib4_pack takes on this machine
const char* nib4_impl();

// A PIECE: rows genomes as one contiguous byte buffer, so that one copy (and one all-gather) moves them:
//   [hdr rows x 4 | exc rows x 128 | raw_idx NIB4_RAW_CAP x i32 | nib rows x m/2 | raw NIB4_RAW_CAP x m]
// every part starting at a multiple of 256 bytes.  raw_idx[r] = row (inside the piece) whose bytes sit in raw slot r, -1 = free.
constexpr int NIB4_RAW_CAP = 4;
struct Nib4Piece {
    size_t off_hdr, off_exc, off_rawidx, off_nib, off_raw, bytes;
};
#if defined(__CUDACC__)
#define SELB_PK_HD __host__ __device__ inline
#else
#define SELB_PK_HD inline
#endif
SELB_PK_HD size_t nib4_up256(size_t x) { return (x + 255) & ~(size_t)255; }
SELB_PK_HD Nib4Piece nib4_piece(long long rows, size_t m) {
    Nib4Piece L;
    L.off_hdr = 0;
    L.off_exc = nib4_up256((size_t)rows * sizeof(Nib4Hdr));
    L.off_rawidx = L.off_exc + nib4_up256((size_t)rows * NIB4_EXC_CAP * 4);
    L.off_nib = L.off_rawidx + 256;
    L.off_raw = L.off_nib + nib4_up256((size_t)rows * (m >> 1));
    L.bytes = L.off_raw + (size_t)NIB4_RAW_CAP * m;
    return L;
}
// packs a piece; returns the number of raw genomes (those beyond NIB4_RAW_CAP have no slot: the caller must move the
// piece some other way) — the bytes worth copying are [0, off_raw + min(n_raw, NIB4_RAW_CAP) * m)
int64_t nib4_pack_piece(const uint8_t* regs, int64_t rows, size_t m, uint8_t* piece, int threads);
}  // namespace selb
