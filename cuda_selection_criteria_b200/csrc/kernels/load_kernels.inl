// load_kernels.inl — load-time kernels: validation, cardinalities, sorted re-layout (part of selb200.cu)
// ============================================================================
// K0: register validation — max byte over a buffer (values must be <= 64-p+1)
// ============================================================================
__global__ void __launch_bounds__(256) k_max_byte(const uint4* __restrict__ data, size_t n16, uint32_t* out) {
    uint32_t mx = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = __ldg(data + i);
        mx = __vmaxu4(mx, __vmaxu4(__vmaxu4(v.x, v.y), __vmaxu4(v.z, v.w)));
    }
    uint32_t b = max(max(mx & 0xff, (mx >> 8) & 0xff), max((mx >> 16) & 0xff, mx >> 24));
    for (int o = 16; o; o >>= 1) b = max(b, __shfl_xor_sync(0xffffffffu, b, o));
    if ((threadIdx.x & 31) == 0 && b) atomicMax(out, b);
}
// ============================================================================
// packed upload (hostpack.h): nibbles + per-genome base -> register bytes, then the exceptions, then raw rows.
// `pieces` holds consecutive packed pieces of piece_rows rows each (the last one shorter), piece_stride bytes apart;
// blockIdx.y = piece.  A thread turns 8 bytes of nibbles into 16 register bytes (one 128-bit store); raw rows are skipped
// by the first two kernels and copied by the third.
// ============================================================================
struct Nib4Pieces {
    const uint8_t* pieces;
    size_t piece_stride;
    long long piece_rows, count;      // rows per piece, rows in all
    int log2_m;
    __device__ __forceinline__ long long rows_of(int pi) const { return min(piece_rows, count - (long long)pi * piece_rows); }
};

__global__ void __launch_bounds__(256) k_unpack_nib4(Nib4Pieces a, uint8_t* __restrict__ regs) {
    const int pi = blockIdx.y;
    const long long rows = a.rows_of(pi);
    const selb::Nib4Piece P = selb::nib4_piece(rows, (size_t)1 << a.log2_m);
    const uint8_t* piece = a.pieces + (size_t)pi * a.piece_stride;
    const selb::Nib4Hdr* hdr = reinterpret_cast<const selb::Nib4Hdr*>(piece + P.off_hdr);
    const uint2* nib = reinterpret_cast<const uint2*>(piece + P.off_nib);
    uint4* out = reinterpret_cast<uint4*>(regs + (((size_t)pi * (size_t)a.piece_rows) << a.log2_m));
    const long long total = rows << (a.log2_m - 4);               // 16-register groups
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const selb::Nib4Hdr h = hdr[idx >> (a.log2_m - 4)];
        if (h.raw) continue;
        const uint2 v = __ldg(nib + idx);
        const uint32_t b4 = 0x01010101u * h.base;                  // value - base <= 15 and values < 128: no carry between bytes
        // bytes n0 | n1<<4 -> n0, n1: the even registers of a word are its low nibbles, the odd ones its high nibbles
        const uint32_t lo0 = v.x & 0x0f0f0f0fu, hi0 = (v.x >> 4) & 0x0f0f0f0fu;
        const uint32_t lo1 = v.y & 0x0f0f0f0fu, hi1 = (v.y >> 4) & 0x0f0f0f0fu;
        uint4 o;
        o.x = __byte_perm(lo0, hi0, 0x5140) + b4;                  // registers 0..3 (nibble bytes 0, 1)
        o.y = __byte_perm(lo0, hi0, 0x7362) + b4;                  // registers 4..7 (nibble bytes 2, 3)
        o.z = __byte_perm(lo1, hi1, 0x5140) + b4;
        o.w = __byte_perm(lo1, hi1, 0x7362) + b4;
        out[idx] = o;
    }
}
__global__ void __launch_bounds__(256) k_apply_nib4_exc(Nib4Pieces a, uint8_t* __restrict__ regs) {
    const int pi = blockIdx.y;
    const long long rows = a.rows_of(pi);
    const selb::Nib4Piece P = selb::nib4_piece(rows, (size_t)1 << a.log2_m);
    const uint8_t* piece = a.pieces + (size_t)pi * a.piece_stride;
    const selb::Nib4Hdr* hdr = reinterpret_cast<const selb::Nib4Hdr*>(piece + P.off_hdr);
    const uint32_t* exc = reinterpret_cast<const uint32_t*>(piece + P.off_exc);
    uint8_t* out = regs + (((size_t)pi * (size_t)a.piece_rows) << a.log2_m);
    const long long total = rows * selb::NIB4_EXC_CAP;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long g = idx / selb::NIB4_EXC_CAP;
        const int e = (int)(idx - g * selb::NIB4_EXC_CAP);
        const selb::Nib4Hdr h = hdr[g];
        if (h.raw || e >= (int)h.n_exc) continue;
        const uint32_t x = exc[idx];
        out[((size_t)g << a.log2_m) + (x >> 8)] = (uint8_t)(x & 0xffu);
    }
}
// raw rows of a piece: slot r = blockIdx.x of the raw area -> row raw_idx[r]
__global__ void __launch_bounds__(256) k_apply_nib4_raw(Nib4Pieces a, uint8_t* __restrict__ regs) {
    const int pi = blockIdx.y, r = blockIdx.x;
    const long long rows = a.rows_of(pi);
    const selb::Nib4Piece P = selb::nib4_piece(rows, (size_t)1 << a.log2_m);
    const uint8_t* piece = a.pieces + (size_t)pi * a.piece_stride;
    const int g = reinterpret_cast<const int32_t*>(piece + P.off_rawidx)[r];
    if (g < 0 || g >= rows) return;
    const uint4* raw = reinterpret_cast<const uint4*>(piece + P.off_raw + ((size_t)r << a.log2_m));
    uint4* out = reinterpret_cast<uint4*>(regs + (((size_t)pi * (size_t)a.piece_rows + (size_t)g) << a.log2_m));
    for (int i = threadIdx.x; i < (1 << (a.log2_m - 4)); i += blockDim.x) out[i] = raw[i];
}

__global__ void k_iota_i32(int32_t* v, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) v[i] = (int32_t)i;
}

// sorted cardinalities -> truncated e (size_t e = card, selection.cpp:275,280) + tie detection
__global__ void k_sorted_prep(const double* __restrict__ cards_sorted, long long n, unsigned long long* __restrict__ e,
                              uint32_t* __restrict__ tie_flag) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double cd = cards_sorted[i];
    e[i] = (unsigned long long)cd;
    // a tie sends the load to the reference's own (unstable) std::sort on the host.  Ties at cardinality 0 do not count:
    // empty sketches (and the all-zero padding rows of a sharded load) pair with nothing (e == 0 columns are skipped, an
    // e == 0 row fails CB for every tau > 0), so their order among themselves never reaches the output
    if (i + 1 < n && !(cd < cards_sorted[i + 1]) && cd > 0.) *tie_flag = 1;
}

// per-genome cardinality: hll.h:834-837 (sum) / :1138-1141 (trusted stored value)
__global__ void k_genome_cards(const uint32_t* __restrict__ hist, const double* __restrict__ stored, long long n,
                               int p, double* __restrict__ cards, const uint32_t* __restrict__ max_seen,
                               uint32_t max_ok, uint16_t* __restrict__ grange) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (*max_seen > max_ok) { cards[i] = 0.; grange[i] = 0; return; }   // malformed input: the load fails after the sync
    {   // smallest and largest register value of the genome (window choice of the bit-plane union kernel)
        int vmin = 63, vmax = 0;
        for (int b = 0; b < 64; ++b)
            if (hist[i * 64 + b]) { vmin = min(vmin, b); vmax = max(vmax, b); }
        grange[i] = (uint16_t)(min(vmin, vmax) | (vmax << 8));
    }
    if (stored && stored[i] >= 0.) { cards[i] = stored[i]; return; }
    cards[i] = selb::ertl_mle(hist + i * 64, p);
}

__global__ void k_mle_only(const uint32_t* __restrict__ hist, long long n, int p, double* __restrict__ out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) out[i] = selb::ertl_mle(hist + i * 64, p);
}
// ============================================================================
// load-time re-layout
// ============================================================================
// dst[i][:] = src[order[i]][:], rows of row_words uint32
__global__ void k_gather_rows(const uint32_t* __restrict__ src, const int32_t* __restrict__ order, long long n,
                              int row_words, uint32_t* __restrict__ dst) {
    const long long total = n * row_words;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long i = idx / row_words;
        const int j = (int)(idx - i * row_words);
        dst[idx] = src[(size_t)order[i] * row_words + j];
    }
}

// auxT[j][g] = word j of the aux HLL of the g-th genome in sorted order (pad columns stay 0)
__global__ void k_aux_transpose(const uint32_t* __restrict__ src, const int32_t* __restrict__ order, long long n,
                                long long npad, int row_words, uint32_t* __restrict__ dst) {
    const long long total = n * row_words;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(idx / n);
        const long long g = idx - (long long)j * n;
        dst[(size_t)j * npad + g] = src[(size_t)order[g] * row_words + j];
    }
}
