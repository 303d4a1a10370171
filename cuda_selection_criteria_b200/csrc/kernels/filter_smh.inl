// filter_smh.inl — K3/K4: LSH signatures, smh_a tile pre-filter, exact verification, CB-only enumeration (part of selb200.cu)
// ============================================================================
// K3: LSH band signatures.  For the g-th sorted genome and band b, sig(b,g) = 16 bits of a mix
// of the band's n_rows buckets.  Equal bands => equal signatures, so "some band equal"
// (criteria_sketch.hpp:71-79) implies "some signature equal"; the converse is checked exactly
// by k_smh_verify.  Two bands share one 32-bit word: word w holds bands 2w (low half) and 2w+1.
//   sigR[w][g] =  sig              (row operand)
//   sigC[w][g] = -sig per half     (column operand), so  r + c == 0 (mod 2^16)  <=>  equal
// An odd band count leaves a pad half that can never match (row 0, column 1); pad genomes hold
// row 0 / column 0x0101.
// ============================================================================
__device__ __forceinline__ uint32_t band_sig16(const uint64_t* v, int n_rows) {
    uint64_t h = 0x243F6A8885A308D3ull;
    for (int r = 0; r < n_rows; ++r) h = mix64(h ^ v[r]);
    return (uint32_t)(h >> 48);
}

__global__ void __launch_bounds__(256)
k_smh_signatures(const uint64_t* __restrict__ aux_sorted, long long n, long long npad, int m_aux,
                 int n_rows, int n_bands, uint32_t* __restrict__ sigR, uint32_t* __restrict__ sigC) {
    // thread = (genome, band), band fastest: a warp reads consecutive bands of one genome, i.e. one contiguous
    // run of its sketch; lane pairs then pack two bands into a word
    const int nw = (n_bands + 1) >> 1;
    const int nb2 = nw * 2;                                    // bands rounded up to even (pad band never matches)
    const long long total = npad * nb2;                        // pad genomes included: row halves 0, column halves 0x0101
    const long long stride = (long long)gridDim.x * blockDim.x;          // even: lane pairs stay together
    for (long long idx0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx0 - (threadIdx.x & 31) < total;
         idx0 += stride) {
        const bool live = idx0 < total;
        const long long g = live ? idx0 / nb2 : 0;
        const int b = live ? (int)(idx0 - g * nb2) : 0;
        uint32_t sig = 0;
        const bool real = live && b < n_bands && g < n;
        if (real) sig = band_sig16(aux_sorted + (size_t)g * m_aux + (size_t)b * n_rows, n_rows);
        const uint32_t other = __shfl_down_sync(0xffffffffu, sig, 1);
        const bool other_real = __shfl_down_sync(0xffffffffu, (int)real, 1) != 0;
        if (live && !(b & 1) && g >= n) {                      // pad genome: never matches anything
            const int w = b >> 1;
            sigR[(size_t)w * npad + g] = 0u;
            sigC[(size_t)w * npad + g] = 0x01010101u;
        } else if (live && !(b & 1)) {
            const uint32_t r1 = other_real ? other : 0u;
            const uint32_t c1 = other_real ? ((0u - other) & 0xffffu) : 1u;   // pad half: row 0, column 1
            const int w = b >> 1;
            sigR[(size_t)w * npad + g] = sig | (r1 << 16);
            sigC[(size_t)w * npad + g] = ((0u - sig) & 0xffffu) | (c1 << 16);
        }
    }
}

// ============================================================================
// K4: smh_a tile pre-filter.  One CTA = one 128x128 tile of the sorted pair space,
// 256 threads, each an 8x8 register micro-tile.  Per signature word (two bands): 4 LDS.128 and
// 64 x VIADDMNMX.U16x2 (acc = min(acc, r + c) per 16-bit half); a zero half at the end
// <=> some band signature matched.  Candidates (rare) leave through warp-aggregated atomics.
// ============================================================================
// Accumulate: acc = min(acc, r + c) per 16-bit half in ONE instruction (VIADDMNMX.U16x2); the column
// operand holds the negated halves, so a half reaches 0 exactly when the two signatures are equal.
// Measured alternatives on B200 (n=100k, 4.66e8 CB pairs): XOR+MIN on 32-bit signatures 1.22 ms;
// (LOP3, IADD, LOP3) zero-half test on packed halves 1.07 ms; this form 0.86 ms.
#ifndef SELB_EMUL   // tests/emul/cuda_emul.h supplies host versions
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
#endif   // SELB_EMUL

__global__ void __launch_bounds__(256, FILTER_CTAS_PER_SM)
k_tile_filter_smh(const uint32_t* __restrict__ sigR, const uint32_t* __restrict__ sigC, long long npad, int n_words,
                  TileWalk tw, const int32_t* __restrict__ lo, const int32_t* __restrict__ hi, int n,
                  uint2* __restrict__ cand, unsigned long long* __restrict__ cand_count,
                  unsigned long long cand_cap) {
    // three buffers: the signature words of the next two (tile, chunk) items stream in with cp.async while
    // the current one is being compared — a tile's 8 KiB arrive in about the time its 512 instructions per
    // thread take, so without the overlap the ALU pipe idles half the time; with three buffers ONE barrier
    // per item both publishes the item's copies and frees the buffer of the item before it
    constexpr int NBUF = 3;
    __shared__ __align__(16) uint32_t sR[NBUF][SIG_CHUNK][TILE];
    __shared__ __align__(16) uint32_t sC[NBUF][SIG_CHUNK][TILE];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int jend = tw.count();
    const int nchunk = (n_words + SIG_CHUNK - 1) / SIG_CHUNK;
    // persistent CTAs: the shard's tile count lives in device memory, so no host sync sizes the grid
    const int j_first = tw.j0 + (int)blockIdx.x;
    if (j_first >= jend) return;
    const int my_tiles = (jend - 1 - j_first) / (int)gridDim.x + 1;
    const int n_items = my_tiles * nchunk;                 // items = (tile, chunk of SIG_CHUNK words), no divisions below

    auto tile_at = [&](int t) -> int2 { return t < my_tiles ? tw.tile(j_first + t * (int)gridDim.x) : make_int2(0, 0); };
    auto stage = [&](int ch, int2 rc, int buf) {           // queue the loads of one item
        const int b0 = ch * SIG_CHUNK;
        const int nb = min(SIG_CHUNK, n_words - b0);
        const int r0 = rc.x * TILE, c0 = rc.y * TILE;
        for (int idx = tid; idx < nb * 64; idx += 256) {   // nb words x (128 row + 128 col) / 4 per copy
            const int bb = idx >> 6, part = idx & 63, x = (part & 31) * 4;
            if (part < 32) cp_async16(&sR[buf][bb][x], sigR + (size_t)(b0 + bb) * npad + r0 + x);
            else cp_async16(&sC[buf][bb][x], sigC + (size_t)(b0 + bb) * npad + c0 + x);
        }
        cp_async_commit();
    };
    // zero 16-bit half somewhere in x
    auto has_zero_half = [](uint32_t x) { return ((x - 0x00010001u) & ~x & 0x80008000u) != 0u; };

    uint32_t acc[8][8];
    // (t, ch) = item being compared; (ts, chs) = the next item to copy, two items ahead.  The coordinates of the
    // copy cursor's tile and of the tile after it are fetched ahead of use (rc_s, rc_sn).
    int t = 0, ch = 0, buf = 0;
    int ts = 0, chs = 0, bufs = 0;
    int2 rc_s = tile_at(0), rc_sn = tile_at(1);
    __shared__ int2 s_rc[4];                               // coordinates of the tiles in flight, by tile number & 3
    auto stage_next = [&]() {
        if (chs == 0 && tid == 0) s_rc[ts & 3] = rc_s;
        stage(chs, rc_s, bufs);
        bufs = bufs + 1 == NBUF ? 0 : bufs + 1;
        if (++chs == nchunk) {
            chs = 0;
            ++ts;
            rc_s = rc_sn;
            rc_sn = tile_at(ts + 1);
        }
    };
    int staged = 0;
    for (; staged < 2 && staged < n_items; ++staged) stage_next();
    for (int item = 0; item < n_items; ++item) {
        if (item + 1 < staged) cp_async_wait<1>();          // everything but the newest group has landed
        else cp_async_wait<0>();
        __syncthreads();
        if (staged < n_items) { stage_next(); ++staged; }   // into the buffer item-1 was compared from
        const int nb = min(SIG_CHUNK, n_words - ch * SIG_CHUNK);
        int bb = 0;
        if (ch == 0) {                                      // first word of a tile: no accumulator to read
            const uint4 ra = *reinterpret_cast<const uint4*>(&sR[buf][0][ty * 8]);
            const uint4 rb = *reinterpret_cast<const uint4*>(&sR[buf][0][ty * 8 + 4]);
            const uint4 ca = *reinterpret_cast<const uint4*>(&sC[buf][0][tx * 4]);
            const uint4 cb = *reinterpret_cast<const uint4*>(&sC[buf][0][64 + tx * 4]);
            const uint32_t rs[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
            const uint32_t cs[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int b = 0; b < 8; ++b) acc[a][b] = __viaddmin_u16x2(rs[a], cs[b], 0xffffffffu);
            bb = 1;
        }
        for (; bb < nb; ++bb) {
            const uint4 ra = *reinterpret_cast<const uint4*>(&sR[buf][bb][ty * 8]);
            const uint4 rb = *reinterpret_cast<const uint4*>(&sR[buf][bb][ty * 8 + 4]);
            const uint4 ca = *reinterpret_cast<const uint4*>(&sC[buf][bb][tx * 4]);
            const uint4 cb = *reinterpret_cast<const uint4*>(&sC[buf][bb][64 + tx * 4]);
            const uint32_t rs[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
            const uint32_t cs[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int b = 0; b < 8; ++b) acc[a][b] = __viaddmin_u16x2(rs[a], cs[b], acc[a][b]);
        }
        buf = buf + 1 == NBUF ? 0 : buf + 1;
        if (++ch < nchunk) continue;
        const int t_done = t;
        ch = 0;
        ++t;
        // per-row minima first: a 16-bit signature collides by chance once per ~64 thread-tiles, so four
        // warps in ten come here with ONE row to look at, not 64 cells
        uint32_t rowmin[8];
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            const uint32_t m0 = __vimin3_u16x2(acc[a][0], acc[a][1], acc[a][2]);
            const uint32_t m1 = __vimin3_u16x2(acc[a][3], acc[a][4], acc[a][5]);
            rowmin[a] = __vimin3_u16x2(m0, m1, __vminu2(acc[a][6], acc[a][7]));
        }
        const uint32_t any = __vimin3_u16x2(__vimin3_u16x2(rowmin[0], rowmin[1], rowmin[2]),
                                            __vimin3_u16x2(rowmin[3], rowmin[4], rowmin[5]),
                                            __vminu2(rowmin[6], rowmin[7]));
        if (!has_zero_half(any)) continue;
        const int2 rc = s_rc[t_done & 3];                   // written when the tile was queued, barriers ago
        const int r0 = rc.x * TILE, c0 = rc.y * TILE;
        unsigned long long cells = 0ull;                    // bit a*8+b: cell (a,b) has a matching band signature
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            if (!has_zero_half(rowmin[a])) continue;
            uint32_t rowbits = 0;
#pragma unroll
            for (int b = 0; b < 8; ++b) rowbits |= has_zero_half(acc[a][b]) ? (1u << b) : 0u;
            cells |= (unsigned long long)rowbits << (a * 8);
        }
        while (cells) {
            const int bit = __ffsll((long long)cells) - 1;
            cells &= cells - 1;
            const int a = bit >> 3, b = bit & 7;
            const int i = r0 + ty * 8 + a;
            const int k = c0 + (b < 4 ? tx * 4 + b : 64 + tx * 4 + (b - 4));
            if (i >= n || k < lo[i] || k > hi[i]) continue;
            const unsigned long long slot = warp_claim(cand_count);
            if (slot < cand_cap) cand[slot] = make_uint2((uint32_t)i, (uint32_t)k);
        }
    }
}

// buckets of band b of the two genomes equal?  All loads are issued before the first compare (a chain of load, compare,
// branch per bucket costs a DRAM round trip each: the sketches are not cache-resident)
__device__ __forceinline__ bool smh_band_equal(const uint64_t* __restrict__ v1, const uint64_t* __restrict__ v2, int b, int n_rows) {
    const uint64_t* p1 = v1 + (size_t)b * n_rows;
    const uint64_t* p2 = v2 + (size_t)b * n_rows;
    if (!(n_rows & 1) && !(((size_t)b * n_rows) & 1)) {            // 16-byte aligned rows of pairs (sketch rows are 8m bytes, m even here)
        const ulonglong2* q1 = reinterpret_cast<const ulonglong2*>(p1);
        const ulonglong2* q2 = reinterpret_cast<const ulonglong2*>(p2);
        unsigned long long diff = 0ull;
        int r = 0;
        for (; r + 4 <= (n_rows >> 1); r += 4) {
            const ulonglong2 a0 = __ldg(q1 + r), a1 = __ldg(q1 + r + 1), a2 = __ldg(q1 + r + 2), a3 = __ldg(q1 + r + 3);
            const ulonglong2 c0 = __ldg(q2 + r), c1 = __ldg(q2 + r + 1), c2 = __ldg(q2 + r + 2), c3 = __ldg(q2 + r + 3);
            diff |= (a0.x ^ c0.x) | (a0.y ^ c0.y) | (a1.x ^ c1.x) | (a1.y ^ c1.y) | (a2.x ^ c2.x) | (a2.y ^ c2.y) | (a3.x ^ c3.x) | (a3.y ^ c3.y);
            if (diff) return false;
        }
        for (; r < (n_rows >> 1); ++r) {
            const ulonglong2 a0 = __ldg(q1 + r), c0 = __ldg(q2 + r);
            diff |= (a0.x ^ c0.x) | (a0.y ^ c0.y);
        }
        return diff == 0ull;
    }
    for (int r = 0; r < n_rows; ++r)
        if (__ldg(p1 + r) != __ldg(p2 + r)) return false;
    return true;
}

// exact smh_a on the candidates: include/criteria_sketch.hpp:66-81.  Thread per candidate.
// A band can only be equal if its 16-bit signatures are, so the thread re-reads the (L2-resident)
// signature words of both genomes, and compares bucket by bucket only the bands whose signatures
// match, stopping at the first band that is really equal: ~2 x 64 B of auxiliary sketch per
// candidate instead of 2 x 8m B.
__global__ void __launch_bounds__(256)
k_smh_verify(const uint64_t* __restrict__ aux_sorted, const uint32_t* __restrict__ sigR, long long npad, int m_aux,
             int n_rows, int n_bands, const uint2* __restrict__ cand,
             const unsigned long long* __restrict__ ncand_dev, unsigned long long cand_cap,
             uint2* __restrict__ pairs, unsigned long long* __restrict__ pair_count, unsigned long long pair_cap) {
    const long long ncand = (long long)min(*ncand_dev, cand_cap);
    const int n_words = (n_bands + 1) >> 1;
    for (long long ci = blockIdx.x * (long long)blockDim.x + threadIdx.x; ci < ncand;
         ci += (long long)gridDim.x * blockDim.x) {
        const uint2 pr = cand[ci];
        const uint64_t* v1 = aux_sorted + (size_t)pr.x * m_aux;
        const uint64_t* v2 = aux_sorted + (size_t)pr.y * m_aux;
        bool hit = false;
        // signature words four at a time (independent loads), then the bands whose halves agree, exactly
        for (int w0 = 0; w0 < n_words && !hit; w0 += 4) {
            uint32_t x[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                x[u] = w0 + u < n_words ? (__ldg(sigR + (size_t)(w0 + u) * npad + pr.x) ^ __ldg(sigR + (size_t)(w0 + u) * npad + pr.y)) : 0xffffffffu;
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int b = 2 * (w0 + u) + half;
                    if (hit || b >= n_bands || ((x[u] >> (16 * half)) & 0xffffu) != 0) continue;
                    hit = smh_band_equal(v1, v2, b, n_rows);
                }
        }
        if (hit) {
            const unsigned long long slot = warp_claim(pair_count);
            if (slot < pair_cap) pairs[slot] = pr;
        }
    }
}

// ============================================================================
// K3'/K4': smh_a as an EQUALITY JOIN (default below four shards; the tile filter above stays for more shards and as
// SELB200_SMHFILTER=tiles).
// "Some band equal" (criteria_sketch.hpp:71-79) is a join on (band, band contents): instead of testing all P_cb pairs
// of the band against all bands (O(P_cb * bands): 3.7e9 half-word operations at n = 100k), the n * bands keys
// (band << sbits | top sbits of the 16-bit signature of the band's buckets; sbits = 16 up to 256 bands, fewer beyond so
// that the bucket table stays below 2^24 entries) are bucketed by a counting sort and every bucket is walked:
// work O(n * bands + matches).
//   k_smh_sigkeys   : key of every (genome, band), its rank inside its bucket (the atomicAdd that counts the bucket),
//                     and the full signatures genome-major (sigG: two bands per word, nbw words per genome = 32 B at 16 bands)
//   exclusive scan  : bucket offsets (cub)
//   k_smh_scatter   : positions into their buckets (order inside a bucket: arbitrary)
//   k_smh_join_expand + k_smh_join : for the element (band b, genome i) the members k of its bucket with i < k <= hi(i) are
//                     the pairs of the CB band whose band-b keys agree (every unordered pair exactly once, from its smaller
//                     position); every one is written out as an item and handled by a thread of its own (a bucket of a
//                     thousand identical bands is 5e5 items for as many threads, not a serial walk).  A pair is handled
//                     ONCE, by the first band whose KEYS agree (the thread reads the earlier bands' signatures of both
//                     genomes: 32 B each): that handler compares the buckets of every band from b on whose 16-bit signatures
//                     agree, exactly, and emits the pair at the first band that really is equal — an equal band has equal
//                     signatures, hence equal keys, hence lies at or behind the handler's band: the decision is that of
//                     k_smh_verify, so P_aux is the reference's.
// (First version: stable radix sort of (key, position) + galloping search for the followers: 86 + 56 us at n = 100k
// against 8 + 5 + 40 for scan, scatter and this expansion.)
// Shards: keys and buckets are replicated (0.06 ms at n = 100k), expansion and item walk cover the shard's own rows
// (i mod n_shards == shard): they divide by the shard count.
// ============================================================================
__global__ void __launch_bounds__(256)
k_smh_sigkeys(const uint64_t* __restrict__ aux_sorted, long long n, int m_aux, int n_rows, int n_bands, int nbw, int sbits,
              uint32_t* __restrict__ keys, uint32_t* __restrict__ rank, uint32_t* __restrict__ bucket_cnt,
              uint32_t* __restrict__ sigG) {
    const int nb2 = nbw * 2;
    const long long total = n * nb2;
    const long long stride = (long long)gridDim.x * blockDim.x;          // even: lane pairs stay together
    for (long long idx0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx0 - (threadIdx.x & 31) < total;
         idx0 += stride) {
        const bool live = idx0 < total;
        const long long g = live ? idx0 / nb2 : 0;
        const int b = live ? (int)(idx0 - g * nb2) : 0;
        const bool real = live && b < n_bands;
        uint32_t sig = 0;
        if (real) {
            sig = band_sig16(aux_sorted + (size_t)g * m_aux + (size_t)b * n_rows, n_rows);
            const uint32_t key = ((uint32_t)b << sbits) | (sig >> (16 - sbits));
            keys[g * n_bands + b] = key;
            rank[g * n_bands + b] = atomicAdd(bucket_cnt + key, 1u);
        }
        const uint32_t other = __shfl_down_sync(0xffffffffu, sig, 1);
        if (live && !(b & 1)) sigG[g * nbw + (b >> 1)] = sig | (other << 16);
    }
}

// element e = (genome, band) -> its slot of its bucket
__global__ void __launch_bounds__(256)
k_smh_scatter(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ rank, const uint32_t* __restrict__ bucket_off,
              long long n_keys, int n_bands, uint32_t* __restrict__ members) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n_keys; e += (long long)gridDim.x * blockDim.x)
        members[bucket_off[keys[e]] + rank[e]] = (uint32_t)(e / n_bands);
}

// element e = (genome i, band b) in [e0, e1): the members k of its bucket with i < k <= hi(i) become ITEMS {i, k, band}.  A warp
// claims the room for its lanes' items with one atomic and the lanes write them — the walk itself is then one thread per
// item (k_smh_join), whatever the size of the buckets.  Most buckets hold one element: settled by two loads.
__global__ void __launch_bounds__(256)
k_smh_join_expand(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ bucket_off, const uint32_t* __restrict__ members,
                  long long e0, long long e1, int n_bands, int sbits, const int32_t* __restrict__ lo, const int32_t* __restrict__ hi,
                  uint4* __restrict__ items, unsigned long long* __restrict__ item_count, unsigned long long item_cap,
                  int shard = 0, int n_shards = 1, long long n = 0) {
    // Shards: a pair belongs to the shard of its ROW, i mod n_shards — this shard expands the elements of its own rows only
    // (rows shard, shard + n_shards, ...: the walk below runs over that compact sequence, 16 bands of two rows to a warp),
    // so expansion, items and item walk all divide by the shard count; rows are in cardinality order, cluster mates sit next
    // to one another, and dealing them round-robin balances the pairs.  n: number of genomes (only read when n_shards > 1).
    const int lane = threadIdx.x & 31;
    const long long t_end = n_shards > 1 ? ((n - shard + n_shards - 1) / n_shards) * n_bands : e1 - e0;
    for (long long tb = blockIdx.x * (long long)blockDim.x + threadIdx.x - lane; tb < t_end; tb += (long long)gridDim.x * blockDim.x) {
        const long long t = tb + lane;
        long long e = e0 + t;                            // one shard: the elements of [e0, e1) in order
        if (n_shards > 1) {
            const long long li = t / n_bands;
            e = ((long long)shard + li * n_shards) * n_bands + (t - li * n_bands);
        }
        unsigned long long c = 0;
        uint32_t key = 0, a = 0, b = 0;
        int i = 0, hi_i = -1;
        if (t < t_end && e >= e0 && e < e1) {
            key = keys[e];
            a = bucket_off[key]; b = bucket_off[key + 1];
            if (b - a > 1u) {
                i = (int)(e / n_bands);
                hi_i = hi[i];
                if (hi_i >= lo[i])                      // lo(i) = max(i + 1, first genome with e > 0): implied by k > i for such rows
                    for (uint32_t t = a; t < b; ++t) {
                        const int k = (int)members[t];
                        c += (k > i && k <= hi_i) ? 1u : 0u;
                    }
            }
        }
        unsigned long long pre = c;                    // inclusive scan over the lanes
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long v = __shfl_up_sync(0xffffffffu, pre, o);
            if (lane >= o) pre += v;
        }
        const unsigned long long total = __shfl_sync(0xffffffffu, pre, 31);
        if (total == 0ull) continue;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(item_count, total);
        base = __shfl_sync(0xffffffffu, base, 0) + (pre - c);
        if (c) {
            unsigned long long d = 0;
            for (uint32_t t = a; t < b; ++t) {
                const int k = (int)members[t];
                if (k > i && k <= hi_i) {
                    if (base + d < item_cap) items[base + d] = make_uint4((uint32_t)i, (uint32_t)k, key >> sbits, 0u);
                    ++d;
                }
            }
        }
    }
}

// one thread per item {i, k, band}: see the header above
__global__ void __launch_bounds__(256)
k_smh_join(const uint4* __restrict__ items, const unsigned long long* __restrict__ item_count, unsigned long long item_cap,
           const uint32_t* __restrict__ sigG, int nbw, int sbits, const uint64_t* __restrict__ aux_sorted, int m_aux, int n_rows, int n_bands,
           uint2* __restrict__ pairs, unsigned long long* __restrict__ pair_count,
           unsigned long long pair_cap, unsigned long long* __restrict__ cand_count, unsigned long long* __restrict__ item_max) {
    const unsigned long long n_items = min(*item_count, item_cap);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicMax(item_max, *item_count);     // the host grows the list and redoes the pass if it overflowed
    __shared__ uint32_t s_warp_hits[8];
    __shared__ unsigned long long s_base;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t n_cand = 0;
    // the loop is uniform over the CTA (block barriers inside): every thread runs the same number of rounds
    for (unsigned long long w0 = blockIdx.x * (unsigned long long)blockDim.x; w0 < n_items; w0 += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long w = w0 + threadIdx.x;
        bool hit = false;
        int i = 0, k = 0;
        if (w < n_items) {
            const uint4 it = __ldg(items + w);
            i = (int)it.x; k = (int)it.y;
            const int bnd = (int)it.z;
            {   // (every item of the list is this shard's: the expansion only wrote the items of its own rows)
                const uint32_t* si = sigG + (size_t)i * nbw;
                const uint32_t* sk = sigG + (size_t)k * nbw;
                // an earlier band with equal KEYS (the top sbits of the signatures: what the buckets are made of, so an item
                // exists for it) handles the pair (words of four at a time: independent loads)
                const uint32_t kmask = (0xffffu << (16 - sbits)) & 0xffffu;
                bool earlier = false;
                uint32_t x_own = 0u;
                for (int wd = 0; wd <= (bnd >> 1) && !earlier; wd += 4) {
                    uint32_t x[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) x[u] = wd + u <= (bnd >> 1) ? (__ldg(si + wd + u) ^ __ldg(sk + wd + u)) : 0xffffffffu;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (2 * (wd + u) < bnd && (x[u] & kmask) == 0u) earlier = true;
                        if (2 * (wd + u) + 1 < bnd && ((x[u] >> 16) & kmask) == 0u) earlier = true;
                        if (wd + u == (bnd >> 1)) x_own = x[u];
                    }
                }
                if (!earlier) {
                    ++n_cand;
                    const uint64_t* v1 = aux_sorted + (size_t)i * m_aux;
                    const uint64_t* v2 = aux_sorted + (size_t)k * m_aux;
                    // equal buckets need equal 16-bit signatures: with coarse keys (sbits < 16) this band may already be out
                    hit = ((x_own >> (16 * (bnd & 1))) & 0xffffu) == 0u && smh_band_equal(v1, v2, bnd, n_rows);
                    for (int b2 = bnd + 1; b2 < n_bands && !hit; ++b2) {
                        const uint32_t x = __ldg(si + (b2 >> 1)) ^ __ldg(sk + (b2 >> 1));
                        if (((x >> (16 * (b2 & 1))) & 0xffffu) == 0u) hit = smh_band_equal(v1, v2, b2, n_rows);
                    }
                }
            }
        }
        // one claim per CTA and round: hits are common here (half a million on the bench), and claims of one word
        // from every warp serialise in the L2
        const uint32_t bal = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) s_warp_hits[wid] = (uint32_t)__popc(bal);
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t tot = 0;
            for (int q = 0; q < 8; ++q) { const uint32_t c = s_warp_hits[q]; s_warp_hits[q] = tot; tot += c; }
            s_base = tot ? atomicAdd(pair_count, (unsigned long long)tot) : 0ull;
        }
        __syncthreads();
        if (hit) {
            const unsigned long long slot = s_base + s_warp_hits[wid] + (unsigned long long)__popc(bal & ((1u << lane) - 1u));
            if (slot < pair_cap) pairs[slot] = make_uint2((uint32_t)i, (uint32_t)k);
        }
        __syncthreads();                               // s_warp_hits / s_base are rewritten in the next round
    }
    // one atomic per warp: 600 k threads adding to one word is what the kernel's time was (same-address atomics serialise)
    for (int o = 16; o; o >>= 1) n_cand += __shfl_xor_sync(0xffffffffu, n_cand, o);
    if ((threadIdx.x & 31) == 0 && n_cand) atomicAdd(cand_count, (unsigned long long)n_cand);
}

// CB only: every pair of the band inside this tile
__global__ void __launch_bounds__(256)
k_tile_enum(TileWalk tw, const int32_t* __restrict__ lo, const int32_t* __restrict__ hi, int n,
            uint2* __restrict__ pairs, unsigned long long* __restrict__ pair_count, unsigned long long pair_cap) {
    const int jend = tw.count();
    for (int j = tw.j0 + (int)blockIdx.x; j < jend; j += (int)gridDim.x) {
        const int2 rc = tw.tile(j);
        const int r0 = rc.x * TILE, c0 = rc.y * TILE;
        for (int idx = threadIdx.x; idx < TILE * TILE; idx += 256) {
            const int i = r0 + (idx >> 7), k = c0 + (idx & (TILE - 1));
            if (i >= n || k >= n) continue;
            if (k < lo[i] || k > hi[i]) continue;
            const unsigned long long slot = warp_claim(pair_count);
            if (slot < pair_cap) pairs[slot] = make_uint2((uint32_t)i, (uint32_t)k);
        }
    }
}
