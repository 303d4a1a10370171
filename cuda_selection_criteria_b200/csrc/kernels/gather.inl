// gather.inl — peer-memory gather of the pair lists over NVLink / CUDA IPC (part of selb200.cu)
// ============================================================================
// Peer-memory gather (multi-GPU, one process per GPU): every rank pushes its emitted (key, J)
// list straight into the ROOT GPU's landing zone with plain stores over NVLink/NVSwitch (the zone
// is mapped into each process with CUDA IPC).  One system-scope atomicAdd claims a contiguous
// block per rank, a second one signals completion; the root spins on its own memory until all
// ranks have signalled, then sorts the merged list.  No NCCL call and no host round trip between
// the emit kernel and the merged result (SURVEY.md §8e "gather of (i,k,J) lists to GPU 0").
//   landing zone: GatherHdr | keys[2][cap] | jac[2][cap] | near_keys[2][ncap] | near_j[2][ncap]
//   two buffers (epoch parity) so that a fast rank may already push run e+1 while the root still
//   merges run e; `consumed` stops it from getting two runs ahead.
// ============================================================================
struct GatherHdr {
    unsigned long long count[2];        // slots claimed per parity
    unsigned long long near_count[2];
    unsigned int done[2];               // ranks whose push is complete, per parity
    unsigned int consumed;              // runs the root has merged (monotone)
    unsigned int error;                 // 1: a wait timed out
    unsigned long long pad[26];
};
static_assert(sizeof(GatherHdr) == 256, "landing-zone header is 256 bytes");

struct GatherPush {                     // local to each rank
    unsigned long long base, near_base;
    unsigned int go, blocks_done;
};

struct GatherZone {                     // pointers into the (local or IPC-mapped) landing zone
    GatherHdr* hdr;
    uint64_t* keys;                     // [2][cap]
    double* jac;
    uint64_t* near_keys;                // [2][near_cap]
    double* near_j;
    unsigned long long cap, near_cap;
};

#ifndef SELB_EMUL   // tests/emul/cuda_emul.h: host clock, and a timeout short enough to test
__device__ __forceinline__ unsigned long long gtime_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
constexpr unsigned long long GATHER_TIMEOUT_NS = 20ull * 1000 * 1000 * 1000;
#endif

// one thread: (optionally) make sure the pass did not overflow, wait until the buffer of this parity
// has been merged by the root two runs ago, claim the rank's block in the root's lists
__global__ void k_gather_claim(GatherZone z, unsigned int epoch, unsigned long long* __restrict__ meta, int check,
                               unsigned long long cand_cap, unsigned long long pair_lim, unsigned long long out_cap,
                               unsigned long long tile_cap, unsigned long long near_cap_local,
                               GatherPush* __restrict__ st) {
    if (threadIdx.x | blockIdx.x) return;
    st->go = 0;
    st->blocks_done = 0;
    if (check && (meta[M_CAND] > cand_cap || meta[M_PAIRS] > pair_lim || meta[M_OUT] > out_cap ||
                  meta[M_TILES] > tile_cap || meta[M_NEAR] > near_cap_local))
        return;                              // the host redoes the pass and pushes afterwards
    if (epoch >= 2) {
        const unsigned long long t0 = gtime_ns();
        while (*(volatile unsigned int*)&z.hdr->consumed + 1u < epoch) {
            if (gtime_ns() - t0 > GATHER_TIMEOUT_NS) { meta[M_PUSHED] = 2; return; }
            __nanosleep(200);
        }
    }
    const unsigned int b = epoch & 1u;
    st->base = atomicAdd_system(&z.hdr->count[b], meta[M_OUT]);
    st->near_base = atomicAdd_system(&z.hdr->near_count[b], min(meta[M_NEAR], near_cap_local));
    st->go = 1;
    meta[M_PUSHED] = 1;
}

// all CTAs: copy the rank's lists into its block of the root's lists; the last CTA signals
__global__ void __launch_bounds__(256)
k_gather_copy(GatherZone z, unsigned int epoch, const unsigned long long* __restrict__ meta,
              unsigned long long near_cap_local, const uint64_t* __restrict__ keys, const double* __restrict__ jac,
              const uint64_t* __restrict__ near_keys, const double* __restrict__ near_j, GatherPush* __restrict__ st) {
    if (!st->go) return;
    const unsigned int b = epoch & 1u;
    const unsigned long long cnt = meta[M_OUT], ncnt = min(meta[M_NEAR], near_cap_local);
    const unsigned long long base = st->base, nbase = st->near_base;
    uint64_t* dk = z.keys + b * z.cap;
    double* dj = z.jac + b * z.cap;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < cnt;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long slot = base + i;
        if (slot < z.cap) { dk[slot] = keys[i]; dj[slot] = jac[i]; }
    }
    uint64_t* nk = z.near_keys + b * z.near_cap;
    double* nj = z.near_j + b * z.near_cap;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < ncnt;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long slot = nbase + i;
        if (slot < z.near_cap) { nk[slot] = near_keys[i]; nj[slot] = near_j[i]; }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(&st->blocks_done, 1u) == gridDim.x - 1) {
            st->blocks_done = 0;
            __threadfence_system();
            atomicAdd_system(&z.hdr->done[b], 1u);
        }
    }
}

// root: wait for every rank's signal, then publish the merged counts where the host can read them
__global__ void k_gather_wait(GatherZone z, unsigned int epoch, unsigned int world, const GatherPush* __restrict__ st,
                              unsigned long long* __restrict__ merged /* [count, near_count, error] */) {
    if (threadIdx.x | blockIdx.x) return;
    if (!st->go) { merged[2] = 2; return; }      // the root's own pass is being redone: nothing to wait for yet
    const unsigned int b = epoch & 1u;
    const unsigned long long t0 = gtime_ns();
    unsigned long long err = 0;
    while (*(volatile unsigned int*)&z.hdr->done[b] < world) {
        if (gtime_ns() - t0 > GATHER_TIMEOUT_NS) { err = 1; z.hdr->error = 1; break; }
        __nanosleep(100);
    }
    __threadfence_system();
    merged[0] = *(volatile unsigned long long*)&z.hdr->count[b];
    merged[1] = *(volatile unsigned long long*)&z.hdr->near_count[b];
    merged[2] = err;
}

// root, after the merge of this parity has been copied out: hand the buffer back
__global__ void k_gather_release(GatherZone z, unsigned int epoch) {
    if (threadIdx.x | blockIdx.x) return;
    const unsigned int b = epoch & 1u;
    z.hdr->count[b] = 0;
    z.hdr->near_count[b] = 0;
    z.hdr->done[b] = 0;
    __threadfence_system();
    *(volatile unsigned int*)&z.hdr->consumed = epoch + 1u;
}
