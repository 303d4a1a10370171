/*
 * selb200.h — C-ABI of the B200-native all-pairs genome *selection* hot path.
 *
 * Drop-in boundary for sanhue903/CUDA_Selection_Criteria (citations relative to the
 * reference tree):
 *   - process level: the CLIs `selection` (src/selection.cpp:70-303) and
 *     `selection_cuda` (src/selection_cuda.cpp:59-189) call, after loading the sketch
 *     files, exactly one "compare everything" step.  selb200_load_* + selb200_run
 *     replace src/selection.cpp:241-291 (cardinalities, sort, CB, criterion, HLL
 *     union, Jaccard >= tau) and src/selection_cuda.cpp:111-180.
 *   - link level: src/selection_kernels_wrapper.hpp:11-45 (launch_kernel_smh /
 *     launch_kernel_CBsmh, C++ linkage).  Shims with those exact mangled names are
 *     exported by libselb200.so itself (see selb200_shims.h / INTEGRATION.md).
 *
 * Plain pointers and sizes only.  All functions return 0 on success or a negative
 * SELB200_E* code; selb200_last_error() gives the message of the last failure on the
 * calling thread.  There is no CPU fallback: every compute entry point needs a CUDA
 * device and fails with SELB200_ECUDA otherwise.
 */
#ifndef SELB200_H
#define SELB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SELB200_ABI_VERSION 1

enum {
    SELB200_OK = 0,
    SELB200_EINVAL = -1,   /* bad argument / malformed sketch (register > 64-p+1, mixed p, ...) */
    SELB200_ECUDA = -2,    /* CUDA runtime error or no device */
    SELB200_ESTATE = -3,   /* call order violated (run before load, ...) */
    SELB200_ENOMEM = -4,
    SELB200_EUNSUPPORTED = -5 /* estimator other than ERTL_MLE stored in a sketch header */
};

/* Selection criterion applied after the cardinality bound (CB).
 * Reference: `-c` flag, src/selection.cpp:107-109,126,180,227,292-294. */
enum {
    SELB200_CRIT_CB = 0,     /* CB only (README.md:86 "no criterion"; BASELINE config 2) */
    SELB200_CRIT_SMH_A = 1,  /* include/criteria_sketch.hpp:66-81 */
    SELB200_CRIT_HLL_A = 2,  /* include/criteria_sketch.hpp:60-64 + 36-43 */
    SELB200_CRIT_HLL_AN = 3  /* include/criteria_sketch.hpp:52-58 + 22-34 */
};

/* Kind of auxiliary sketch handed to selb200_load_*. */
enum {
    SELB200_AUX_NONE = 0,
    SELB200_AUX_SMH = 1,  /* uint64 [n][aux_len]   (.smh<m> files, src/selection.cpp:12-33) */
    SELB200_AUX_HLL = 2   /* uint8  [n][2^aux_len] (.hll_<p> files, sketch/include/sketch/hll.h:1126-1143) */
};

typedef struct selb200_ctx selb200_ctx;

/* Parameters of one selection run.  Defaults follow src/selection.cpp:76-81. */
typedef struct selb200_params {
    float tau;        /* -h threshold, a FLOAT in the reference (selection.cpp:81)            */
    int32_t criterion;/* SELB200_CRIT_*                                                        */
    float z_score;    /* 1.96f (selection.cpp:76)                                              */
    int32_t order_n;  /* 1 (selection.cpp:77)                                                  */
    int32_t n_rows;   /* smh_a band shape; 0,0 = derive with selb200_band_params(cpu_variant=1) */
    int32_t n_bands;
    int32_t shard;    /* this process's shard: tiles shard, shard+n_shards, ... of the tile list */
    int32_t n_shards; /* 1 = whole pair space                                                  */
    int32_t sort_output; /* 1: order results by (i,k) like the reference prints them          */
    int32_t no_cb;    /* 1: skip the cardinality bound (the "smh_a" loop of experiments/src/time_smh.cpp:229-257; the e2==0 skip stays) */
    int32_t gather;   /* 1: push this shard's pairs into the root GPU's landing zone (selb200_gather_*); the
                         root's results are then the merged list of all n_shards ranks                     */
    int32_t host_results; /* 1: the result lists are also copied to pinned host memory inside the run
                             (selb200_result_host; copy_results then needs no further device copy)     */
    int32_t reserved[4];
} selb200_params;

/* Per-run statistics (stage counts + per-kernel device times). */
typedef struct selb200_stats {
    int64_t n;            /* genomes                                              */
    int64_t pairs_total;  /* n(n-1)/2 (whole job, not per shard)                   */
    int64_t pairs_cb;     /* pairs inside the CB band (whole job)                  */
    int64_t pairs_cb_shard;/* CB-band pairs inside this shard's tiles              */
    int64_t pairs_cand;   /* shard: candidates after the tile pre-filter           */
    int64_t pairs_aux;    /* shard: pairs passing the exact auxiliary criterion    */
    int64_t pairs_out;    /* shard: emitted pairs (J >= tau)                       */
    int64_t pairs_near;   /* shard: evaluated pairs with |J - tau| <= 1e-6*|tau|   */
    int64_t tiles_total;  /* whole job                                             */
    int64_t tiles_shard;
    int32_t n_bands, n_rows;
    int32_t batches;      /* filter->union passes this run needed                  */
    int32_t launches;     /* kernels of this library launched by the run           */
    float ms_bounds;      /* CUDA-event times on the run stream, summed over batches */
    float ms_filter;
    float ms_verify;
    float ms_union;       /* HLL register-max + histogram kernel                   */
    float ms_estimate;    /* MLE + Jaccard + emit                                  */
    float ms_sort;
    float ms_total;       /* first launch to last, device time                     */
    int32_t reserved0;
    int64_t filter_steps; /* hll_a / hll_an plane filter, first pass: warp steps executed (one step = 64 auxiliary
                             registers of 32 pairs); below pairs_cb_shard/32 * 2^p_aux/64 when steps stop early */
    int32_t reserved[5];
} selb200_stats;

/* ---- library / device ---------------------------------------------------- */
int selb200_abi_version(void);
const char* selb200_last_error(void);
int selb200_device_count(void);

/* ---- context --------------------------------------------------------------
 * One context per process per GPU.  `device` is a CUDA ordinal.  `stream` is a
 * cudaStream_t (NULL = a stream owned by the context). */
int selb200_create(int device, void* stream, selb200_ctx** out);
void selb200_destroy(selb200_ctx* ctx);

/* ---- load ------------------------------------------------------------------
 * Sketches arrive in FILE-LIST order (the order of the `-l` list).
 *   regs   uint8 [n][2^p]  primary HLL registers (p = 14 for build_sketch output)
 *   stored double [n] or NULL: the `value_` field of each .hll header; entries >= 0
 *          are trusted like hll.h:1138-1141 does, anything else is recomputed
 *   aux    per aux_kind, see above; NULL with SELB200_AUX_NONE
 * The call computes per-genome histograms and Ertl-MLE cardinalities on the device
 * (hll.h:834-837,564-581,628-688), sorts by cardinality with the reference's
 * comparator (selection.cpp:251-256) and lays the auxiliary sketches out for the
 * tile kernels.  load_host copies from host memory (pinned or pageable);
 * load_device BORROWS device pointers, which must stay valid until the next load
 * or destroy. */
int selb200_load_host(selb200_ctx* ctx, int64_t n, int p, const uint8_t* regs, const double* stored,
                      int aux_kind, int aux_len, const void* aux);
int selb200_load_device(selb200_ctx* ctx, int64_t n, int p, const uint8_t* d_regs, const double* stored_host,
                        int aux_kind, int aux_len, const void* d_aux);

/* Streaming load (what the C++ CLI uses): the caller decodes sketch files straight into pinned
 * staging slots, chunk by chunk, while the device copy, validation, histogram and cardinality of
 * earlier chunks are already running.
 *   begin   : declares n, p and the aux kind; *rows_per_chunk = largest chunk (64 MiB of registers)
 *   acquire : rows [g0, g0+count) must follow the previous chunk; returns host pointers to fill:
 *             regs uint8[count][2^p], stored double[count] (value_ of each header, <0 = recompute),
 *             aux per aux_kind (NULL pointer returned for SELB200_AUX_NONE).  Blocks only when
 *             all three slots are still being copied.
 *   commit  : queues the H2D copies + device work of the acquired rows (asynchronous)
 *   end     : sort by cardinality, aux re-layout; afterwards selb200_run may be called
 * Equivalent to one selb200_load_host call with the concatenated rows. */
int selb200_load_begin(selb200_ctx* ctx, int64_t n, int p, int aux_kind, int aux_len, int64_t* rows_per_chunk);
int selb200_load_acquire(selb200_ctx* ctx, int64_t g0, int64_t count, uint8_t** regs, double** stored, void** aux);
int selb200_load_commit(selb200_ctx* ctx);
int selb200_load_end(selb200_ctx* ctx);

/* Device matrices that become complete piece by piece — e.g. while the other ranks' slices are still
 * arriving over NVLink.  begin BORROWS the (full-size) device matrices like selb200_load_device; every
 * selb200_load_device_rows call declares rows [g0, g0+count) complete IN THE ORDER OF THE CONTEXT'S STREAM
 * (enqueue it behind the copy / collective that fills them, on that stream) and queues their validation,
 * histograms, cardinalities and bit planes without synchronising; each row must be declared exactly once;
 * selb200_load_end (above) then sorts.  Stored header cardinalities are not supported on this route. */
int selb200_load_device_begin(selb200_ctx* ctx, int64_t n, int p, const uint8_t* d_regs, int aux_kind, int aux_len,
                              const void* d_aux);
int selb200_load_device_rows(selb200_ctx* ctx, int64_t g0, int64_t count);

/* Packed transport of register rows (csrc/hostpack.h).  HLL registers of one sketch sit in a narrow band above the
 * smallest one, so a row crosses PCIe / NVLink as one base byte + 4-bit offsets + a short exception list: 0.51 of its
 * bytes.  selb200_load_host packs, copies and unpacks internally (SELB200_H2D=raw turns that off, =packed packs every
 * piece even when the link idles); these three entry
 * points expose the same pieces to a caller that moves the bytes itself (cuda_selection_criteria_b200/dist.py: per-rank
 * slices, H2D + NCCL all-gather of the packed pieces):
 *   piece_bytes : size of the buffer holding `rows` packed rows
 *   pack_piece  : host rows -> piece (threads <= 0: hardware threads / LOCAL_WORLD_SIZE, or SELB200_PACK_THREADS).
 *                 Returns the number of rows that had to be kept as raw bytes (more than 32 registers 15 or more above
 *                 the row's smallest): up to 4 of them travel inside the piece, a return value > 4 means the piece
 *                 cannot be used (send those rows raw); negative = error
 *   load_device_rows_packed : like selb200_load_device_rows, for rows that arrived as pieces in device memory:
 *                 ceil(count / piece_rows) pieces of piece_rows rows each (the last one shorter, packed as such),
 *                 selb200_nib4_piece_bytes(piece_rows, p) bytes apart; the rows are unpacked INTO the matrix given to
 *                 selb200_load_device_begin (which must be writable)
 * Unpacking reproduces the bytes exactly: nothing downstream can depend on the transport. */
/* What the last selb200_load_host moved over PCIe for the REGISTER matrix (auxiliary sketches always travel as they are):
 * bytes copied, rows that went packed, rows that went raw.  The loader packs a piece unless the copy engine has run out
 * of work, in which case the piece goes raw (pinned source memory only) — host packing and PCIe run side by side. */
int selb200_load_info(selb200_ctx* ctx, int64_t* h2d_register_bytes, int64_t* rows_packed, int64_t* rows_raw);
int64_t selb200_nib4_piece_bytes(int64_t rows, int p);
int64_t selb200_nib4_pack_piece(const uint8_t* regs, int64_t rows, int p, uint8_t* piece, int threads);
int selb200_load_device_rows_packed(selb200_ctx* ctx, int64_t g0, int64_t count, const uint8_t* d_pieces, int64_t piece_rows);

/* After a load: cards_sorted[i] = cardinality (double) of the i-th genome in sorted
 * order; order[i] = its index in file-list order.  Either pointer may be NULL. */
int selb200_get_order(selb200_ctx* ctx, double* cards_sorted, int32_t* order);

/* ---- run -------------------------------------------------------------------- */
void selb200_default_params(selb200_params* p);
int selb200_run(selb200_ctx* ctx, const selb200_params* params, selb200_stats* stats);

/* Results of the last run.  i,k are SORTED positions (i < k); map through `order`
 * for file-list indices.  With sort_output they come in the reference's print order
 * (selection.cpp:297-300).  jaccard is the double the reference formats with
 * std::to_string (selection.cpp:288).  The *_near list holds every evaluated pair
 * with |J - tau| <= 1e-6*|tau| (emitted or not) for the tolerance report. */
int64_t selb200_result_count(selb200_ctx* ctx);
int selb200_copy_results(selb200_ctx* ctx, int64_t cap, int32_t* i, int32_t* k, double* jaccard);
int64_t selb200_near_count(selb200_ctx* ctx);
int selb200_copy_near(selb200_ctx* ctx, int64_t cap, int32_t* i, int32_t* k, double* jaccard);
/* Device views of the same lists for NCCL gathers: keys[c] = (uint64)i<<32 | k. */
int selb200_result_device(selb200_ctx* ctx, const uint64_t** d_keys, const double** d_jaccard);
/* Pinned host views of the same lists, valid after a run with params.host_results = 1. */
int selb200_result_host(selb200_ctx* ctx, const uint64_t** keys, const double** jaccard);

/* ---- multi-GPU: gather of the pair lists over peer memory -------------------------------------
 * One process per GPU on one NVLink/NVSwitch node.  The ROOT rank allocates a landing
 * zone for `cap_pairs` pairs on its GPU and exports an opaque handle (CUDA IPC inside); every rank,
 * the root included, attaches with its (rank, world) and that handle — how the bytes travel between
 * the processes is the caller's business (torch.distributed, MPI, a file).  A run with
 * params.gather = 1, shard = rank, n_shards = world then ends with each rank storing its (i,k,J) list
 * straight into the root's memory from a kernel (plain stores over NVLink, one system-scope atomicAdd
 * to claim a block, one to signal); the root waits on the device for all `world` signals and sorts
 * the merged list.  On the root, selb200_result_count / copy_results / copy_near then describe the
 * WHOLE job; on the other ranks they are empty.  All ranks must issue the same sequence of gather
 * runs.  Replaces the cudaMemcpy of `out_count` + `Result[]` of src/selection_cuda.cpp:177-180. */
#define SELB200_GATHER_HANDLE_BYTES 128
int selb200_gather_create(selb200_ctx* ctx, int64_t cap_pairs, void* handle_out);
int selb200_gather_attach(selb200_ctx* ctx, int rank, int world, const void* root_handle);
void selb200_gather_close(selb200_ctx* ctx);

/* ---- host helpers (no device needed) -------------------------------------- */
/* LSH band search.  cpu_variant=1: src/selection.cpp:258-267 (falls through to (m,1));
 * cpu_variant=0: src/selection_cuda.cpp:119-128 (falls through to (1,1)). */
int selb200_band_params(int m, float tau, int cpu_variant, int* n_bands, int* n_rows);
/* std::sort order of the reference (selection.cpp:251-256) for n cardinalities. */
int selb200_sort_order(int64_t n, const double* cards, int32_t* order);

/* ---- diagnostics used by the parity tests --------------------------------- */
/* Per-pair union estimate through the same kernels: for c in [0,count):
 * t[c] = union_size(sketch a[c], sketch b[c]) over the loaded PRIMARY (which=0) or
 * auxiliary-HLL (which=1) sketches; a,b are file-list indices (host arrays). */
int selb200_debug_union(selb200_ctx* ctx, int which, int64_t count, const int32_t* a, const int32_t* b, double* t);

/* ---- sketch builder (what `build_sketch` computes) ---------------------
 * For genome g the sequence characters are seq[offsets[g] .. offsets[g+1]) with ONE non-ACGT byte
 * between FASTA records (a record restarts the rolling 31-mer, src/build_sketch.cpp:61-92).
 * Writes the primary HLL registers (uint8 [n][2^p], sketch/include/sketch/hll.h:886-894 with
 * WangHash, hash.h:44-53) and the auxiliary sketch: SELB200_AUX_HLL -> uint8 [n][2^aux_len];
 * SELB200_AUX_SMH -> uint64 [n][selb200_smh_size(aux_len)] SuperMinHash buckets
 * (sketch/include/sketch/bbmh.h:639-670).  Host pointers in and out; byte-identical to the
 * reference's sketches.  selb200_sketch_last_error() reports this family's failures. */
int selb200_smh_size(int m_arg);   /* bucket count after SizePow2Policy rounding (policy.h:14-19) */
int selb200_sketch_host(int device, int64_t n_genomes, const uint8_t* seq, const int64_t* offsets, int p,
                        int aux_kind, int aux_len, uint8_t* out_hll, void* out_aux);
const char* selb200_sketch_last_error(void);
/* Create the CUDA context of `device` ahead of time (call it on a side thread while reading files). */
int selb200_warmup(int device);

/* ---- synthetic sketches ("synth-v1", integer-exact; bench + tests) --------
 * Registers follow P(reg <= k) = T[k]/2^64 with T supplied by the caller
 * (64 thresholds per stream); a member's registers are byte-max(core, private),
 * its SuperMinHash buckets element-wise min(core, private) of uniform draws below
 * R.  on_device=1 writes to a device pointer with a kernel; on_device=0 fills host
 * memory with the same integer arithmetic, so both agree bit for bit. */
int selb200_synth_hll(int on_device, int device, int64_t n, int p, const int32_t* cluster, int64_t n_clusters,
                      const uint64_t* thr_core, const uint64_t* thr_priv, uint64_t seed, uint32_t tag,
                      uint8_t* out);
int selb200_synth_smh(int on_device, int device, int64_t n, int m, const int32_t* cluster, int64_t n_clusters,
                      const uint64_t* range_core, const uint64_t* range_priv, uint64_t seed, uint32_t tag,
                      uint64_t* out);

#ifdef __cplusplus
}
#endif
#endif /* SELB200_H */
