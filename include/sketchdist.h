/*
 * sketchdist.h -- C ABI of libsketchdist.so, the B200 (sm_100a) implementation of the pairwise
 * sketch-distance path of bacpop/sketchlib.rust.
 *
 * The reference has no FFI seam for this path; its seam is the public Rust API of
 * `sketchlib::distances` as called from `main` (reference src/lib.rs:341-451).  Every entry point
 * below names the reference function it replaces (file:line relative to the reference checkout).
 * A Rust `-sys` crate binding these is sketched in INTEGRATION.md; in this repository the host
 * layer above the ABI is C++ (sketchlib.rust_b200/csrc/host) and Python ctypes
 * (sketchlib.rust_b200/__init__.py).
 *
 * Conventions
 *   - every function returns an skd_status (0 = ok); no exception or unwind crosses the boundary;
 *     where the reference panics, the matching status is returned and skd_last_error() holds the
 *     reference's message;
 *   - all pointers are caller-owned; sizes are in elements unless the name says bytes;
 *   - "loc" arguments say where a caller buffer lives: SKD_HOST (pageable or pinned host memory)
 *     or SKD_DEVICE (memory of the context's GPU);
 *   - one context per GPU per process (one process per GPU for multi-GPU runs; rows are sharded by
 *     the row/ref/query ranges every compute call takes); a context is not thread-safe, but a thread may
 *     create a second context on the same GPU (a loader filling the next shard: skd_db_adopt).
 *   - there is no CPU fallback: without a CUDA device skd_ctx_create fails.
 */
#ifndef SKETCHDIST_H
#define SKETCHDIST_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum skd_status {
    SKD_OK = 0,
    SKD_ERR_INVALID = 1,      /* bad argument */
    SKD_ERR_CUDA = 2,         /* CUDA runtime error, text in skd_last_error */
    SKD_ERR_NEED_TWO_K = 3,   /* reference panic: src/distances/jaccard.rs:70-72 */
    SKD_ERR_K_NOT_FOUND = 4,  /* reference Err: src/distances/mod.rs:28-30 */
    SKD_ERR_EMPTY_DB = 5,     /* reference panic: src/distances/mod.rs:318-323 */
    SKD_ERR_INCOMPATIBLE = 6, /* ref/query sketch size or k-mer list differ (multisketch.rs:222-226) */
    SKD_ERR_UNSUPPORTED = 7,  /* outside the implemented envelope */
    SKD_ERR_IO = 8,           /* file could not be read */
    SKD_ERR_NOMEM = 9
} skd_status;

typedef enum skd_loc { SKD_HOST = 0, SKD_DEVICE = 1 } skd_loc;

/* DistType: src/distances/distance_matrix.rs:55-60 */
typedef enum skd_dist_kind { SKD_CORE_ACC = 0, SKD_JACCARD = 1, SKD_ANI = 2 } skd_dist_kind;
typedef struct skd_dist_type {
    int32_t kind;  /* skd_dist_kind */
    int32_t k_idx; /* index into the k-mer list for SKD_JACCARD / SKD_ANI; ignored for SKD_CORE_ACC */
} skd_dist_type;

typedef struct skd_ctx skd_ctx;
typedef struct skd_db skd_db; /* a MultiSketch resident in HBM (src/sketch/multisketch.rs:23-43) */

/* ---- context ---------------------------------------------------------------------------- */
/* Number of CUDA devices visible to the process (0 without a driver or GPU). */
int skd_device_count(void);
/* One context per GPU and host thread/process.  With SKD_BIND_CPU=1 in the environment the calling thread is first
 * bound to the CPUs next to the GPU (NVML's ideal affinity), so that the pinned host buffers of skd_host_alloc and the
 * thread filling them sit on the GPU's NUMA node; off by default because it changes the caller's affinity. */
skd_status skd_ctx_create(int device, skd_ctx **out);
void skd_ctx_destroy(skd_ctx *ctx);
/* Message of the last failing call on this context (or of a failed skd_ctx_create when ctx == NULL). */
const char *skd_last_error(const skd_ctx *ctx);
int skd_ctx_device(const skd_ctx *ctx);
/* Kernels launched by this context since creation (bench.py's gpu_launches). */
uint64_t skd_ctx_launch_count(const skd_ctx *ctx);
/* Work actually done by the dense core/accessory kernel since the context was created (or the last reset): like the
 * reference's loop (src/distances/jaccard.rs:89-91: `break` at the first k-mer length without a match) the kernel stops
 * counting a 16 x 64 block of pairs after the third k-mer length when (almost) every pair of it is already final.
 * steps_computed / steps_total = the share of the (warp, 64-bin word) match-count steps that were executed;
 * pairs_deferred = pairs finished one by one afterwards.  SKD_EARLY_OUT=0 in the environment disables the early exit
 * (steps_computed == steps_total).  Measurement aid; synchronises the context's streams. */
/* Turns that early exit off (0) or back on (non-zero; the default unless SKD_EARLY_OUT=0): with it off the kernel
 * counts every k-mer length of every pair, which is the work the roofline figures are quoted on.  Results are the
 * same either way. */
skd_status skd_ctx_set_early_exit(skd_ctx *ctx, int enabled);
skd_status skd_ctx_work_counters(skd_ctx *ctx, uint64_t *steps_computed, uint64_t *steps_total, uint64_t *pairs_deferred,
                                 int reset);
/* Device time in ms of the distance kernels of the last compute call, measured with CUDA events on
 * the context's compute stream (the stream the kernels are launched on). */
float skd_ctx_last_kernel_ms(const skd_ctx *ctx);
/* The part of it spent in the match-count kernels (tile_kernel / stream_kernel launches), the path's dominant
 * kernels: what bench.py's roofline is computed from. */
float skd_ctx_last_tile_kernel_ms(const skd_ctx *ctx);
/* Block until all work queued by this context is complete. */
skd_status skd_ctx_synchronize(skd_ctx *ctx);
/* The context's compute stream as a cudaStream_t (for callers that time with their own events). */
void *skd_ctx_stream(const skd_ctx *ctx);

/* Pinned host memory for inputs/results (plays the part of the Vec<f32> the reference's
 * DistanceMatrix::new allocates, src/distances/distance_matrix.rs:148-152). */
skd_status skd_host_alloc(skd_ctx *ctx, size_t bytes, void **out);
void skd_host_free(skd_ctx *ctx, void *ptr);
/* Device memory helpers for callers without a CUDA runtime binding. */
skd_status skd_device_alloc(skd_ctx *ctx, size_t bytes, void **out);
void skd_device_free(skd_ctx *ctx, void *ptr);
skd_status skd_memcpy(skd_ctx *ctx, void *dst, skd_loc dst_loc, const void *src, skd_loc src_loc, size_t bytes);

/* ---- sketch database loader ---------------------------------------------------------------
 * Replaces MultiSketch::read_sketch_data / get_sketch_slice (src/sketch/multisketch.rs:167-184,
 * 213-219).  `bins` is the reference's flat u64 array: word
 * [s*sample_stride + k_idx*kmer_stride + w*14 + p] = bit plane p of bins 64w..64w+63 of sample s at
 * k-mer index k_idx.  The library keeps its own re-tiled copy in HBM (k-major, 64-sample blocks,
 * plane pairs interleaved); the caller's buffer is not referenced afterwards. */
skd_status skd_db_upload(skd_ctx *ctx, const uint64_t *bins, skd_loc bins_loc, uint64_t n_samples,
                         uint64_t sample_stride, uint64_t kmer_stride, uint64_t sketchsize64,
                         uint64_t sketch_size, const uint64_t *kmer_lengths, uint32_t n_kmers, skd_db **out);
/* Same as skd_db_upload for a PINNED host array (skd_host_alloc), but returns as soon as the host-to-device
 * copies are queued: the caller must keep `bins` unchanged until the first compute call on the database (or
 * skd_ctx_synchronize) has returned.  The planes are re-tiled chunk by chunk by that first call; a dense
 * skd_self_dists_all interleaves the re-tiling with the distance tiles of the chunks that have arrived, so the
 * PCIe transfer overlaps the computation.  Falls back to the synchronous path for pageable memory. */
skd_status skd_db_upload_async(skd_ctx *ctx, const uint64_t *bins, uint64_t n_samples, uint64_t sample_stride,
                               uint64_t kmer_stride, uint64_t sketchsize64, uint64_t sketch_size,
                               const uint64_t *kmer_lengths, uint32_t n_kmers, skd_db **out);
/* Replaces read_sketch_data / read_sketch_data_block + SketchArrayReader (multisketch.rs:167-210,
 * src/sketch/sketch_datafile.rs:159-194): streams an .skd file into HBM through pinned staging.
 * `sample_indices` (nullable) lists the FILE positions (Sketch::index) of the samples to load, in the
 * order they are to appear (the --subset order, block_reindex); NULL loads samples 0..n_samples-1. */
skd_status skd_db_load_skd(skd_ctx *ctx, const char *skd_path, const uint64_t *sample_indices,
                           uint64_t n_samples, uint64_t sample_stride, uint64_t kmer_stride,
                           uint64_t sketchsize64, uint64_t sketch_size, const uint64_t *kmer_lengths,
                           uint32_t n_kmers, skd_db **out);
void skd_db_free(skd_db *db);
/* Hands a completely loaded database to another context of the SAME device.  A loader thread with its own context
 * can then fill the next shard of a large reference file (skd_db_load_skd is synchronous) while this context
 * compares the previous one: the out-of-core / overlapped form of read_sketch_data + cross_dists_knn
 * (multisketch.rs:167-184, src/distances/mod.rs:306-395).  The previous owner must not use `db` afterwards. */
skd_status skd_db_adopt(skd_ctx *ctx, skd_db *db);
uint64_t skd_db_n_samples(const skd_db *db);       /* number_samples_loaded, multisketch.rs:106-111 */
uint32_t skd_db_n_kmers(const skd_db *db);
uint64_t skd_db_sketch_size(const skd_db *db);     /* bins */
uint64_t skd_db_hbm_bytes(const skd_db *db);
/* is_compatible_with (multisketch.rs:222-226) minus the hash type, which the ABI does not see. */
int skd_db_compatible(const skd_db *a, const skd_db *b);

/* set_k: src/distances/mod.rs:25-37.  kmer < 0 means "no -k given" => core/accessory. */
skd_status skd_set_k(skd_ctx *ctx, const skd_db *db, int64_t kmer, int ani, skd_dist_type *out);

/* ---- distances --------------------------------------------------------------------------------
 * completeness vectors are host arrays of n_samples doubles or NULL (Option<&Vec<f64>>).        */

/* self_dists_all: src/distances/mod.rs:58-130, restricted to rows [row_begin, row_end) of the upper
 * triangle.  `out` receives, in condensed row-major order, the distances of all pairs (i, j) with
 * row_begin <= i < row_end, i < j < n: element 0 is pair (row_begin, row_begin+1).  1 float per pair
 * (Jaccard distance 1-J, or ANI) or 2 interleaved (core, accessory).  row range (0, n) is the whole
 * matrix of the reference. */
skd_status skd_self_dists_all(skd_ctx *ctx, const skd_db *db, skd_dist_type dist, const double *completeness,
                              double completeness_cutoff, uint64_t row_begin, uint64_t row_end, float *out,
                              skd_loc out_loc);
/* Number of pairs skd_self_dists_all writes for a row range (square_to_condensed arithmetic,
 * src/distances/distance_matrix.rs:11-14). */
uint64_t skd_self_pairs_in_rows(uint64_t n, uint64_t row_begin, uint64_t row_end);

/* cross_dists_all: src/distances/mod.rs:227-297, refs [ref_begin, ref_end).  Ref-major:
 * out[((i - ref_begin) * n_query + j) * ncols]. */
skd_status skd_cross_dists_all(skd_ctx *ctx, const skd_db *ref, const skd_db *query, skd_dist_type dist,
                               const double *ref_completeness, const double *query_completeness,
                               double completeness_cutoff, uint64_t ref_begin, uint64_t ref_end, float *out,
                               skd_loc out_loc);

/* self_dists_knn: src/distances/mod.rs:133-224, rows [row_begin, row_end).  Outputs are row-major
 * (row_end-row_begin) x knn: neighbour index, first distance (Jaccard dist / ANI / core), second
 * distance (accessory; may be NULL unless core/acc).  Rows ascend by the reference's key (1-J, 1-ANI,
 * core); ties are broken by neighbour index.  Requires 1 <= knn <= n-1 (the CLI clamps,
 * src/lib.rs:379-382). */
skd_status skd_self_dists_knn(skd_ctx *ctx, const skd_db *db, uint64_t knn, skd_dist_type dist,
                              const double *completeness, double completeness_cutoff, uint64_t row_begin,
                              uint64_t row_end, uint64_t *idx_out, float *d0_out, float *d1_out, skd_loc out_loc);

/* cross_dists_knn: src/distances/mod.rs:306-395, one row per query in [query_begin, query_end), each
 * scanning all refs; knn is clamped to n_ref like the reference (:325) and returned in *knn_used, which
 * is also the row pitch of the outputs. */
skd_status skd_cross_dists_knn(skd_ctx *ctx, const skd_db *ref, const skd_db *query, uint64_t knn,
                               skd_dist_type dist, const double *ref_completeness,
                               const double *query_completeness, double completeness_cutoff,
                               uint64_t query_begin, uint64_t query_end, uint64_t *idx_out, float *d0_out,
                               float *d1_out, skd_loc out_loc, uint64_t *knn_used);

/* ---- partial kNN lists for multi-GPU runs that split the COLUMNS of the problem -----------------------------
 * A partial list is, per row, `knn` opaque 64-bit composites (distance key and neighbour index) in ascending
 * order, padded with UINT64_MAX, plus (core/accessory only) the accessory distance of each entry.  Ranks exchange
 * them (rows x knn x 12 bytes per rank: through the peer groups below, or any transport of the host's) and
 * skd_knn_merge turns `n_parts` of them into the reference's sparse rows.  Each candidate must appear in exactly one
 * partial.  A list may hold fewer than knn real entries even when more candidates were seen: candidates at distance 1
 * whose neighbour index exceeds knn can never reach a final row (every sample with a smaller index precedes them)
 * and are dropped at the source.
 *
 * skd_self_knn_partial: symmetric self mode -- evaluates every unordered pair (i, j > i) with i in
 *   [row_begin, row_end) ONCE and credits it to both rows, so the lists cover ALL n rows (n x knn entries); the
 *   union over a partition of the rows (row_begin a multiple of 64; use equal-pair triangle blocks) is the
 *   result of self_dists_knn (src/distances/mod.rs:133-224) at half its pair evaluations.  For knn > 64 the rows
 *   [row_begin, row_end) are evaluated against all columns instead and only they get lists (same merged result).
 * skd_cross_knn_partial: cross mode with the REFERENCES sharded -- the best knn of this shard for every query
 *   (n_query x knn entries); neighbour indices are shifted by ref_index_offset (the shard's first reference). */
skd_status skd_self_knn_partial(skd_ctx *ctx, const skd_db *db, uint64_t knn, skd_dist_type dist,
                                const double *completeness, double completeness_cutoff, uint64_t row_begin,
                                uint64_t row_end, uint64_t *keys_out, float *aux_out, skd_loc out_loc);
skd_status skd_cross_knn_partial(skd_ctx *ctx, const skd_db *ref_shard, const skd_db *query, uint64_t knn,
                                 skd_dist_type dist, const double *ref_completeness, const double *query_completeness,
                                 double completeness_cutoff, uint64_t ref_index_offset, uint64_t *keys_out,
                                 float *aux_out, skd_loc out_loc);
/* keys: [n_parts][n_rows][knn] (aux likewise; NULL unless core/accessory).  Outputs as skd_self_dists_knn. */
skd_status skd_knn_merge(skd_ctx *ctx, const uint64_t *keys, const float *aux, uint32_t n_parts, uint64_t n_rows,
                         uint64_t knn, skd_dist_type dist, skd_loc in_loc, uint64_t *idx_out, float *d0_out,
                         float *d1_out, skd_loc out_loc);

/* ---- peer groups: the GPUs of one box working on ONE distance problem -------------------------------------------
 * The reference has a single shared-memory thread pool (rayon, src/cli.rs:75-86); its multi-GPU counterpart is a
 * group of `world` contexts, one per GPU -- one process per GPU (handles exchanged by the launcher, memory mapped
 * with CUDA IPC) or several host threads of one process (direct peer access).  Every rank owns a window of device
 * memory that its peers can read over NVLink; the library's own kernels do the exchange (the merge kernel pulls the
 * partial kNN lists it needs straight from the peers' windows, the loader pulls the peers' re-tiled sample slices),
 * synchronised by flag barriers in peer memory.  There is no NCCL or MPI dependency.
 *
 * All skd_comm_* compute calls are COLLECTIVE: every rank of the group calls them in the same order with the same
 * arguments (except where noted); a rank that never arrives makes the others fail with SKD_ERR_CUDA after the
 * barrier timeout (SKD_PEER_TIMEOUT_S, default 60 s) instead of hanging. */
typedef struct skd_comm skd_comm;
#define SKD_COMM_HANDLE_BYTES 128
#define SKD_COMM_MAX_WORLD 16

skd_status skd_comm_create(skd_ctx *ctx, int rank, int world, size_t window_bytes, skd_comm **out);
void skd_comm_destroy(skd_comm *comm);
/* handle_out: SKD_COMM_HANDLE_BYTES bytes describing this rank's window, to be sent to every peer. */
skd_status skd_comm_export(const skd_comm *comm, void *handle_out);
/* handles: world x SKD_COMM_HANDLE_BYTES bytes in rank order (the own entry is ignored). */
skd_status skd_comm_connect(skd_comm *comm, const void *handles);
int skd_comm_rank(const skd_comm *comm);
int skd_comm_world(const skd_comm *comm);
/* Device time in ms of the last collective call on this rank (CUDA events on the context's compute stream, from
 * the entry barrier to the last kernel or copy of the call). */
float skd_comm_last_op_ms(const skd_comm *comm);
/* The part of it spent after the local distance kernels: barriers + peer merge + gather of the final rows. */
float skd_comm_last_exchange_ms(const skd_comm *comm);
/* The triangle rows this rank evaluated in its last skd_comm_self_dists_knn call (knn <= 64).  The first call of a
 * problem shape splits the triangle into blocks of equal pair counts (skd_triangle_rows); later calls of the same
 * shape re-split it from the ranks' measured times so that they finish together (SKD_ADAPTIVE_SPLIT=0: never).  The
 * results do not depend on the split. */
void skd_comm_last_row_block(const skd_comm *comm, uint64_t *row_begin, uint64_t *row_end);
/* Window sizes the collective calls need (the window must hold the largest one used). */
size_t skd_comm_window_bytes_knn(uint64_t n_rows, uint64_t knn, int world);
size_t skd_comm_window_bytes_db(uint64_t n_samples, uint64_t sketchsize64, uint32_t n_kmers, int world);

/* The partitions the collective calls use (exported so that every host layer derives them from one rule):
 * skd_triangle_rows: upper-triangle row block of `rank` with about 1/world of the n(n-1)/2 pairs (boundaries at
 *   n(1 - sqrt(1 - r/world)), rounded to the 64-sample tile) -- dense self outputs and symmetric self kNN;
 * skd_slice_rows: contiguous slices of ceil(n/world) items rounded up to `align` -- sample slices of the
 *   distributed loader (align 64), reference shards (align 64), owned result rows (align 1). */
void skd_triangle_rows(uint64_t n, int rank, int world, uint64_t *row_begin, uint64_t *row_end);
void skd_slice_rows(uint64_t n, int rank, int world, uint64_t align, uint64_t *begin, uint64_t *end);
/* The triangle split for given pair shares (shares[world], summing to 1; NULL = equal shares = skd_triangle_rows):
 * bounds[world + 1], bounds[0] = 0, bounds[world] = n, rank r takes rows [bounds[r], bounds[r + 1]).  What
 * skd_comm_self_dists_knn uses when it re-splits the triangle from the ranks' measured times. */
void skd_triangle_rows_weighted(uint64_t n, int world, const double *shares, uint64_t *bounds);

/* Distributed form of read_sketch_data (multisketch.rs:167-184): every rank ends up with the WHOLE database in its
 * HBM, but moves only its own slice skd_slice_rows(n_total, rank, world, 64) over PCIe: `bins_slice` points at the
 * first sample of that slice (reference layout, host or device memory).  The slice is re-tiled locally into the
 * window and every rank pulls the other slices from its peers over NVLink. */
skd_status skd_comm_db_upload(skd_comm *comm, const uint64_t *bins_slice, skd_loc bins_loc, uint64_t n_total,
                              uint64_t sample_stride, uint64_t kmer_stride, uint64_t sketchsize64,
                              uint64_t sketch_size, const uint64_t *kmer_lengths, uint32_t n_kmers, skd_db **out);
/* Same from an .skd file: `sample_indices` (nullable) lists the file positions of all n_total samples in load order
 * (the --subset order); every rank reads only the entries of its slice. */
skd_status skd_comm_db_load_skd(skd_comm *comm, const char *skd_path, const uint64_t *sample_indices,
                                uint64_t n_total, uint64_t sample_stride, uint64_t kmer_stride,
                                uint64_t sketchsize64, uint64_t sketch_size, const uint64_t *kmer_lengths,
                                uint32_t n_kmers, skd_db **out);

/* self_dists_knn (src/distances/mod.rs:133-224) over the group, database replicated on every rank.  knn <= 64:
 * every rank evaluates its skd_triangle_rows block once per unordered pair (skd_self_knn_partial), the owner of a
 * row (skd_slice_rows(n, rank, world, 1)) merges the partial lists of that row straight out of its peers' windows;
 * larger knn: every rank computes its owned rows against all columns.  The final n x knn rows are gathered on rank
 * `root` (every rank if root < 0) into idx_out / d0_out / d1_out (ignored on the other ranks, may be NULL there). */
skd_status skd_comm_self_dists_knn(skd_comm *comm, const skd_db *db, uint64_t knn, skd_dist_type dist,
                                   const double *completeness, double completeness_cutoff, int root,
                                   uint64_t *idx_out, float *d0_out, float *d1_out, skd_loc out_loc);
/* cross_dists_knn (src/distances/mod.rs:306-395) with the REFERENCES sharded: rank r holds the references
 * [ref_index_offset, ref_index_offset + n(ref_shard)) of n_ref_total (`ref_completeness` covers the shard) and all
 * queries; per-shard partial lists are merged by the owners of the query rows and gathered on `root`.  knn is
 * clamped to n_ref_total (:325), *knn_used is the row pitch of the outputs. */
skd_status skd_comm_cross_dists_knn(skd_comm *comm, const skd_db *ref_shard, const skd_db *query, uint64_t knn,
                                    skd_dist_type dist, const double *ref_completeness,
                                    const double *query_completeness, double completeness_cutoff,
                                    uint64_t ref_index_offset, uint64_t n_ref_total, int root, uint64_t *idx_out,
                                    float *d0_out, float *d1_out, skd_loc out_loc, uint64_t *knn_used);
/* The exchange half alone, for callers that produce the partial lists themselves (skd_self_knn_partial /
 * skd_cross_knn_partial; e.g. one list set per reference shard streamed through HBM): `keys` holds n_parts_local
 * list sets [n_parts_local][n_rows][knn] of this rank (aux likewise, NULL unless core/accessory); they are folded
 * into one, merged across the group by the rows' owners out of peer memory and gathered on `root`. */
skd_status skd_comm_knn_merge(skd_comm *comm, const uint64_t *keys, const float *aux, uint32_t n_parts_local,
                              skd_loc in_loc, uint64_t n_rows, uint64_t knn, skd_dist_type dist, int root,
                              uint64_t *idx_out, float *d0_out, float *d1_out, skd_loc out_loc);
/* Flag barrier over the peers' windows on the compute stream, then a stream synchronise (a collective no-op). */
skd_status skd_comm_barrier(skd_comm *comm);

/* The distance half of self_dists_knn_precluster (src/distances/mod.rs:399-553): per row i the candidates
 * cand_idx[cand_offsets[i] .. cand_offsets[i+1]) (.skd sample indices; what Inverted::any_shared_bins returned,
 * src/inverted.rs:259-268; the row itself may be listed and is skipped) are compared with row i and the knn
 * nearest kept, ascending; short rows are padded with (i, 1.0) (:537-544).  retain_unmatched: 0 = none,
 * 1 = singleton (a row without any neighbour becomes (i, 0.0), (i, 1.0), ...; :491-499), 2 = brute force (the
 * caller lists all samples for such rows).  Single-k distance types only (the reference's CoreAcc arm is
 * unimplemented!, :548-550).  All buffers are host memory; outputs are n x knn. */
skd_status skd_self_dists_knn_candidates(skd_ctx *ctx, const skd_db *db, uint64_t knn, skd_dist_type dist,
                                         const double *completeness, double completeness_cutoff,
                                         const uint64_t *cand_offsets, const uint64_t *cand_idx, int retain_unmatched,
                                         uint64_t *idx_out, float *d0_out);

/* Distances for an explicit list of (ref_idx, query_idx) pairs (the per-pair calls inside
 * self_dists_knn_precluster, src/distances/mod.rs:399-553).  out: 1 or 2 floats per pair. */
skd_status skd_pair_dists(skd_ctx *ctx, const skd_db *ref, const skd_db *query, skd_dist_type dist,
                          const double *ref_completeness, const double *query_completeness,
                          double completeness_cutoff, const uint64_t *pairs, uint64_t n_pairs, float *out);

/* Parity/debug: raw per-k bin-match counts (`samebits`, src/distances/jaccard.rs:15-25) for explicit
 * pairs, counts[p * n_kmers + k]; and for a dense tile range exactly as the tiled kernel computes them:
 * counts[((i - ref_begin) * n_query + j) * n_kmers + k]. */
skd_status skd_match_counts(skd_ctx *ctx, const skd_db *ref, const skd_db *query, const uint64_t *pairs,
                            uint64_t n_pairs, uint32_t *counts);
skd_status skd_match_counts_dense(skd_ctx *ctx, const skd_db *ref, const skd_db *query, uint64_t ref_begin,
                                  uint64_t ref_end, uint32_t *counts);

#ifdef __cplusplus
}
#endif
#endif /* SKETCHDIST_H */
