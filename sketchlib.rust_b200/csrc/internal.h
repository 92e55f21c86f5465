// Internals of libsketchdist shared by capi.cu (single-device entry points) and peer.cu (peer groups of several
// devices): context / database objects and the host-side drivers the multi-device operations are built from.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "../../include/sketchdist.h"
#include "kernels.h"
#include "layout.h"

namespace skdi {
constexpr size_t kMaxSmem = 232448;           // 227 KB opt-in dynamic shared memory per CTA on sm_100
constexpr uint32_t kSymKnnMax = 64;           // symmetric self-kNN keeps 128-slot lists: knn + 64 candidates per step
}  // namespace skdi

struct skd_ctx {
    int device = 0;
    cudaStream_t compute = nullptr, copy = nullptr;
    cudaStream_t aux[2] = {nullptr, nullptr};  // async-upload dense runs: re-tiling / every other column band
    std::string err;
    uint64_t launches = 0;
    float last_ms = 0.f;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timers;  // event pairs around kernel launches of a call
    std::vector<int> timer_tags;  // 1 = a tile_kernel / stream_kernel launch (the path's dominant kernel)
    size_t timers_used = 0;
    float last_tile_ms = 0.f;
    cudaEvent_t ev_a[2] = {nullptr, nullptr}, ev_b[2] = {nullptr, nullptr};
    std::map<std::string, std::pair<void *, size_t>> scratch;  // named device buffers reused across calls
    void *pinned[2] = {nullptr, nullptr};
    size_t pinned_bytes = 0;
    int sm_count = 148;
    bool knn_symmetric = true;  // SKD_KNN_SYMMETRIC=0: whole-matrix self kNN evaluates both (i,j) and (j,i)
    bool dense_v2 = true;    // dense modes run the persistent barrier-free dense_kernel (SKD_DENSE_V1=1: tile_kernel)
    bool zero_copy = true;   // dense kernels store straight into pinned host results (SKD_ZEROCOPY=0 disables)
    bool early_out = true;   // core/accessory dense_kernel: warps stop counting after the third k-mer length when (almost)
                             // all their pairs are final, the rest goes to defer_kernel (SKD_EARLY_OUT=0 disables)
    uint32_t defer_max = 32; // SKD_DEFER_MAX: open pairs of a 16 x 64 warp block that may be deferred
    uint32_t defer_cap = 8u << 20;  // SKD_DEFER_CAP: entries (32 B) of the deferred-pair list of one launch; warps that
                                    // find it full keep counting
    unsigned long long *work_stats = nullptr;  // device: [1] warp x word steps in all, [3] steps skipped, [2] deferred pairs
    // Tunables, read once from the environment when the context is created.
    size_t upload_chunk_bytes = 64ull << 20;  // SKD_UPLOAD_CHUNK_MB: staging chunk of a pageable/file upload
    size_t result_chunk_bytes = 64ull << 20;  // SKD_RESULT_CHUNK_MB: staged D2H chunk of dense results
    uint32_t group = 24;                      // SKD_GROUP: row blocks per L2 group of the tile scheduler (B200: DRAM reads of a
                                              // dense 10k run 1.52 GB at 8, 0.75 GB at 24, 1.07 GB at 48, 1.61 GB at 64)
    size_t sym_chunk_bytes = 8ull << 30;      // SKD_SYM_CHUNK_MB: HBM key buffer of one symmetric-kNN row group (100k kNN
                                              // step on B200: 547 ms at 2 GB, 519 ms at 8 GB, 513 ms at 16 GB: fewer, longer
                                              // column folds in select_sym_kernel)
    uint64_t few_queries = 1ull << 20;        // SKD_FEW_QUERIES: cross kNN up to this many queries goes dense keys + select
    size_t read_threads = 8;                  // SKD_READ_THREADS: host threads copying an .skd chunk into pinned staging
                                              // (default: the host's hardware threads, at most 16)
};

struct skd_db {
    skd_ctx *ctx = nullptr;
    skd::DevDb dev{};
    uint32_t *tiled = nullptr;
    size_t tiled_bytes = 0;
    std::vector<uint64_t> kmers;
    uint64_t sketch_size = 0, W = 0, n = 0;
    double *lny_dev = nullptr;  // [nbins + 1]
    uint64_t *rows = nullptr;   // sample-major copy of the planes, made on demand for the early exit's pair kernel
    // skd_db_upload_async: the raw planes are still arriving (copy stream, one event per chunk of whole sample
    // blocks); they are re-tiled by the first compute call -- band by band, interleaved with the tile kernel, when
    // that call is a dense self run, so that the PCIe transfer overlaps the distance computation.
    struct Pending {
        uint64_t *raw = nullptr;                 // device copy in the reference layout
        uint64_t sample_stride = 0, kmer_stride = 0;
        std::vector<cudaEvent_t> ready;          // chunk c has arrived
        std::vector<uint32_t> block_end;         // chunk c covers sample blocks [block_end[c-1], block_end[c])
    } pending;
    bool is_pending() const { return pending.raw != nullptr; }
};

namespace skdi {

skd_status fail(skd_ctx *ctx, skd_status st, const char *fmt, ...);

#define CU(ctx, expr)                                                                                   \
    do {                                                                                                \
        cudaError_t e__ = (expr);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return skdi::fail(ctx, SKD_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, \
                              __LINE__);                                                                \
    } while (0)

cudaError_t preload_peer_kernels();  // peer.cu
skd_status scratch_get(skd_ctx *ctx, const char *name, size_t bytes, void **out);
void timers_reset(skd_ctx *ctx);
skd_status timer_begin(skd_ctx *ctx, cudaStream_t stream = nullptr, int tag = 0);
skd_status timer_end(skd_ctx *ctx, cudaStream_t stream = nullptr);
skd_status timers_collect(skd_ctx *ctx);  // call after the compute stream is idle

skd_status build_epi(skd_ctx *ctx, const skd_db *A, const skd_db *B, skd_dist_type dist, const double *a_comp,
                     const double *b_comp, double cutoff, skd::EpiParams *E);
skd_status check_pair(skd_ctx *ctx, const skd_db *A, const skd_db *B, bool keep_pending = false);

// Row-owned kNN (capi.cu): final sparse rows, or (key_out set) one sorted composite list per row.
skd_status run_knn(skd_ctx *ctx, const skd_db *R, const skd_db *Q, bool self, const skd::EpiParams &E, uint64_t knn,
                   uint64_t row_begin, uint64_t row_end, uint64_t *idx_out, float *d0_out, float *d1_out, skd_loc out_loc,
                   uint64_t *key_out = nullptr, float *aux_out = nullptr, uint64_t idx_offset = 0);
// The same contract for the corners of the envelope (knn beyond the per-row list capacity, k-mer counts beyond shared
// memory): dense keys + a full sort of every row (sortknn.cu).
skd_status run_knn_sorted(skd_ctx *ctx, const skd_db *R, const skd_db *Q, bool self, const skd::EpiParams &E, uint64_t knn,
                          uint64_t row_begin, uint64_t row_end, uint64_t *idx_out, float *d0_out, float *d1_out,
                          skd_loc out_loc, uint64_t *key_out, float *aux_out, uint64_t idx_offset);
// Dense launches (capi.cu): tile-group offsets, kernel choice, one launch over rows [r0, r1) into device memory.
void build_groups(uint32_t group, bool self, uint32_t rb0, uint32_t n_rb, uint32_t nbB, std::vector<uint64_t> *off,
                  bool pairs, uint32_t cb_lo = 0, uint32_t cb_hi = 0);
bool uses_dense_kernel(const skd_ctx *ctx, int mode, const skd_db *B, const skd::EpiParams &E);
skd_status launch_dense(skd_ctx *ctx, const skd_db *A, const skd_db *B, const skd::EpiParams &E, int mode, uint64_t r0,
                        uint64_t r1, float *dout, const std::vector<uint64_t> &off, const uint64_t *d_grp_off,
                        uint64_t rect_ld = 0, uint32_t cb_lo = 0, uint32_t cb_hi = 0, uint64_t out_row_begin = ~0ull,
                        cudaStream_t stream = nullptr);
// Final sparse rows from n_parts device-resident partial lists (keys[part][n_rows][knn]).
skd_status finalize_lists(skd_ctx *ctx, const unsigned long long *dkeys, const float *daux, uint32_t n_parts,
                          uint64_t part_stride, uint64_t n_rows, uint64_t knn, int kind, uint64_t *idx_out, float *d0_out,
                          float *d1_out, skd_loc out_loc);
// Symmetric self-kNN over triangle rows [row_begin, row_end): lists dkey/daux [n][knn] in device memory.
skd_status run_knn_symmetric(skd_ctx *ctx, const skd_db *db, const skd::EpiParams &E, uint64_t knn, uint64_t row_begin,
                             uint64_t row_end, unsigned long long *dkey, float *daux);

// Source of a database build: an array in the reference layout (host or device), an .skd file, or nothing
// (allocate only: the tiled planes are filled by the caller, peer.cu gathers them from the peers' slices).
struct ChunkSource {
    const uint64_t *direct = nullptr;  // array in the reference layout (host or device)
    skd_loc loc = SKD_HOST;
    int fd = -1;                       // .skd file
    const uint64_t *indices = nullptr; // file positions, nullable
    bool async = false;                // pinned host array: return once the copies are queued (skd_db_upload_async)
    bool alloc_only = false;
};
skd_status db_build(skd_ctx *ctx, const ChunkSource &src, uint64_t n, uint64_t sample_stride, uint64_t kmer_stride,
                    uint64_t W, uint64_t sketch_size, const uint64_t *kmers, uint32_t K, skd_db **out);

}  // namespace skdi
