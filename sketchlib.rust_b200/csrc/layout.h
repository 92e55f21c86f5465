// Shared host/device definitions: the HBM layout of a sketch database and the kernel parameter blocks.
#pragma once
#include <cstdint>

namespace skd {

constexpr int BBITS = 14;                  // bit planes per bin (reference src/sketch/mod.rs:34)
constexpr int PP = BBITS / 2;              // plane pairs
constexpr int TS = 64;                     // samples per block = tile edge
constexpr int HALF_U32 = PP * TS * 2;      // u32 per (32-bin word column, sample block) = 896 (3584 B)
constexpr int STAGE_U32 = 2 * HALF_U32;    // u32 per (64-bin word, sample block): one pipeline stage operand
constexpr int STAGE_BYTES = STAGE_U32 * 4; // 7168 B
constexpr int NST = 3;                     // pipeline stages
#ifndef SKD_TJ
#define SKD_TJ 4
#endif
constexpr int TJ = SKD_TJ;                 // columns per thread: 4 (8 warps, warp tile 16x32) or 8 (4 warps, 16x64)
constexpr int THREADS = TJ == 4 ? 256 : 128;  // thread tile 4 x TJ pairs
constexpr int MAXK = 64;                   // k-mer lengths supported by the fused core/accessory epilogue

// Tiled layout ("k-major, 64-bin-word, sample-block"): one (k, w, b) slab = 64 samples x 14 planes x 64 bins =
// STAGE_U32 u32 = 7168 contiguous bytes, exactly what one bulk (TMA) copy moves into shared memory.  Inside a
// slab, for each 32-bin half h (HALF_U32 u32 each):
//   planes 4q..4q+3 (q = 0, 1, 2):  u32 [h*896 + q*256 + s*4 + (p - 4q)]     one 16-byte item per sample
//   planes 12, 13:                  u32 [h*896 + 768   + s*2 + (p - 12)]      one  8-byte item per sample
// so a thread fetches the 14 planes of a sample with three LDS.128 and one LDS.64, and the 64 samples of a
// plane group are adjacent items: a warp's operand loads are conflict free (row operands broadcast).
// Slab order in HBM: index ((k * W + w) * nb + b).  Samples past n in the last block are zero.
constexpr int QUAD_U32 = TS * 4;           // u32 per plane quad of one half (256)
constexpr int TAIL_U32 = 3 * QUAD_U32;     // offset of the plane 12/13 pair inside a half (768)
struct DevDb {
    const uint32_t *tiled;
    uint64_t n;   // samples
    uint32_t nb;  // sample blocks
    uint32_t W;   // 64-bin words per sketch (sketchsize64)
    uint32_t K;   // k-mer lengths
    uint32_t pad_;
};

__host__ __device__ inline uint64_t slab_index(const DevDb &d, uint32_t k, uint32_t w, uint32_t b) {
    return ((uint64_t)k * d.W + w) * d.nb + b;
}

#ifdef __CUDACC__
// Total-order key for non-NaN floats, so that (key, neighbour index) sorts as one u64 composite.
__device__ __forceinline__ uint32_t float_order_bits(float f) {
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float float_from_order_bits(uint32_t b) {
    return __uint_as_float((b & 0x80000000u) ? (b & 0x7fffffffu) : ~b);
}
#endif

enum : int { KIND_CORE_ACC = 0, KIND_JACCARD = 1, KIND_ANI = 2 };

// Everything the fused epilogue needs.  Tables are indexed by the per-k match count (`samebits`).
struct EpiParams {
    int kind;
    int k0;          // first k index the tile loop visits
    int nk;          // number of k indices visited (K for core/acc, 1 for single-k modes)
    int kp;          // nk rounded up so that one pair's counts fill whole 16-byte shared-memory words
    int has_c;       // completeness vectors supplied for both sides
    uint32_t nbins;  // 64 * W
    uint32_t expected; // expected_samebits = nbins >> 14
    double cutoff;
    double k_single;   // k as f64 for ANI
    double tolerance;  // ln(2 / (sketch_size * 64)), jaccard.rs:75
    const double *a_comp; // device, per row-side sample (nullable)
    const double *b_comp; // device, per column-side sample (nullable)
    const double *lny;    // [nbins+1] ln(J(count)) computed on the host with the platform libm
    const float *val_tab; // [nbins+1] dense single-k output: (f32)(1-J) or (f32)ani
    const float *key_tab; // [nbins+1] kNN key: (f32)(1-J) or (f32)(1-ani)
    double kf[MAXK];          // k-mer lengths as f64
    double xsum[MAXK + 1];    // running sums after n k-mers, accumulated in the reference's order
    double xbar[MAXK + 1];
    double x_diff[MAXK + 1];
    double xstddev[MAXK + 1];
    double rcp_n[MAXK + 1];   // correctly rounded 1/n, for the exact division-by-n sequence
};

enum : int { MODE_DENSE_SELF = 0, MODE_DENSE_CROSS = 1, MODE_COUNTS = 2, MODE_KNN = 3 };

// Tile schedule of one launch.  Rows are the A side (self: the database, cross: refs), columns the B side.
struct SchedParams {
    int mode;
    int self_mode;        // knn: skip i == j; dense self: only j > i
    uint32_t rb0;         // first row block
    uint32_t n_rb;        // row blocks
    uint32_t group;       // row blocks per L2 group
    uint32_t n_groups;
    const uint64_t *grp_off; // [n_groups + 1] first linear tile of each group
    uint64_t row_begin, row_end; // sample-level row range (clips the first/last block)
    uint64_t out_base;    // dense self: condensed index of (row_begin, row_begin+1)
    uint64_t rect_ld;     // dense self, if non-zero: write kNN keys row-major at [(i - row_begin) * rect_ld + (j - rect_c0)]
    uint64_t rect_c0;     //   (only j > i is written): the symmetric self-kNN path reads them back by row and by column
    uint64_t idx_offset;  // kNN: added to the neighbour index stored in the composite (column shards of a bigger database)
    uint32_t cb_lo, cb_hi; // dense: only column blocks [cb_lo, cb_hi) (cb_hi == 0: all); bands follow an upload in flight
    float *out;           // dense: distances; counts mode: reinterpret as uint32_t*
    uint32_t one;         // always 1: a multiplier the compiler cannot fold, see tile_counts
    uint32_t pad0_;
    // kNN: each CTA owns one row block and column blocks [split * cb_per_split, ...)
    uint32_t knn;
    uint32_t n_split;
    uint32_t cb_per_split;
    uint32_t cap;         // candidate slots per row: a power of two >= knn + TS
    unsigned long long *part_key; // [n_rb][n_split][TS][cap] composite (key bits << 32 | idx); the first knn
                                  // of each list are its best, ascending, when the kernel ends
    float *part_aux;              // same shape: accessory distance (core/acc only)
    // dense modes: match-count scratch in L2 (SM count x 8192 pairs x kp counts, behind a 256-byte header holding
    // dense_kernel's tile counter): used by dense_kernel and by the one-pass 16-query stream kernel
    void *counts_scratch;
    // the launch goes to dense_kernel (tile groups were built over column block PAIRS); 0 = tile_kernel
    uint32_t dense_pairs;
    // dense_kernel, core/accessory: early exit of a warp's 16 x 64 pair block after the third k-mer length when at
    // most defer_max of its pairs can still need the regression; those go to this list (defer_cap entries, fill
    // count in the scratch header) and are finished by defer_kernel.  defer_cap == 0 disables.
    uint4 *defer_list;               // two uint4 per entry: (row, column, output index lo, hi), (count k0, k1, k2, -)
    uint32_t defer_cap, defer_max;
    const uint64_t *a_rows, *b_rows; // sample-major copies of the planes (u64 [sample][k][w][plane]) for defer_kernel
    unsigned long long *work_stats;  // nullable: [1] += warp x word steps in all, [3] += steps skipped, [2] += deferred pairs
};

}  // namespace skd
