// CUDA kernels (sm_100a) for the pairwise sketch-distance path.
//
//   repack_kernel      reference .skd layout -> tiled HBM layout (layout.h)            [HBM bound]
//   tile_kernel        64x64 pair tiles: TMA-staged bit planes, LOP3/POPC match counts,
//                      fused Jaccard / ANI / core-accessory epilogue, dense or kNN output [INT-ALU bound]
//   knn_merge_kernel   merges per-split candidate lists into the final sparse rows
//   pair_kernel        explicit (ref, query) pair list: counts and distances (debug / candidate lists)
//
// Arithmetic restated from the reference: src/distances/jaccard.rs (whole file).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>

#include "kernels.h"
#include "layout.h"

namespace skd {

// ------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + bulk async copy (TMA, SASS UBLKCP)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// Non-blocking phase probe: lets a warp learn early (mid-iteration) that the next stage has landed, so the
// ~90-cycle latency of the phase check is off the critical path at the top of the next iteration.
__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity));
    return ok != 0;
}
// Wait with back-off for a thread that has nothing else to do (the dense kernel's producer lane).  try_wait alone
// returns within a few cycles when the phase is still open, and the retry loop (SYNCS + BRA + YIELD) then eats the
// issue slots of the scheduler the lane shares with consumer warps (ncu, round 2: 14.7 % of all warp instructions of
// the kernel, ~40 % of that scheduler's slots).  A sleeping warp issues nothing.
#ifndef SKD_PROD_SLEEP_NS
#define SKD_PROD_SLEEP_NS 200
#endif
__device__ __forceinline__ void mbar_wait_sleep(uint64_t *bar, uint32_t parity) {
    while (!mbar_test(bar, parity)) {
        if (SKD_PROD_SLEEP_NS > 0) __nanosleep(SKD_PROD_SLEEP_NS);
    }
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ------------------------------------------------------------------------------------------------
// Loader: reference layout (u64 [sample][k][w][plane]) -> tiled layout.  One CTA per (block, w, k).
__global__ void __launch_bounds__(256) repack_kernel(const uint64_t *__restrict__ raw, uint64_t raw_samples,
                                                     uint64_t sample_stride, uint64_t kmer_stride,
                                                     uint32_t *__restrict__ tiled, DevDb db, uint32_t b0) {
    __shared__ uint64_t sh[TS][BBITS + 1];
    const uint32_t b = b0 + blockIdx.x, w = blockIdx.y, k = blockIdx.z;
    for (int t = threadIdx.x; t < TS * BBITS; t += blockDim.x) {
        int s = t / BBITS, p = t % BBITS;
        uint64_t ls = (uint64_t)blockIdx.x * TS + s;  // sample index inside `raw`
        uint64_t v = 0;
        if (ls < raw_samples) v = raw[ls * sample_stride + (uint64_t)k * kmer_stride + (uint64_t)w * BBITS + p];
        sh[s][p] = v;
    }
    __syncthreads();
    uint32_t *dst = tiled + slab_index(db, k, w, b) * STAGE_U32;
    for (int v = threadIdx.x; v < 2 * 4 * TS; v += blockDim.x) {
        const int half = v / (4 * TS), r = v % (4 * TS), q = r / TS, s = r % TS;
        const int shift = 32 * half;
        if (q < 3) {
            uint4 o;
            o.x = (uint32_t)(sh[s][4 * q + 0] >> shift);
            o.y = (uint32_t)(sh[s][4 * q + 1] >> shift);
            o.z = (uint32_t)(sh[s][4 * q + 2] >> shift);
            o.w = (uint32_t)(sh[s][4 * q + 3] >> shift);
            *reinterpret_cast<uint4 *>(dst + half * HALF_U32 + q * QUAD_U32 + s * 4) = o;
        } else {
            uint2 o;
            o.x = (uint32_t)(sh[s][12] >> shift);
            o.y = (uint32_t)(sh[s][13] >> shift);
            *reinterpret_cast<uint2 *>(dst + half * HALF_U32 + TAIL_U32 + s * 2) = o;
        }
    }
}

void launch_repack(const uint64_t *raw, uint64_t raw_samples, uint64_t sample_stride, uint64_t kmer_stride,
                   uint32_t *tiled, const DevDb &db, uint32_t b0, uint32_t n_blocks, cudaStream_t st) {
    dim3 grid(n_blocks, db.W, db.K);
    repack_kernel<<<grid, 256, 0, st>>>(raw, raw_samples, sample_stride, kmer_stride, tiled, db, b0);
}

// The inverse, tiled -> sample-major u64 [sample][k][w][plane] (the reference's .skd layout): one sample's planes of one
// k-mer length are W x 112 contiguous bytes.  Built on demand for defer_kernel, which compares single pairs: from the
// tiled layout a pair costs 16-byte pieces out of 2 x 8 x W x 8 different slabs (half of every 32-byte sector wasted).
__global__ void __launch_bounds__(256) untile_kernel(const __grid_constant__ DevDb db, uint64_t *__restrict__ rows) {
    __shared__ uint64_t sh[TS][BBITS + 1];
    const uint32_t b = blockIdx.x, w = blockIdx.y, k = blockIdx.z;
    const uint32_t *src = db.tiled + slab_index(db, k, w, b) * STAGE_U32;
    for (int t = threadIdx.x; t < TS * BBITS; t += blockDim.x) {
        const int p = t / TS, sm = t % TS;  // consecutive threads: consecutive samples of one plane
        const int q = p / 4;
        uint32_t lo, hi;
        if (q < 3) {
            lo = src[q * QUAD_U32 + sm * 4 + (p - 4 * q)];
            hi = src[HALF_U32 + q * QUAD_U32 + sm * 4 + (p - 4 * q)];
        } else {
            lo = src[TAIL_U32 + sm * 2 + (p - 12)];
            hi = src[HALF_U32 + TAIL_U32 + sm * 2 + (p - 12)];
        }
        sh[sm][p] = (uint64_t)lo | ((uint64_t)hi << 32);
    }
    __syncthreads();
    const uint64_t sample_stride = (uint64_t)db.K * db.W * BBITS;
    for (int t = threadIdx.x; t < TS * BBITS; t += blockDim.x) {
        const int sm = t / BBITS, p = t % BBITS;
        const uint64_t gs = (uint64_t)b * TS + sm;
        if (gs < db.n) rows[gs * sample_stride + ((uint64_t)k * db.W + w) * BBITS + p] = sh[sm][p];
    }
}
void launch_untile(const DevDb &db, uint64_t *rows, cudaStream_t st) {
    if (db.n == 0) return;
    untile_kernel<<<dim3(db.nb, db.W, db.K), 256, 0, st>>>(db, rows);
}

// ------------------------------------------------------------------------------------------------
// Epilogue arithmetic.  All f64 operations use the explicit round-to-nearest intrinsics so that no
// FMA contraction can change a rounding relative to the reference's (jaccard.rs) operation order.
__device__ __forceinline__ double jaccard_from_count(uint32_t c, const EpiParams &E) {
    // jaccard.rs:26-33
    uint32_t diff = c > E.expected ? c - E.expected : 0u;
    double intersize = __ddiv_rn(__dmul_rn((double)diff, (double)E.nbins), (double)(E.nbins - E.expected));
    return __ddiv_rn(intersize, (double)E.nbins);
}
__device__ __forceinline__ double corrected_jaccard(uint32_t c, double c1, double c2, const EpiParams &E) {
    // jaccard.rs:36-42, 55-57
    double j = jaccard_from_count(c, E);
    double c12 = __dmul_rn(c1, c2);
    if (c12 >= E.cutoff) {
        j = __ddiv_rn(j, __ddiv_rn(c12, __dsub_rn(__dadd_rn(c1, c2), c12)));
        j = fmin(j, 1.0);
    }
    return j;
}
__device__ __forceinline__ double ani_pois_dev(double j, double k) {
    // jaccard.rs:49-51
    double q = __ddiv_rn(__dmul_rn(2.0, j), __dadd_rn(1.0, j));
    return fmax(0.0, __dadd_rn(1.0, __dmul_rn(__ddiv_rn(1.0, k), log(q))));
}

// x / n for a small integer n, correctly rounded like IEEE division (so bit-identical to the reference's
// f64 `/`), from the correctly rounded reciprocal: two FMA residual corrections (Markstein).  The first
// makes the quotient faithful, the second rounds it correctly.  Inputs here are far from the
// overflow/underflow ranges where the sequence would need a slow path (zero, inf and NaN pass through).
__device__ __forceinline__ double div_by_n(double x, double dn, double rcp) {
    // tiny or non-finite x (never seen in practice): the plain division is exact by definition
    const int hi = __double2hiint(x) & 0x7fffffff;
    if (__builtin_expect(hi < 0x05000000 || hi >= 0x7a000000, 0)) return __ddiv_rn(x, dn);
    double q = __dmul_rn(x, rcp);
    q = __fma_rn(__fma_rn(-dn, q, x), rcp, q);
    return __fma_rn(__fma_rn(-dn, q, x), rcp, q);
}

// simple_linear_regression (jaccard.rs:105-142) with the x statistics tabulated per n.
__device__ __forceinline__ void regress(double ysum, double xysum, double ysq, int n, const EpiParams &E,
                                        float &core_out, float &acc_out) {
    if (n < 3) {  // ysum cannot be NaN or -inf here: such a y breaks the loop first
        core_out = 1.0f;
        acc_out = 1.0f;
        return;
    }
    const double dn = (double)n, rcp = E.rcp_n[n];
    const double xsum = E.xsum[n];
    double ybar = div_by_n(ysum, dn, rcp);
    double y_diff = __dsub_rn(ysq, div_by_n(__dmul_rn(ysum, ysum), dn, rcp));
    double ystddev = __dsqrt_rn(div_by_n(y_diff, dn, rcp));
    double r = __ddiv_rn(__dsub_rn(xysum, div_by_n(__dmul_rn(xsum, ysum), dn, rcp)),
                         __dsqrt_rn(__dmul_rn(E.x_diff[n], y_diff)));
    double beta = __ddiv_rn(__dmul_rn(r, ystddev), E.xstddev[n]);
    double alpha = __dadd_rn(__dmul_rn(-beta, E.xbar[n]), ybar);
    double core = 0.0, acc = 0.0;
    if (beta < 0.0) {
        core = __dsub_rn(1.0, exp(beta));
    } else if (r > 0.0) {
        core = 1.0;
    }
    if (alpha < 0.0) acc = __dsub_rn(1.0, exp(alpha));
    core_out = (float)core;
    acc_out = (float)acc;
}

// ------------------------------------------------------------------------------------------------
// The tile kernel.
//
// Shared memory:  [NST stages x (A slab 7168 B + B slab 7168 B)] [counts nk x 4096 x CountT]
//                 [mbarriers] [kNN candidate lists]
// Thread tile: rows  arow0 + 4*i (i<4), cols bcol0 + 8*j (j<4) with arow0 = 16*(warp/2) + lane/8,
// bcol0 = 32*(warp%2) + lane%8.  For one plane pair a warp's four A loads are 4 distinct 8-byte words
// (broadcast) and its B loads 8 adjacent ones: no bank conflicts.
struct TileCoord {
    uint32_t rb, cb;
};

// Mismatch counts of one 32-bin half for a thread's 4 x JN pair tile (JN <= TJ column groups of 8 samples).
// JN < TJ is the narrow-column variant for cross runs with few queries: the column operands that do not
// exist are neither loaded nor compared, which turns the launch from INT-bound into memory-bound.
template <int JN>
__device__ __forceinline__ void half_mismatches(const uint32_t *sa, const uint32_t *sb, int arow0, int bcol0,
                                                uint32_t one, uint32_t (&mism)[4][TJ]) {
    uint32_t d[4][JN];
    const uint4 *a4 = reinterpret_cast<const uint4 *>(sa);
    const uint4 *b4 = reinterpret_cast<const uint4 *>(sb);
#pragma unroll
    for (int q = 0; q < 3; ++q) {  // planes 4q .. 4q+3: one LDS.128 per sample
        uint4 av[4], bv[JN];
#pragma unroll
        for (int i = 0; i < 4; ++i) av[i] = a4[q * TS + arow0 + 4 * i];
#pragma unroll
        for (int j = 0; j < JN; ++j) bv[j] = b4[q * TS + bcol0 + 8 * j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < JN; ++j) {
                if (q == 0)
                    d[i][j] = av[i].x ^ bv[j].x;
                else
                    d[i][j] |= av[i].x ^ bv[j].x;
                d[i][j] |= av[i].y ^ bv[j].y;
                d[i][j] |= av[i].z ^ bv[j].z;
                d[i][j] |= av[i].w ^ bv[j].w;
            }
    }
    {  // planes 12, 13: one LDS.64 per sample
        const uint2 *a2 = reinterpret_cast<const uint2 *>(sa + TAIL_U32);
        const uint2 *b2 = reinterpret_cast<const uint2 *>(sb + TAIL_U32);
        uint2 av[4], bv[JN];
#pragma unroll
        for (int i = 0; i < 4; ++i) av[i] = a2[arow0 + 4 * i];
#pragma unroll
        for (int j = 0; j < JN; ++j) bv[j] = b2[bcol0 + 8 * j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < JN; ++j) {
                d[i][j] |= av[i].x ^ bv[j].x;
                d[i][j] |= av[i].y ^ bv[j].y;
            }
    }
    // accumulate on the FMA pipe (IMAD with a run-time 1) so the ALU pipe only sees the LOP3s
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < JN; ++j) mism[i][j] = __popc(d[i][j]) * one + mism[i][j];
}

// COLS = number of leading columns of the tile that are computed (8, 16, 32 or 64 = all).
template <typename CountT, int COLS>
__device__ __forceinline__ void tile_counts(const DevDb &A, const DevDb &B, uint32_t rb, uint32_t cb, int k0, int nk,
                                            uint32_t *stage_mem, uint64_t *full_bar, CountT *counts, int kp,
                                            uint32_t one, uint64_t row_end, uint32_t &git) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int arow0 = (TJ == 4 ? 16 * (warp >> 1) : 16 * warp) + (lane >> 3);
    const int bcol0 = (TJ == 4 ? 32 * (warp & 1) : 0) + (lane & 7);
    (void)row_end;
    const uint32_t W = A.W;
    const uint32_t n_it = (uint32_t)nk * W;

    auto issue = [&](uint32_t it) {
        uint32_t k = (uint32_t)k0 + it / W, w = it % W;
        uint32_t g = git + it;
        uint32_t *dst = stage_mem + (g % NST) * (2 * STAGE_U32);
        uint64_t *bar = &full_bar[g % NST];
        mbar_expect_tx(bar, 2 * STAGE_BYTES);
        bulk_g2s(dst, A.tiled + slab_index(A, k, w, rb) * STAGE_U32, STAGE_BYTES, bar);
        bulk_g2s(dst + STAGE_U32, B.tiled + slab_index(B, k, w, cb) * STAGE_U32, STAGE_BYTES, bar);
    };

    // One CTA barrier per pipeline iteration frees the stage consumed in the previous iteration; thread 0
    // then refills it (two iterations ahead).  A variant with per-stage "empty" mbarriers and a rotating
    // producer warp (no CTA barrier) measured 4 % slower on B200 and was dropped (profiles/README.md).
    if (tid == 0) {
        uint32_t pre = n_it < (uint32_t)(NST - 1) ? n_it : (uint32_t)(NST - 1);
        for (uint32_t p = 0; p < pre; ++p) issue(p);  // all stages are free at tile start (CTA barrier before)
    }

    uint32_t mism[4][TJ];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < TJ; ++j) mism[i][j] = 0;

    uint32_t w_in_k = 0, kk = 0;
    bool ready = false;
    for (uint32_t it = 0; it < n_it; ++it) {
        if (it > 0) __syncthreads();  // every warp is done with iteration it-1: its stage may be refilled
        if (tid == 0 && it + NST - 1 < n_it) issue(it + NST - 1);
        const uint32_t g = git + it;
        if (!ready) mbar_wait(&full_bar[g % NST], (g / NST) & 1u);
        const uint32_t *sa = stage_mem + (g % NST) * (2 * STAGE_U32);
        const uint32_t *sb = sa + STAGE_U32;

        if (COLS == TS) {
            half_mismatches<TJ>(sa, sb, arow0, bcol0, one, mism);
            ready = it + 1 < n_it && mbar_test(&full_bar[(g + 1) % NST], ((g + 1) / NST) & 1u);
            half_mismatches<TJ>(sa + HALF_U32, sb + HALF_U32, arow0, bcol0, one, mism);
        } else if (bcol0 < COLS) {  // narrow variant: warps whose columns lie past COLS only keep the barriers
            constexpr int JN = COLS >= 8 * TJ ? TJ : COLS / 8;
            half_mismatches<JN>(sa, sb, arow0, bcol0, one, mism);
            half_mismatches<JN>(sa + HALF_U32, sb + HALF_U32, arow0, bcol0, one, mism);
        }

        if (++w_in_k == W) {  // finished this k: park the match counts in shared memory
            const uint32_t nbins = 64u * W;
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < TJ; ++j) {
                    counts[((arow0 + 4 * i) * TS + bcol0 + 8 * j) * kp + kk] = (CountT)(nbins - mism[i][j]);
                    mism[i][j] = 0;
                }
            w_in_k = 0;
            ++kk;
        }
    }
    git += n_it;
    __syncthreads();  // counts visible to all; all stages consumed
}

// One pair's distances from its counts in shared memory ([pair][kp], read as 16-byte words).
// Returns d0 (first distance or kNN key) and d1 (accessory distance).
template <typename CountT, bool COMP>
__device__ __forceinline__ void pair_epilogue(const CountT *counts, int pair, const EpiParams &E, uint64_t gi,
                                              uint64_t gj, bool for_knn, float &d0, float &d1) {
    constexpr int PER = 16 / (int)sizeof(CountT);  // counts per 16-byte word
    double c1 = 0.0, c2 = 0.0;
    if (COMP) {
        c1 = E.a_comp[gi];
        c2 = E.b_comp[gj];
    }
    const uint4 *words = reinterpret_cast<const uint4 *>(counts + (size_t)pair * E.kp);
    if (E.kind == KIND_CORE_ACC) {
        // core_acc_dist, jaccard.rs:73-100
        double ysum = 0.0, xysum = 0.0, ysq = 0.0;
        int n = 0;
        bool alive = true;
        for (int kb = 0; kb < E.nk && alive; kb += PER) {
            const uint4 v = words[kb / PER];
            const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
            // the ln J look-ups of this word are independent: issue them together
            double y[PER];
#pragma unroll
            for (int t = 0; t < PER; ++t) {
                const uint32_t c = sizeof(CountT) == 2 ? ((w4[t >> 1] >> (16 * (t & 1))) & 0xffffu) : w4[t];
                if (COMP)
                    y[t] = kb + t < E.nk ? log(corrected_jaccard(c, c1, c2, E)) : 0.0;
                else
                    y[t] = __ldg(&E.lny[c]);  // padding counts are 0 -> a valid table index
            }
#pragma unroll
            for (int t = 0; t < PER; ++t) {
                if (alive && kb + t < E.nk) {
                    if (y[t] < E.tolerance) {
                        alive = false;  // the reference breaks: later k are ignored (jaccard.rs:89-91)
                    } else {
                        ysum = __dadd_rn(ysum, y[t]);
                        xysum = __dadd_rn(xysum, __dmul_rn(E.kf[kb + t], y[t]));
                        ysq = __dadd_rn(ysq, __dmul_rn(y[t], y[t]));
                        ++n;
                    }
                }
            }
        }
        regress(ysum, xysum, ysq, n, E, d0, d1);
    } else {
        const uint32_t c = sizeof(CountT) == 2 ? (words[0].x & 0xffffu) : words[0].x;
        d1 = 0.0f;
        if (!COMP) {
            d0 = for_knn ? __ldg(&E.key_tab[c]) : __ldg(&E.val_tab[c]);
        } else {
            double j = corrected_jaccard(c, c1, c2, E);
            if (E.kind == KIND_ANI) {
                double a = ani_pois_dev(j, E.k_single);
                d0 = for_knn ? (float)__dsub_rn(1.0, a) : (float)a;
            } else {
                d0 = (float)__dsub_rn(1.0, j);
            }
        }
    }
}

// No key exceeds 1.0 (distances and 1 - ANI lie in [0, 1]).  A candidate with key 1.0 and neighbour index j is
// preceded in (key, index) order by every sample with a smaller index, and all of those are candidates of the same
// row somewhere in the run -- so it can only be among the row's knn best if fewer than knn samples precede it.
// Composites at or above this bound never reach a final list; rejecting them up front keeps the lists from filling
// (and being sorted) with the distance-1 pairs that make up nearly everything between unrelated samples.
// first_useless = knn + 1 in self mode (the row itself is no candidate), knn in cross mode.
__device__ __forceinline__ unsigned long long useless_from(uint32_t first_useless) {
    return ((unsigned long long)float_order_bits(1.0f) << 32) | first_useless;
}

// Warp-level bitonic sort of P (power of two) u64 keys (+ optional aux payload) in shared memory.
__device__ __forceinline__ void warp_bitonic_sort(unsigned long long *key, float *aux, int P, int lane) {
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncwarp();
            for (int t = lane; t < P / 2; t += 32) {
                int lo = 2 * t - (t & (stride - 1));  // index with bit `stride` clear
                int hi = lo + stride;
                bool up = (lo & size) == 0;
                unsigned long long a = key[lo], b = key[hi];
                if ((a > b) == up) {
                    key[lo] = b;
                    key[hi] = a;
                    if (aux) {
                        float x = aux[lo];
                        aux[lo] = aux[hi];
                        aux[hi] = x;
                    }
                }
            }
        }
    }
    __syncwarp();
}

template <typename CountT, int COLS>
__global__ void __launch_bounds__(THREADS, 2)
    tile_kernel(const __grid_constant__ DevDb A, const __grid_constant__ DevDb B, const __grid_constant__ EpiParams E,
                const __grid_constant__ SchedParams S) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint32_t *stage_mem = reinterpret_cast<uint32_t *>(smem_raw);
    CountT *counts = reinterpret_cast<CountT *>(smem_raw + NST * 2 * STAGE_BYTES);
    size_t off = (size_t)NST * 2 * STAGE_BYTES + (size_t)E.kp * TS * TS * sizeof(CountT);
    off = (off + 15) & ~(size_t)15;
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem_raw + off);
    off += 64;

    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < NST; ++s) mbar_init(&full_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // padding slots of the per-pair count words must hold a valid table index
    if (E.kp != E.nk)
        for (int e = tid; e < TS * TS * E.kp; e += THREADS) counts[e] = 0;
    __syncthreads();
    uint32_t git = 0;

    if (S.mode != MODE_KNN) {
        // ---- dense / counts: one tile per CTA; linear tile id -> (row block, col block) through L2 groups
        const uint64_t t = blockIdx.x;
        uint32_t lo = 0, hi = S.n_groups;  // largest g with grp_off[g] <= t
        while (hi - lo > 1) {
            uint32_t mid = (lo + hi) >> 1;
            if (S.grp_off[mid] <= t) lo = mid; else hi = mid;
        }
        const uint32_t g = lo;
        const uint32_t rows_in_g = min(S.group, S.n_rb - g * S.group);
        const uint64_t local = t - S.grp_off[g];
        const uint32_t rb = S.rb0 + g * S.group + (uint32_t)(local % rows_in_g);
        const uint32_t cb_first = max(S.mode == MODE_DENSE_SELF ? S.rb0 + g * S.group : 0u, S.cb_lo);
        const uint32_t cb = cb_first + (uint32_t)(local / rows_in_g);
        if (S.mode == MODE_DENSE_SELF && cb < rb) return;  // below the diagonal (only inside a diagonal group)

        tile_counts<CountT, COLS>(A, B, rb, cb, E.k0, E.nk, stage_mem, full_bar, counts, E.kp, S.one, S.row_end, git);

        const uint64_t nA = A.n, nB = B.n;
        for (int m = 0; m < (TS * TS) / THREADS; ++m) {
            const int pair = m * THREADS + tid;
            const uint64_t gi = (uint64_t)rb * TS + pair / TS, gj = (uint64_t)cb * TS + pair % TS;
            if (gi < S.row_begin || gi >= S.row_end || gj >= nB) continue;
            if (S.mode == MODE_COUNTS) {
                uint32_t *o = reinterpret_cast<uint32_t *>(S.out) + ((gi - S.row_begin) * nB + gj) * (uint64_t)E.nk;
                for (int k = 0; k < E.nk; ++k) o[k] = counts[pair * E.kp + k];
                continue;
            }
            uint64_t idx;
            const bool rect = S.rect_ld != 0;
            if (S.mode == MODE_DENSE_SELF) {
                if (gj <= gi) continue;
                if (rect)
                    idx = (gi - S.row_begin) * S.rect_ld + (gj - S.rect_c0);
                else
                    idx = nA * gi - ((gi * (gi + 1)) >> 1) + gj - 1 - gi - S.out_base;  // square_to_condensed
            } else if (rect) {
                idx = (gi - S.row_begin) * S.rect_ld + gj;  // ref-major kNN keys for select_few_kernel
            } else {
                idx = (gi - S.row_begin) * nB + gj;  // ref-major
            }
            float d0, d1;
#ifdef SKD_ABLATE_EPI
            d0 = (float)counts[pair * E.kp];
            d1 = 0.f;
#else
            if (E.has_c)
                pair_epilogue<CountT, true>(counts, pair, E, gi, gj, rect, d0, d1);
            else
                pair_epilogue<CountT, false>(counts, pair, E, gi, gj, rect, d0, d1);
#endif
            if (rect) {  // read back soon by the select kernels: normal (L2-resident) stores
                if (E.kind == KIND_CORE_ACC)
                    reinterpret_cast<float2 *>(S.out)[idx] = make_float2(d0, d1);
                else
                    S.out[idx] = d0;
                continue;
            }
            // results are written once and never re-read by the GPU: streaming stores keep the database in L2
            if (E.kind == KIND_CORE_ACC)
                __stcs(reinterpret_cast<float2 *>(S.out) + idx, make_float2(d0, d1));
            else
                __stcs(S.out + idx, d0);
        }
        return;
    }

    // ---- kNN: this CTA owns row block rb and scans column blocks [cb_lo, cb_hi).
    // Per-row candidate lists live in global memory (L2 resident; appends are rare once the per-row
    // threshold has tightened), so that shared memory stays small enough for two CTAs per SM.  Only the
    // thresholds and fill counts are in shared memory; the pipeline stage buffers double as sort scratch
    // between tiles, when no bulk copy is in flight.
    const bool with_aux = E.kind == KIND_CORE_ACC;
    unsigned long long *thr = reinterpret_cast<unsigned long long *>(smem_raw + off);
    off += TS * sizeof(unsigned long long);
    uint32_t *cand_cnt = reinterpret_cast<uint32_t *>(smem_raw + off);

    const uint32_t rbl = blockIdx.x / S.n_split, split = blockIdx.x % S.n_split;
    const uint32_t rb = S.rb0 + rbl;
    const uint32_t cb_lo = split * S.cb_per_split;
    const uint32_t cb_hi = min(B.nb, cb_lo + S.cb_per_split);
    const int warp = tid >> 5, lane = tid & 31;
    const uint32_t cap = S.cap, knn = S.knn;
    // this CTA's lists: [row][cap]
    unsigned long long *cand_key = S.part_key + ((uint64_t)rbl * S.n_split + split) * TS * cap;
    float *cand_aux = with_aux ? S.part_aux + ((uint64_t)rbl * S.n_split + split) * TS * cap : nullptr;
    // sort scratch: one (cap keys + cap aux) region per sorting warp, carved from the stage buffers
    const uint32_t scratch_bytes = cap * 12;
    const int sorters = min(THREADS / 32, (int)((NST * 2 * STAGE_BYTES) / scratch_bytes));
    unsigned long long *sc_key = reinterpret_cast<unsigned long long *>(smem_raw + (size_t)warp * scratch_bytes);
    float *sc_aux = reinterpret_cast<float *>(sc_key + cap);

    // Sort one row's list through shared memory; keeps the best `keep` entries (ascending) in place.
    auto sort_row = [&](int row, uint32_t count, uint32_t keep) {
        unsigned long long *gk = cand_key + (size_t)row * cap;
        float *ga = with_aux ? cand_aux + (size_t)row * cap : nullptr;
        for (uint32_t e = lane; e < cap; e += 32) {
            sc_key[e] = e < count ? __ldcg(gk + e) : ~0ull;
            if (with_aux) sc_aux[e] = e < count ? __ldcg(ga + e) : 0.0f;
        }
        warp_bitonic_sort(sc_key, with_aux ? sc_aux : nullptr, (int)cap, lane);
        for (uint32_t e = lane; e < keep; e += 32) {
            gk[e] = sc_key[e];
            if (with_aux) ga[e] = sc_aux[e];
        }
        __syncwarp();
    };

    for (int r = tid; r < TS; r += THREADS) {
        thr[r] = useless_from(knn + (S.self_mode ? 1u : 0u));
        cand_cnt[r] = 0;
    }
    __syncthreads();

    for (uint32_t cb = cb_lo; cb < cb_hi; ++cb) {
        tile_counts<CountT, COLS>(A, B, rb, cb, E.k0, E.nk, stage_mem, full_bar, counts, E.kp, S.one, S.row_end, git);
        for (int m = 0; m < (TS * TS) / THREADS; ++m) {
            const int pair = m * THREADS + tid;
            const int row = pair / TS;
            const uint64_t gi = (uint64_t)rb * TS + row, gj = (uint64_t)cb * TS + pair % TS;
            if (gi < S.row_begin || gi >= S.row_end || gj >= B.n) continue;
            if (S.self_mode && gi == gj) continue;
            float d0, d1;
            if (E.has_c)
                pair_epilogue<CountT, true>(counts, pair, E, gi, gj, true, d0, d1);
            else
                pair_epilogue<CountT, false>(counts, pair, E, gi, gj, true, d0, d1);
            unsigned long long comp = ((unsigned long long)float_order_bits(d0) << 32) | (uint32_t)(gj + S.idx_offset);
            if (comp < thr[row]) {
                uint32_t slot = atomicAdd(&cand_cnt[row], 1u);
                cand_key[(size_t)row * cap + slot] = comp;
                if (with_aux) cand_aux[(size_t)row * cap + slot] = d1;
            }
        }
        __syncthreads();  // appends (global) visible CTA-wide; stage buffers idle until the next tile
        // compact rows that could overflow during the next tile (a tile adds at most TS entries per row)
        if (warp < sorters) {
            for (int row = warp; row < TS; row += sorters) {
                const uint32_t c = cand_cnt[row];
                if (c + TS > cap) {
                    sort_row(row, c, knn);
                    if (lane == 0) {
                        cand_cnt[row] = knn;
                        thr[row] = sc_key[knn - 1];
                    }
                    __syncwarp();
                }
            }
        }
        // the scratch lives in the stage buffers the next tile's bulk copies (async proxy) will overwrite
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
    }
    // final: each row's best knn of this split, ascending, padded with ~0
    if (warp < sorters) {
        for (int row = warp; row < TS; row += sorters) sort_row(row, cand_cnt[row], knn);
    }
}

// ------------------------------------------------------------------------------------------------
// The dense kernel (round 2): the same match-count arithmetic and epilogue as tile_kernel's dense modes, restructured
// around what the loop model (tools/loop_model.cu, profiles/README.md) says bounds this loop: a 4 x 4 pair tile per
// thread needs 16 LDS per 128 LOP3 and tops out at 0.77 of the LOP3 issue rate whatever else is done, a 4 x 8 tile
// needs 12 and reaches 0.92 -- if nothing synchronises the CTA.
//   * CTA tile 64 rows x 128 columns (one row block, two adjacent column blocks), 8 consumer warps of 16 x 64 pairs,
//     4 x 8 pairs per thread (~150 registers: one CTA per SM), plus one producer warp.
//   * Persistent CTAs take tiles from a global counter (linear tile id -> (row block, column block pair) through the
//     same L2-friendly groups as tile_kernel), so the SMs stay balanced to the last tile.
//   * A ring of DNST stages (one 64-bin word: the row slab + the two column slabs, 21 KB), filled by bulk copies onto
//     per-stage "full" mbarriers.  Consumers never meet at a barrier: a warp that has finished a stage arrives on the
//     stage's "empty" mbarrier and moves on; the producer lane (suspended in try_wait, not polling) refills a stage
//     once all 8 warps have released it, and its cursor runs on into the next tile, so the pipeline never drains
//     between tiles.
//   * Per-pair match counts do not go through shared memory: a thread keeps the counts of its own 32 pairs in a
//     per-CTA scratch area in global memory (L2 resident, [pair slot][thread][kp], one 16-byte word per pair for
//     K <= 8) and runs the fused epilogue for exactly those pairs, so a warp goes from its last word straight into
//     its epilogue while the other warps are still counting (f64/LSU work overlaps LOP3 work) and on into the next
//     tile.  Shared memory holds only the ring, and any number of k-mer lengths fits.
#ifndef SKD_DNST
#define SKD_DNST 6
#endif
constexpr int DNST = SKD_DNST;
static_assert(DNST >= 2 && DNST <= 10, "ring depth");
constexpr int TRING = 16;                          // published tiles kept in shared memory
constexpr uint32_t kCutUnset = 0xffffffffu;
constexpr int DSTAGE_U32 = 3 * STAGE_U32;          // row slab + two column slabs
constexpr int DENSE_THREADS = 288;                 // 8 consumer warps + 1 producer warp
constexpr int DJ = 8;                              // column groups per thread (4 x 8 pair tile)

// Mismatch counts of one 32-bin half for a thread's 4 x 8 pair tile: rows arow0 + 4 i of the row slab, columns
// bcol0 + 8 j (j < 8: all 64 columns of ONE column slab; a warp works on one of the tile's two column slabs).
__device__ __forceinline__ void half_mismatches_4x8(const uint32_t *sa, const uint32_t *sb, int arow0, int bcol0,
                                                    uint32_t one, uint32_t (&mism)[4][DJ]) {
    uint32_t d[4][DJ];
    const uint4 *a4 = reinterpret_cast<const uint4 *>(sa);
    const uint4 *b4 = reinterpret_cast<const uint4 *>(sb);
#pragma unroll
    for (int q = 0; q < 3; ++q) {  // planes 4q .. 4q+3: one LDS.128 per sample
        uint4 av[4], bv[DJ];
#pragma unroll
        for (int i = 0; i < 4; ++i) av[i] = a4[q * TS + arow0 + 4 * i];
#pragma unroll
        for (int j = 0; j < DJ; ++j) bv[j] = b4[q * TS + bcol0 + 8 * j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < DJ; ++j) {
                if (q == 0)
                    d[i][j] = av[i].x ^ bv[j].x;
                else
                    d[i][j] |= av[i].x ^ bv[j].x;
                d[i][j] |= av[i].y ^ bv[j].y;
                d[i][j] |= av[i].z ^ bv[j].z;
                d[i][j] |= av[i].w ^ bv[j].w;
            }
    }
    {  // planes 12, 13: one LDS.64 per sample
        const uint2 *a2 = reinterpret_cast<const uint2 *>(sa + TAIL_U32);
        const uint2 *b2 = reinterpret_cast<const uint2 *>(sb + TAIL_U32);
        uint2 av[4], bv[DJ];
#pragma unroll
        for (int i = 0; i < 4; ++i) av[i] = a2[arow0 + 4 * i];
#pragma unroll
        for (int j = 0; j < DJ; ++j) bv[j] = b2[bcol0 + 8 * j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < DJ; ++j) {
                d[i][j] |= av[i].x ^ bv[j].x;
                d[i][j] |= av[i].y ^ bv[j].y;
            }
    }
    // accumulate on the FMA pipe (IMAD with a run-time 1) so the ALU pipe only sees the LOP3s
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < DJ; ++j) mism[i][j] = __popc(d[i][j]) * one + mism[i][j];
}

// Distances of one pair from its per-k match counts, for dense_kernel: the out-of-line part of the epilogue (pairs
// that really go through the regression, or any pair when completeness vectors are given).  The counts live k-major
// in the CTA's scratch: the word of k is words[k * kstride]; for 16-bit counts a word packs the counts of two
// neighbouring pairs (`shift` selects the half).  Same arithmetic, order and guards as pair_epilogue
// (jaccard.rs:61-142).
template <typename CountT>
__device__ __forceinline__ float2 pair_epilogue_k(const uint32_t *words, size_t kstride, int shift, const EpiParams &E,
                                                  uint64_t gi, uint64_t gj, bool for_knn) {
    constexpr uint32_t MASK = sizeof(CountT) == 2 ? 0xffffu : 0xffffffffu;
    const bool comp = E.has_c != 0;
    double cc1 = 0.0, cc2 = 0.0;
    if (comp) {
        cc1 = E.a_comp[gi];
        cc2 = E.b_comp[gj];
    }
    float d0, d1 = 0.0f;
    if (E.kind == KIND_CORE_ACC) {
        double ysum = 0.0, xysum = 0.0, ysq = 0.0;
        int n = 0;
        bool alive = true;
        for (int kb = 0; kb < E.nk && alive; kb += 8) {
            uint32_t c[8];
            double y[8];
#pragma unroll
            for (int t = 0; t < 8; ++t)
                c[t] = kb + t < E.nk ? (__ldcg(words + (size_t)(kb + t) * kstride) >> shift) & MASK : 0u;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                if (kb + t >= E.nk)
                    y[t] = 0.0;
                else if (comp)
                    y[t] = log(corrected_jaccard(c[t], cc1, cc2, E));
                else
                    y[t] = __ldg(&E.lny[c[t]]);
            }
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                if (alive && kb + t < E.nk) {
                    if (y[t] < E.tolerance) {
                        alive = false;  // the reference breaks: later k are ignored (jaccard.rs:89-91)
                    } else {
                        ysum = __dadd_rn(ysum, y[t]);
                        xysum = __dadd_rn(xysum, __dmul_rn(E.kf[kb + t], y[t]));
                        ysq = __dadd_rn(ysq, __dmul_rn(y[t], y[t]));
                        ++n;
                    }
                }
            }
        }
        regress(ysum, xysum, ysq, n, E, d0, d1);
    } else {  // single k with completeness vectors
        const uint32_t c = (__ldcg(words) >> shift) & MASK;
        const double j = corrected_jaccard(c, cc1, cc2, E);
        if (E.kind == KIND_ANI) {
            const double a = ani_pois_dev(j, E.k_single);
            d0 = for_knn ? (float)__dsub_rn(1.0, a) : (float)a;
        } else {
            d0 = (float)__dsub_rn(1.0, j);
        }
    }
    return make_float2(d0, d1);
}

template <typename CountT, bool DEFER>
__global__ void __launch_bounds__(DENSE_THREADS, 1)
    dense_kernel(const __grid_constant__ DevDb A, const __grid_constant__ DevDb B, const __grid_constant__ EpiParams E,
                 const __grid_constant__ SchedParams S, const uint64_t n_tiles, unsigned long long *__restrict__ tile_counter,
                 CountT *__restrict__ scratch_all) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint32_t *stage_mem = reinterpret_cast<uint32_t *>(smem_raw);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem_raw + (size_t)DNST * DSTAGE_U32 * 4);
    uint64_t *empty_bar = full_bar + DNST;  // 8 arrivals (one per consumer warp) release a stage
    uint2 *tiles = reinterpret_cast<uint2 *>(empty_bar + DNST);
    volatile uint32_t *published = reinterpret_cast<volatile uint32_t *>(tiles + TRING);  // tiles[j % TRING] valid for j < it
    // early exit: votes[q] = consumer warps that stopped counting tile q; cut[q] = stages the producer issued for it
    // (kCutUnset while it may still issue more)
    uint32_t *votes = const_cast<uint32_t *>(published) + 4;
    volatile uint32_t *cut = votes + TRING;
    uint16_t *warp_lists = reinterpret_cast<uint16_t *>(votes + 2 * TRING);  // [8 warps][1024]
    unsigned char *out_stage = reinterpret_cast<unsigned char *>(warp_lists + 8 * 1024);          // [8 warps][16][64] float2

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t W = A.W, nk = (uint32_t)E.nk;
    const uint32_t n_it = nk * W;
    const uint32_t nbins = 64u * W;
    // column blocks this launch may touch: [.., cb_limit) (a band of an upload in flight ends at cb_hi)
    const uint32_t cb_limit = S.cb_hi ? min(S.cb_hi, B.nb) : B.nb;

    if (tid == 0) {
        for (int s = 0; s < DNST; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 8);
        }
        *published = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // Every wait below ends by construction; should a bug ever break that, trap after 20 s instead of hanging the GPU.
    uint64_t t_start;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_start));
    uint32_t spins = 0;
    auto spin_guard = [&]() {
        if ((++spins & 0xffu) == 0) {
            uint64_t now;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
            if (now - t_start > 20000000000ull) __trap();
        }
    };

    if (warp == 8) {
        // ---- producer warp: one lane walks the schedule (tiles from the global counter, words of a tile in order)
        // and refills a stage as soon as all 8 consumer warps have released its previous word
        if (lane != 0) return;
        uint32_t grp_hint = 0;
        auto fetch_tile = [&]() -> uint2 {  // next valid tile (rb, first cb); (0xffffffff, 0) = schedule exhausted
            for (;;) {
                const uint64_t t = atomicAdd(tile_counter, 1ull);
                if (t >= n_tiles) return make_uint2(0xffffffffu, 0u);
                uint32_t g = grp_hint;
                if (!(S.grp_off[g] <= t && t < S.grp_off[g + 1])) {
                    uint32_t lo = 0, hi = S.n_groups;  // largest g with grp_off[g] <= t
                    while (hi - lo > 1) {
                        const uint32_t mid = (lo + hi) >> 1;
                        if (S.grp_off[mid] <= t) lo = mid; else hi = mid;
                    }
                    g = grp_hint = lo;
                }
                const uint32_t rows_in_g = min(S.group, S.n_rb - g * S.group);
                const uint64_t local = t - S.grp_off[g];
                const uint32_t rb = S.rb0 + g * S.group + (uint32_t)(local % rows_in_g);
                const uint32_t cb_first = max(S.mode == MODE_DENSE_SELF ? S.rb0 + g * S.group : 0u, S.cb_lo);
                const uint32_t cb = cb_first + 2u * (uint32_t)(local / rows_in_g);  // column block PAIRS
                if (S.mode == MODE_DENSE_SELF && cb + 1 < rb) continue;  // wholly below the diagonal
                return make_uint2(rb, cb);
            }
        };
        uint32_t n_pub = 0;
        auto publish_tile = [&]() -> uint2 {
            const uint2 rc = fetch_tile();
            tiles[n_pub % TRING] = rc;
            if (DEFER) {
                votes[n_pub % TRING] = 0;
                cut[n_pub % TRING] = kCutUnset;
            }
            __threadfence_block();
            *published = ++n_pub;
            return rc;
        };
        uint2 rc = publish_tile();
        uint32_t s = 0, use = 0;
        const uint64_t a_step = (uint64_t)A.nb * STAGE_U32, b_step = (uint64_t)B.nb * STAGE_U32;  // next (k, w) slab
        while (rc.x != 0xffffffffu) {
            const uint2 next = publish_tile();  // the tile after this one: a whole tile ahead of its first use
            const uint32_t *a_src = A.tiled + slab_index(A, (uint32_t)E.k0, 0, rc.x) * STAGE_U32;
            const uint32_t *b_src = B.tiled + slab_index(B, (uint32_t)E.k0, 0, rc.y) * STAGE_U32;
            // the second column block may not exist (odd count, end of a band): its slab is then not loaded and its
            // columns are masked in the epilogue
            const uint32_t b_bytes = rc.y + 1 < cb_limit ? 2u * STAGE_BYTES : (uint32_t)STAGE_BYTES;
            const uint32_t q = (n_pub - 2) % TRING;  // this tile's slot (the one after it is published already)
            uint32_t it = 0;
            for (; it < n_it; ++it) {
                // every consumer warp has stopped counting this tile: the rest of its words need not be loaded
                if (DEFER && it > 3 * W && *reinterpret_cast<volatile uint32_t *>(votes + q) == 8u) break;
                // wait (asleep between probes) until all 8 consumer warps have released the stage's previous word
#ifndef SKD_ABL_NOWAIT
                if (use > 0) mbar_wait_sleep(&empty_bar[s], (use - 1) & 1u);
#else
                if (use > 0) break;  // ablation: fill the ring once
#endif
                uint32_t *dst = stage_mem + s * DSTAGE_U32;
                uint64_t *bar = &full_bar[s];
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic reads before async writes
                mbar_expect_tx(bar, STAGE_BYTES + b_bytes);
                bulk_g2s(dst, a_src, STAGE_BYTES, bar);
                bulk_g2s(dst + STAGE_U32, b_src, b_bytes, bar);
                a_src += a_step;
                b_src += b_step;
                if (++s == (uint32_t)DNST) {
                    s = 0;
                    ++use;
                }
            }
            if (DEFER) {  // tell the draining warps where this tile ends, BEFORE any stage of the next tile is issued
                cut[q] = it;
                __threadfence_block();
            }
            rc = next;
        }
        return;
    }

    // ---- consumer warps: warp tile 16 rows x 64 columns, thread tile rows arow0 + 4 i, columns bcol0 + 8 j
    const int arow0 = 16 * (warp >> 1) + (lane >> 3);
    const int bcol0 = 64 * (warp & 1) + (lane & 7);            // column inside the 128-column tile
    // this thread's words in the CTA's count scratch, k-major so that a warp's store is one contiguous line:
    // word [k][slot][tid]; 16-bit counts: slot = pair / 2 holds pairs (i, 2 jj) and (i, 2 jj + 1), else slot = pair
    constexpr int PACK = sizeof(CountT) == 2 ? 2 : 1;
    constexpr int SLOTS = 4 * DJ / PACK;
    const size_t kstride = (size_t)SLOTS * 256;  // words between consecutive k
    uint32_t *my_words = reinterpret_cast<uint32_t *>(scratch_all) + (size_t)blockIdx.x * nk * kstride + tid;

    uint32_t mism[4][DJ];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < DJ; ++j) mism[i][j] = 0;
    uint32_t s = 0, ph = 0, w_in_k = 0, kk = 0;
    bool ready = false;
    // Early exit (the reference's own break, jaccard.rs:89-91, at the granularity a warp can afford): once the third
    // k-mer length is counted every pair without a match at one of them is final, (1, 1).  If at most defer_max pairs
    // of the warp's 16 x 64 block are left, they are handed to defer_kernel (one warp per pair, all k) and the warp
    // does not count the remaining k-mer lengths of the tile: it only passes the stages on.  All of it lives in the
    // once-per-k block below; the counting loop carries no extra state (the kernel sits at its 168-register cap).
    const bool may_defer = DEFER && S.defer_cap != 0 && E.kind == KIND_CORE_ACC && nk > 3;
    // core/accessory: bit c of `dead` = pair c is known to be (1, 1); with fewer than three k-mer lengths every pair is
    const bool track_dead = E.kind == KIND_CORE_ACC;
    const uint32_t dead_init = (E.kind == KIND_CORE_ACC && nk < 3) ? 0xffffffffu : 0u;
    uint32_t dead = dead_init;
    const uint64_t nA = A.n;
    const uint64_t col_end = min((uint64_t)B.n, (uint64_t)cb_limit * TS);
    const bool rect = S.rect_ld != 0;
    // operand offsets inside a stage: the warp's column half selects the first or second column slab
    const int b_slab = (warp & 1) * STAGE_U32;
    const int bcol_in_slab = lane & 7;

    for (uint32_t tj = 0;; ++tj) {
        while (*published <= tj) {  // (published a tile ahead of use: spins only at the very start)
            __nanosleep(20);
            spin_guard();
        }
        __threadfence_block();
        const uint2 rc = tiles[tj % TRING];
        if (rc.x == 0xffffffffu) break;
        for (uint32_t it = 0; it < n_it; ++it) {
#ifndef SKD_ABL_NOWAIT
            if (!ready) mbar_wait(&full_bar[s], ph);  // try_wait suspends the warp in hardware: no issue slots burned
#endif
            const uint32_t *sa = stage_mem + s * DSTAGE_U32;
            const uint32_t *sb = sa + STAGE_U32 + b_slab;
            half_mismatches_4x8(sa, sb, arow0, bcol_in_slab, S.one, mism);
#ifndef SKD_ABL_NOWAIT
            {   // probe the next stage now: the ~90-cycle latency of the phase check hides behind the second half
                const uint32_t s1 = s + 1 == (uint32_t)DNST ? 0u : s + 1;
                ready = mbar_test(&full_bar[s1], s1 == 0 ? ph ^ 1u : ph);
            }
#endif
            half_mismatches_4x8(sa + HALF_U32, sb + HALF_U32, arow0, bcol_in_slab, S.one, mism);
#ifndef SKD_ABL_NOWAIT
            __syncwarp();  // every lane has consumed its operands of this stage
            if (lane == 0) mbar_arrive(&empty_bar[s]);  // release
#endif
            if (++w_in_k == W) {  // finished this k
                // A pair without any match at one of the first three k-mer lengths is (1, 1) whatever follows (the
                // reference's loop breaks there with n < 3, jaccard.rs:89-91, 117-119).  With unrelated samples that
                // is nearly every pair: remember it in a bit mask, and neither store nor ever read its counts.
                if (track_dead && kk < 3) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < DJ; ++j)
                            if (nbins - mism[i][j] <= E.expected) dead |= 1u << (i * DJ + j);
                }
                // park the match counts of the pairs still alive (coalesced: 128 contiguous bytes per warp and slot)
                uint32_t *kw = my_words + (size_t)kk * kstride;
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < DJ; j += PACK) {
                        uint32_t v = nbins - mism[i][j];
                        if (PACK == 2) v |= (nbins - mism[i][j + 1]) << 16;
                        const uint32_t both = (PACK == 2 ? 3u : 1u) << (i * DJ + j);
#ifdef SKD_ABL_NOCOUNTS
                        if (v == 0x12345u)
#endif
                        if ((dead & both) != both) __stcg(kw + (size_t)((i * DJ + j) / PACK) * 256, v);
                    }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < DJ; ++j) mism[i][j] = 0;
                w_in_k = 0;
                if (++kk == nk) kk = 0;
                if (may_defer && kk == 3) {
                    // pairs of this thread that are still open: in range, above the diagonal (self), not dead
                    uint32_t open = 0;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint64_t gi = (uint64_t)rc.x * TS + arow0 + 4 * i;
                        if (gi < S.row_begin || gi >= S.row_end) continue;
#pragma unroll
                        for (int j = 0; j < DJ; ++j) {
                            const uint64_t gj = (uint64_t)rc.y * TS + bcol0 + 8 * j;
                            if (gj < col_end && !(S.mode == MODE_DENSE_SELF && gj <= gi)) open |= 1u << (i * DJ + j);
                        }
                    }
                    open &= ~dead;
                    const uint32_t cnt = __popc(open);
                    uint32_t incl = cnt;
#pragma unroll
                    for (int dlt = 1; dlt < 32; dlt <<= 1) {
                        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, dlt);
                        if (lane >= dlt) incl += t;
                    }
                    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
                    if (total <= S.defer_max) {
                        unsigned long long base = 0;
                        if (total) {
                            if (lane == 0) base = atomicAdd(reinterpret_cast<unsigned long long *>(tile_counter) + 1, (unsigned long long)total);
                            base = __shfl_sync(0xffffffffu, base, 0);
                        }
                        const bool fits = base + total <= (unsigned long long)S.defer_cap;
                        unsigned long long slot = base + incl - cnt;
                        for (uint32_t m = open; m; m &= m - 1, ++slot) {
                            if (slot >= (unsigned long long)S.defer_cap) break;
                            uint4 e = make_uint4(0xffffffffu, 0u, 0u, 0u);  // reserved but unused (list overflow)
                            uint4 e2 = make_uint4(0u, 0u, 0u, 0u);           // the three counts already known
                            if (fits) {
                                const int c = __ffs(m) - 1;
                                {   // this thread's own count words of k-mer lengths 0..2 (stored a moment ago)
                                    const uint32_t *cw = my_words + (size_t)(c / PACK) * 256;
                                    const int sh = PACK == 2 ? 16 * (c % PACK) : 0;
                                    const uint32_t msk = PACK == 2 ? 0xffffu : 0xffffffffu;
                                    e2.x = (__ldcg(cw) >> sh) & msk;
                                    e2.y = (__ldcg(cw + kstride) >> sh) & msk;
                                    e2.z = (__ldcg(cw + 2 * kstride) >> sh) & msk;
                                }
                                const uint64_t gi = (uint64_t)rc.x * TS + arow0 + 4 * (c / DJ);
                                const uint64_t gj = (uint64_t)rc.y * TS + bcol0 + 8 * (c % DJ);
                                uint64_t idx;  // output position, the rule of row_base_of below
                                if (rect) idx = (gi - S.row_begin) * S.rect_ld - (S.mode == MODE_DENSE_SELF ? S.rect_c0 : 0) + gj;
                                else if (S.mode == MODE_DENSE_SELF) idx = nA * gi - ((gi * (gi + 1)) >> 1) - 1 - gi - S.out_base + gj;
                                else idx = (gi - S.row_begin) * B.n + gj;
                                e = make_uint4((uint32_t)gi, (uint32_t)gj, (uint32_t)idx, (uint32_t)(idx >> 32));
                            }
                            S.defer_list[2 * slot] = e;
                            S.defer_list[2 * slot + 1] = e2;
                        }
                        if (fits) {  // (warp-uniform)
                            // deferred pairs look dead from here on: the epilogue writes (1, 1) for them and
                            // defer_kernel, which runs after this kernel, overwrites that with their distances
                            dead |= open;
                            if (S.work_stats && lane == 0) {
                                atomicAdd(S.work_stats + 3, (unsigned long long)(n_it - (it + 1)));
                                if (total) atomicAdd(S.work_stats + 2, (unsigned long long)total);
                            }
                            // Pass the remaining stages of this tile on without counting.  When all 8 warps have
                            // voted the producer stops loading the tile and publishes how many stages it issued
                            // (cut); until then a stage that lands belongs to this tile -- the producer sets cut
                            // before it issues anything of the next one.
                            const uint32_t q = tj % TRING;
                            if (lane == 0) atomicAdd(votes + q, 1u);
                            for (uint32_t it2 = it + 1; it2 < n_it; ++it2) {
                                uint32_t s1 = s + 1, ph1 = ph;
                                if (s1 == (uint32_t)DNST) {
                                    s1 = 0;
                                    ph1 ^= 1u;
                                }
                                // lane 0 watches, the warp follows its verdict (one decision per warp, whatever the
                                // lanes would have read individually)
                                uint32_t c = kCutUnset, landed = 0;
                                if (lane == 0) {
                                    for (;;) {
                                        c = cut[q];
                                        if (c != kCutUnset) break;
                                        if (mbar_test(&full_bar[s1], ph1)) {
                                            c = cut[q];  // read again: unset now means the stage that landed is ours
                                            landed = 1;
                                            break;
                                        }
                                        __nanosleep(32);
                                        spin_guard();
                                    }
                                }
                                c = __shfl_sync(0xffffffffu, c, 0);
                                landed = __shfl_sync(0xffffffffu, landed, 0);
                                if (c != kCutUnset && it2 >= c) break;
                                s = s1;
                                ph = ph1;
                                if (!landed) mbar_wait(&full_bar[s], ph);
                                __syncwarp();
                                if (lane == 0) mbar_arrive(&empty_bar[s]);
                            }
                            it = n_it - 1;  // the loop's tail advances the ring once more and ends the tile
                            kk = 0;
                            ready = false;
                        }
                    }
                }
            }
            if (++s == (uint32_t)DNST) {
                s = 0;
                ph ^= 1u;
            }
        }
        // ---- fused epilogue for this thread's own 32 pairs (no CTA-wide step: other warps may still be counting).
        // Pass 1, every lane for its own pairs: core/accessory pairs known to be (1, 1) are stored as such, single-k
        // pairs take one table look-up (or the completeness arithmetic).  Core/accessory pairs that need the
        // regression are only collected: with 1 % of them spread over the lanes, running the regression where they sit
        // would take the whole warp through that long, latency-bound code once per pair slot (measured: 10 % of the
        // kernel).  Pass 2 instead compacts them over the warp (list in shared memory) and gives every lane one.
#ifdef SKD_ABL_NOEPI
        if (dead != 0x12345u) dead = 0x12345u; else
#endif
        {
            const bool two = E.kind == KIND_CORE_ACC;
            const bool single_plain = !two && !E.has_c;
            const float *tab = rect ? E.key_tab : E.val_tab;
            auto row_base_of = [&](uint64_t gi) -> uint64_t {  // output index of column 0 of row gi
                if (rect) return (gi - S.row_begin) * S.rect_ld - (S.mode == MODE_DENSE_SELF ? S.rect_c0 : 0);
                if (S.mode == MODE_DENSE_SELF) return nA * gi - ((gi * (gi + 1)) >> 1) - 1 - gi - S.out_base;  // condensed
                return (gi - S.row_begin) * B.n;
            };
            // Dense outputs are staged per warp in shared memory (16 rows x 64 columns) and written out row by row,
            // 256 contiguous bytes per store: when the destination is pinned host memory (zero-copy results) the
            // PCIe writes are full-size instead of the 64-byte pieces a thread-private store pattern gives.
            // kNN key matrices (rect) stay in HBM and are stored directly.
            float2 *wst2 = reinterpret_cast<float2 *>(out_stage) + warp * (16 * 64);
            float *wst1 = reinterpret_cast<float *>(wst2);
            auto put = [&](int lrow, int lcol, uint64_t idx, float2 d) {
#ifdef SKD_ABL_NOSTORE
                if (d.x == 12345.f)
#endif
                if (rect) {  // read back soon by the select kernels: normal (L2-resident) stores
                    if (two)
                        reinterpret_cast<float2 *>(S.out)[idx] = d;
                    else
                        S.out[idx] = d.x;
                } else if (two) {
                    wst2[lrow * 64 + lcol] = d;
                } else {
                    wst1[lrow * 64 + lcol] = d.x;
                }
            };
            uint32_t todo = 0;  // bit c: pair c needs the regression
#pragma unroll 1
            for (int i = 0; i < 4; ++i) {
                const uint64_t gi = (uint64_t)rc.x * TS + arow0 + 4 * i;
                if (gi < S.row_begin || gi >= S.row_end) continue;
                const uint64_t row_base = row_base_of(gi);
                uint32_t w_single[DJ];  // single-k: the count words of this row's slots, all in flight at once
                if (!two) {
#pragma unroll
                    for (int j = 0; j < DJ; j += PACK) w_single[j] = __ldcg(my_words + (size_t)((i * DJ + j) / PACK) * 256);
                }
#pragma unroll
                for (int j = 0; j < DJ; ++j) {
                    const uint64_t gj = (uint64_t)rc.y * TS + bcol0 + 8 * j;
                    if (gj >= col_end || (S.mode == MODE_DENSE_SELF && gj <= gi)) continue;
                    const int c = i * DJ + j;
                    if (two) {
                        if ((dead >> c) & 1u)
                            put((lane >> 3) + 4 * i, (lane & 7) + 8 * j, row_base + gj, make_float2(1.0f, 1.0f));
                        else
                            todo |= 1u << c;
                    } else if (single_plain) {
                        const uint32_t cnt = (w_single[j - j % PACK] >> (16 * (j % PACK))) & (PACK == 2 ? 0xffffu : 0xffffffffu);
                        put((lane >> 3) + 4 * i, (lane & 7) + 8 * j, row_base + gj, make_float2(__ldg(&tab[cnt]), 0.0f));
                    } else {
                        put((lane >> 3) + 4 * i, (lane & 7) + 8 * j, row_base + gj,
                            pair_epilogue_k<CountT>(my_words + (size_t)(c / PACK) * 256, kstride, 16 * (c % PACK), E, gi, gj, rect));
                    }
                }
            }
            if (two) {  // pass 2 (warp-uniform branch)
                const uint32_t cnt = __popc(todo);
                uint32_t incl = cnt;
#pragma unroll
                for (int dlt = 1; dlt < 32; dlt <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, dlt);
                    if (lane >= dlt) incl += t;
                }
                const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
                if (total) {
                    uint16_t *list = warp_lists + warp * 1024;  // (owner lane << 5) | pair
                    uint32_t r = incl - cnt;
                    for (uint32_t m = todo; m; m &= m - 1) list[r++] = (uint16_t)((lane << 5) | (__ffs(m) - 1));
                    __syncwarp();  // the list, and the owners' count words, are visible to the whole warp
                    for (uint32_t item = lane; item < total; item += 32) {
                        const uint32_t e = list[item], o = e >> 5, c = e & 31u;
                        const uint64_t gi = (uint64_t)rc.x * TS + 16 * (warp >> 1) + (o >> 3) + 4 * (c / DJ);
                        const uint64_t gj = (uint64_t)rc.y * TS + 64 * (warp & 1) + (o & 7) + 8 * (c % DJ);
                        const uint32_t *words = my_words - lane + o + (size_t)(c / PACK) * 256;
                        put((int)((o >> 3) + 4 * (c / DJ)), (int)((o & 7) + 8 * (c % DJ)), row_base_of(gi) + gj,
                            pair_epilogue_k<CountT>(words, kstride, 16 * (c % PACK), E, gi, gj, rect));
                    }
                    __syncwarp();  // the list is reused by the next tile
                }
            }
            if (!rect) {  // write the warp's 16 x 64 block out, one row per step, coalesced
                __syncwarp();
                const uint64_t gj0 = (uint64_t)rc.y * TS + 64 * (warp & 1);
#pragma unroll 1
                for (int lrow = 0; lrow < 16; ++lrow) {
                    const uint64_t gi = (uint64_t)rc.x * TS + 16 * (warp >> 1) + lrow;
                    if (gi < S.row_begin || gi >= S.row_end) continue;
                    const uint64_t row_base = row_base_of(gi);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int lcol = h * 32 + lane;
                        const uint64_t gj = gj0 + lcol;
                        if (gj >= col_end || (S.mode == MODE_DENSE_SELF && gj <= gi)) continue;
                        if (two)
                            __stcs(reinterpret_cast<float2 *>(S.out) + row_base + gj, wst2[lrow * 64 + lcol]);
                        else
                            __stcs(S.out + row_base + gj, wst1[lrow * 64 + lcol]);
                    }
                }
                __syncwarp();  // the staging block is reused by the next tile
            }
        }
        dead = dead_init;
        if (S.work_stats && lane == 0) atomicAdd(S.work_stats + 1, (unsigned long long)n_it);  // steps of this warp and tile
    }
}

// Finishes the pairs dense_kernel deferred: one warp per pair.  The match counts of the first three k-mer lengths come
// with the list entry; the others are counted from the sample-major copies of the planes (untile_kernel): the two
// half-warps take two consecutive k-mer lengths, a lane compares the 14 planes of a 64-bin word (2 x 112 contiguous
// bytes) -- the reference's own loop (jaccard.rs:15-25) spread over 16 lanes.  Then the arithmetic of pair_epilogue
// (jaccard.rs:61-142) in the reference's order, and the result goes to its place in the dense or rect output.
// Runs after dense_kernel on the same stream; the entry count comes from the scratch header.
__global__ void __launch_bounds__(256) defer_kernel(const uint64_t *__restrict__ a_rows, const uint64_t *__restrict__ b_rows,
                                                    uint32_t W, const __grid_constant__ EpiParams E,
                                                    const uint4 *__restrict__ list,
                                                    const unsigned long long *__restrict__ count, uint32_t cap,
                                                    float2 *__restrict__ out) {
    const int lane = threadIdx.x & 31, half = lane >> 4, hl = lane & 15;
    const unsigned long long n = min(*count, (unsigned long long)cap);
    const bool comp = E.has_c != 0;
    const uint64_t kmer_stride = (uint64_t)W * BBITS, sample_stride = (uint64_t)E.nk * kmer_stride;
    for (unsigned long long p = (unsigned long long)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32; p < n;
         p += (unsigned long long)gridDim.x * (blockDim.x / 32)) {
        const uint4 e = list[2 * p], e2 = list[2 * p + 1];
        if (e.x == 0xffffffffu) continue;
        const uint64_t gi = e.x, gj = e.y, idx = (uint64_t)e.z | ((uint64_t)e.w << 32);
        double c1 = 0.0, c2 = 0.0;
        if (comp) {
            c1 = E.a_comp[gi];
            c2 = E.b_comp[gj];
        }
        double ysum = 0.0, xysum = 0.0, ysq = 0.0;
        int cnt = 0;
        bool alive = true;
        auto take = [&](int k, uint32_t c) {  // the reference's loop body: stop at the first k without a usable Jaccard
            if (!alive) return;
            const double y = comp ? log(corrected_jaccard(c, c1, c2, E)) : E.lny[c];
            if (y < E.tolerance) {
                alive = false;
                return;
            }
            ysum = __dadd_rn(ysum, y);
            xysum = __dadd_rn(xysum, __dmul_rn(E.kf[k], y));
            ysq = __dadd_rn(ysq, __dmul_rn(y, y));
            ++cnt;
        };
        take(0, e2.x);
        take(1, e2.y);
        take(2, e2.z);
        const ulonglong2 *pa = reinterpret_cast<const ulonglong2 *>(a_rows + gi * sample_stride);
        const ulonglong2 *pb = reinterpret_cast<const ulonglong2 *>(b_rows + gj * sample_stride);
        for (int k0 = 3; k0 < E.nk && alive; k0 += 2) {
            const int k = k0 + half;
            uint32_t mism = 0;
            if (k < E.nk) {
                for (uint32_t w = hl; w < W; w += 16) {
                    const size_t o = ((size_t)k * kmer_stride + (size_t)w * BBITS) / 2;
                    unsigned long long d = 0;
#pragma unroll
                    for (int t = 0; t < BBITS / 2; ++t) {
                        const ulonglong2 x = __ldg(pa + o + t), y = __ldg(pb + o + t);
                        d |= (x.x ^ y.x) | (x.y ^ y.y);
                    }
                    mism += __popcll(d);
                }
            }
            for (int o = 8; o > 0; o >>= 1) mism += __shfl_xor_sync(0xffffffffu, mism, o);  // within each half-warp
            const uint32_t m0 = __shfl_sync(0xffffffffu, mism, 0), m1 = __shfl_sync(0xffffffffu, mism, 16);
            take(k0, E.nbins - m0);
            if (k0 + 1 < E.nk) take(k0 + 1, E.nbins - m1);
        }
        if (lane == 0) {
            float d0, d1;
            regress(ysum, xysum, ysq, cnt, E, d0, d1);
            out[idx] = make_float2(d0, d1);
        }
    }
}

size_t dense_kernel_smem() {
    return (size_t)DNST * DSTAGE_U32 * 4 + 2 * DNST * 8 + TRING * sizeof(uint2) + 16 + 2 * TRING * 4 + 8 * 1024 * sizeof(uint16_t) +
           8 * 16 * 64 * sizeof(float2);
}
// per-CTA count scratch (one CTA per SM, 64 x 128 pairs each, nk <= kp words of 4 bytes per pair slot) + 256 bytes
// in front for the tile counter
size_t dense_kernel_scratch_bytes(int kp, bool count32, int sm_count) {
    return 256 + (size_t)sm_count * 2 * TS * TS * kp * (count32 ? 4 : 2);
}

int counts_per_pair(int nk, bool count32) {
    const int per = count32 ? 4 : 8;
    return (nk + per - 1) / per * per;
}

size_t tile_kernel_smem(int nk, bool count32, bool knn_mode, uint32_t cap, bool with_aux) {
    size_t off = (size_t)NST * 2 * STAGE_BYTES + (size_t)counts_per_pair(nk, count32) * TS * TS * (count32 ? 4 : 2);
    off = (off + 15) & ~(size_t)15;
    off += 64;
    (void)cap;
    (void)with_aux;
    if (knn_mode) off += TS * 8 + TS * 4;  // thresholds + fill counts; the lists themselves are in global memory
    return off;
}


// ------------------------------------------------------------------------------------------------
// Few queries (<= 8) against many references: HBM-bound (second roofline).
// The tile kernel stages both sides through shared memory and needs a CTA barrier per 64-bin word; with so few
// columns the work per word is too small to hide that, and it stalls at about half of the HBM rate.  Here the
// whole query side lives in shared memory (nq x K x W x 112 bytes, read with broadcast LDS), every thread owns ONE
// reference sample and streams its bit planes straight from global memory (a warp reads the quads of 32
// consecutive samples: 512 contiguous bytes per request, one word ahead in registers), and nothing in the main
// loop synchronises.  Counts go to a per-warp slice of shared memory in the [pair][kp] layout pair_epilogue reads.
constexpr int STREAM_THREADS = 512;

struct RefWord {  // the 14 planes of both 32-bin halves of one 64-bin word of one sample
    uint4 q[2][3];
    uint2 t[2];
};

// GC (9..16 queries in ONE pass over the references instead of two passes of 8): the query side alone fills shared
// memory (16 x 14 336 B at 1024 bins x 8 k), so the match counts go k-major to the L2-resident count scratch of the
// dense kernel ([warp][k][query pair][lane], two 16-bit counts per word, 128 contiguous bytes per warp store) and the
// epilogue reads them back from there (pair_epilogue_k).
template <int NQ, bool GC>
__global__ void __launch_bounds__(STREAM_THREADS, 1)
    stream_kernel(const __grid_constant__ DevDb A, const __grid_constant__ DevDb B, const __grid_constant__ EpiParams E,
                  const __grid_constant__ SchedParams S, const uint32_t q0) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t W = A.W, n_words = (uint32_t)E.nk * W;
    uint4 *qquad = reinterpret_cast<uint4 *>(smem_raw);                   // [word][half][quad][NQ]
    uint2 *qtail = reinterpret_cast<uint2 *>(qquad + (size_t)n_words * 6 * NQ);  // [word][half][NQ]
    uint16_t *counts_all = reinterpret_cast<uint16_t *>(qtail + (size_t)n_words * 2 * NQ);
    static_assert(!GC || NQ % 2 == 0, "count words pack two queries");
    constexpr size_t GKSTRIDE = (size_t)(NQ / 2) * 32;  // GC: words between consecutive k of one warp

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t nq_all = (uint32_t)B.n;                      // this launch: queries [q0, q0 + nq)
    const uint32_t nq = nq_all - q0 < (uint32_t)NQ ? nq_all - q0 : (uint32_t)NQ;
    const uint32_t hb0 = (uint32_t)(S.row_begin / 32), hb1 = (uint32_t)((S.row_end + 31) / 32);
    const uint32_t n_warps = gridDim.x * (STREAM_THREADS / 32);
    const uint32_t first = hb0 + blockIdx.x * (STREAM_THREADS / 32) + warp;

    for (uint32_t e = tid; e < n_words * 2 * NQ; e += STREAM_THREADS) {
        const uint32_t q = e % NQ, step = e / NQ, h = step & 1, kw = step >> 1;
        const uint32_t *slab = B.tiled + slab_index(B, (uint32_t)E.k0 + kw / W, kw % W, 0) * STAGE_U32 + h * HALF_U32;
        const bool live = q < nq;
#pragma unroll
        for (int quad = 0; quad < 3; ++quad)
            qquad[((size_t)step * 3 + quad) * NQ + q] =
                live ? *reinterpret_cast<const uint4 *>(slab + quad * QUAD_U32 + (q0 + q) * 4) : make_uint4(0, 0, 0, 0);
        qtail[(size_t)step * NQ + q] =
            live ? *reinterpret_cast<const uint2 *>(slab + TAIL_U32 + (q0 + q) * 2) : make_uint2(0, 0);
    }
    uint16_t *counts = counts_all + (size_t)warp * 32 * NQ * E.kp;
    uint32_t *gwords = nullptr;  // GC: this lane's count words
    if (GC) {
        gwords = reinterpret_cast<uint32_t *>(reinterpret_cast<unsigned char *>(S.counts_scratch) + 256) +
                 ((size_t)blockIdx.x * (STREAM_THREADS / 32) + warp) * (size_t)E.nk * GKSTRIDE + lane;
    } else if (E.kp != E.nk) {  // padding slots of the per-pair count words must hold a valid table index
        for (int e = lane; e < 32 * NQ * E.kp; e += 32) counts[e] = 0;
    }
    __syncthreads();

    const uint32_t nbins = 64u * W;
    const uint64_t word_stride = (uint64_t)A.nb * STAGE_U32;  // u32 between consecutive (k, w) slabs of one block
    for (uint32_t hb = first; hb < hb1; hb += n_warps) {
        const uint32_t b = hb >> 1, s = (hb & 1) * 32 + lane;
        const uint64_t gi = (uint64_t)b * TS + s;
        const uint32_t *p = A.tiled + slab_index(A, (uint32_t)E.k0, 0, b) * STAGE_U32 + s * 4;
        // the tail pair of sample s sits at s * 2, not s * 4: fold the difference into the tail offset
        const int tail_fix = -(int)(s * 2);
        RefWord cur, nxt;
        auto load = [&](RefWord &r, const uint32_t *base) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#pragma unroll
                for (int q = 0; q < 3; ++q)
                    r.q[h][q] = __ldcs(reinterpret_cast<const uint4 *>(base + h * HALF_U32 + q * QUAD_U32));
                r.t[h] = __ldcs(reinterpret_cast<const uint2 *>(base + h * HALF_U32 + TAIL_U32 + tail_fix));
            }
        };
        load(nxt, p);
        uint32_t mism[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) mism[q] = 0;
        uint32_t w_in_k = 0, kk = 0;
        for (uint32_t kw = 0; kw < n_words; ++kw) {
            cur = nxt;
            p += word_stride;
            if (kw + 1 < n_words) load(nxt, p);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint4 *qq = qquad + ((size_t)kw * 2 + h) * 3 * NQ;
                const uint2 *qt = qtail + ((size_t)kw * 2 + h) * NQ;
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    const uint4 a = qq[q], bq = qq[NQ + q], c = qq[2 * NQ + q];
                    const uint2 t = qt[q];
                    uint32_t d = cur.q[h][0].x ^ a.x;
                    d |= cur.q[h][0].y ^ a.y;
                    d |= cur.q[h][0].z ^ a.z;
                    d |= cur.q[h][0].w ^ a.w;
                    d |= cur.q[h][1].x ^ bq.x;
                    d |= cur.q[h][1].y ^ bq.y;
                    d |= cur.q[h][1].z ^ bq.z;
                    d |= cur.q[h][1].w ^ bq.w;
                    d |= cur.q[h][2].x ^ c.x;
                    d |= cur.q[h][2].y ^ c.y;
                    d |= cur.q[h][2].z ^ c.z;
                    d |= cur.q[h][2].w ^ c.w;
                    d |= cur.t[h].x ^ t.x;
                    d |= cur.t[h].y ^ t.y;
                    mism[q] += __popc(d);
                }
            }
            if (++w_in_k == W) {
                if (GC) {
#pragma unroll
                    for (int q = 0; q < NQ; q += 2)
                        __stcg(gwords + (size_t)kk * GKSTRIDE + (q / 2) * 32, (nbins - mism[q]) | ((nbins - mism[q + 1]) << 16));
                }
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    if (!GC) counts[(lane * NQ + q) * E.kp + kk] = (uint16_t)(nbins - mism[q]);
                    mism[q] = 0;
                }
                w_in_k = 0;
                ++kk;
            }
        }
        __syncwarp();
        if (gi >= S.row_begin && gi < S.row_end) {
            const bool rect = S.rect_ld != 0;
            for (uint32_t q = 0; q < nq; ++q) {
                const int pair = lane * NQ + (int)q;
                const uint64_t gj = q0 + q;
                if (S.mode == MODE_COUNTS) {
                    uint32_t *o = reinterpret_cast<uint32_t *>(S.out) + ((gi - S.row_begin) * nq_all + gj) * (uint64_t)E.nk;
                    for (int k = 0; k < E.nk; ++k)
                        o[k] = GC ? (__ldcg(gwords + (size_t)k * GKSTRIDE + (q / 2) * 32) >> (16 * (q & 1))) & 0xffffu
                                  : counts[pair * E.kp + k];
                    continue;
                }
                const uint64_t idx = (gi - S.row_begin) * (rect ? S.rect_ld : (uint64_t)nq_all) + gj;
                float d0, d1;
                if (GC) {
                    const uint32_t *words = gwords + (q / 2) * 32;
                    if (E.kind != KIND_CORE_ACC && !E.has_c) {
                        const uint32_t cnt = (__ldcg(words) >> (16 * (q & 1))) & 0xffffu;
                        d0 = __ldg(&(rect ? E.key_tab : E.val_tab)[cnt]);
                        d1 = 0.0f;
                    } else {
                        const float2 d = pair_epilogue_k<uint16_t>(words, GKSTRIDE, 16 * (int)(q & 1), E, gi, gj, rect);
                        d0 = d.x;
                        d1 = d.y;
                    }
                } else if (E.has_c)
                    pair_epilogue<uint16_t, true>(counts, pair, E, gi, gj, rect, d0, d1);
                else
                    pair_epilogue<uint16_t, false>(counts, pair, E, gi, gj, rect, d0, d1);
                if (rect) {  // kNN keys, read back soon by select_few_kernel
                    if (E.kind == KIND_CORE_ACC)
                        reinterpret_cast<float2 *>(S.out)[idx] = make_float2(d0, d1);
                    else
                        S.out[idx] = d0;
                } else if (E.kind == KIND_CORE_ACC) {
                    __stcs(reinterpret_cast<float2 *>(S.out) + idx, make_float2(d0, d1));
                } else {
                    __stcs(S.out + idx, d0);
                }
            }
        }
        __syncwarp();
    }
}

// Shared memory of stream_kernel<NQ>; 0 if the query side does not fit.  gc: counts in the L2 scratch instead.
static size_t stream_smem(const DevDb &A, const EpiParams &E, int NQ, bool gc = false) {
    const size_t n_words = (size_t)E.nk * A.W;
    const size_t bytes = n_words * 2 * NQ * (3 * 16 + 8) + (gc ? 0 : (size_t)(STREAM_THREADS / 32) * 32 * NQ * E.kp * 2);
    return bytes <= 232448 ? bytes : 0;
}
// 9..16 queries in one pass: needs the count scratch (dense_kernel_scratch_bytes covers 16 warps x 32 x 16 pairs per SM)
static bool stream_one_pass_16(const DevDb &A, const DevDb &B, const EpiParams &E, const SchedParams &S) {
    return B.n > 8 && B.n <= 16 && S.counts_scratch != nullptr && stream_smem(A, E, 16, true) != 0;
}

static bool stream_applies(const DevDb &A, const DevDb &B, const EpiParams &E, const SchedParams &S, bool count32) {
    // up to 16 queries: one or two passes over the references still beat the narrow tile variants (B200: 16 queries
    // 2.67 ms vs 2.86 ms per 7.2 GB of reference planes; 32 queries 5.3 ms vs 4.5 ms)
    if (count32 || B.n == 0 || B.n > 16 || A.W != B.W) return false;
    if (S.mode != MODE_DENSE_CROSS && S.mode != MODE_COUNTS) return false;
    if (S.mode == MODE_COUNTS && S.self_mode) return false;
    const int NQ = B.n <= 1 ? 1 : B.n <= 2 ? 2 : B.n <= 4 ? 4 : 8;  // more than 8 queries: passes of 8
    return stream_smem(A, E, NQ) != 0;
}

static cudaError_t launch_stream(const DevDb &A, const DevDb &B, const EpiParams &E, const SchedParams &S, cudaStream_t st) {
    int dev = 0, sm_count = 148;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sm_count <= 0)
        sm_count = 148;
    const int NQ = B.n <= 1 ? 1 : B.n <= 2 ? 2 : B.n <= 4 ? 4 : 8;
    const bool one16 = stream_one_pass_16(A, B, E, S);
    const size_t smem = one16 ? stream_smem(A, E, 16, true) : stream_smem(A, E, NQ);
    const uint64_t units = (S.row_end + 31) / 32 - S.row_begin / 32;
    if (units == 0) return cudaSuccess;
    const unsigned ctas = (unsigned)std::min<uint64_t>((uint64_t)sm_count, (units + STREAM_THREADS / 32 - 1) / (STREAM_THREADS / 32));
    cudaError_t err = cudaSuccess;
    if (one16) {
        // (a variant with two references per thread, to halve the broadcast LDS per LOP3, measured 9 % slower: it has
        // no registers left to keep the next word's global loads in flight)
        err = cudaFuncSetAttribute(stream_kernel<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        stream_kernel<16, true><<<ctas, STREAM_THREADS, smem, st>>>(A, B, E, S, 0);
        return cudaGetLastError();
    }
#define SKD_STREAM(N)                                                                                          \
    do {                                                                                                       \
        err = cudaFuncSetAttribute((stream_kernel<N, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (err != cudaSuccess) return err;                                                                    \
        for (uint32_t q0 = 0; q0 < B.n; q0 += N) stream_kernel<N, false><<<ctas, STREAM_THREADS, smem, st>>>(A, B, E, S, q0); \
    } while (0)
    if (NQ == 1) SKD_STREAM(1);
    else if (NQ == 2) SKD_STREAM(2);
    else if (NQ == 4) SKD_STREAM(4);
    else SKD_STREAM(8);
#undef SKD_STREAM
    return cudaGetLastError();
}

cudaError_t launch_tiles(const DevDb &A, const DevDb &B, const EpiParams &E, const SchedParams &S, uint64_t n_ctas,
                         bool count32, size_t smem, cudaStream_t st) {
    if (n_ctas == 0) return cudaSuccess;
    if (n_ctas > 0x7fffffffull) return cudaErrorInvalidConfiguration;
    if (stream_applies(A, B, E, S, count32)) return launch_stream(A, B, E, S, st);
    // Narrow-column variants: when the column side has at most 8 / 16 / 32 samples (a handful of queries against a
    // large reference database) only those columns are loaded and compared.
    const int cols = S.mode == MODE_KNN || TJ != 4 ? TS : (B.n <= 8 ? 8 : B.n <= 16 ? 16 : B.n <= 32 ? 32 : TS);
    cudaError_t err = cudaSuccess;
    // the caller chose dense_kernel (count scratch given, tile groups built over column block pairs)
    if (TJ == 4 && S.dense_pairs && S.counts_scratch && (S.mode == MODE_DENSE_SELF || S.mode == MODE_DENSE_CROSS)) {
        // the barrier-free persistent kernel: 2 CTAs per SM, each walking up to MAXT tiles per launch
        int dev = 0, sm_count = 148;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sm_count <= 0)
            sm_count = 148;
        const uint64_t max_ctas = (uint64_t)sm_count;  // one CTA per SM
        const size_t dsmem = dense_kernel_smem();
        // the early-exit instantiation only where it can act: the plain one carries none of its code
        const bool defer = S.defer_cap != 0 && E.kind == KIND_CORE_ACC && E.nk > 3;
        const void *kern = count32 ? (defer ? (const void *)dense_kernel<uint32_t, true> : (const void *)dense_kernel<uint32_t, false>)
                                   : (defer ? (const void *)dense_kernel<uint16_t, true> : (const void *)dense_kernel<uint16_t, false>);
        err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsmem);
        if (err != cudaSuccess) return err;
        // the tile counter sits in front of the count scratch and is reset (stream-ordered) before every launch
        unsigned long long *counter = reinterpret_cast<unsigned long long *>(S.counts_scratch);
        unsigned char *counts = reinterpret_cast<unsigned char *>(S.counts_scratch) + 256;
        err = cudaMemsetAsync(counter, 0, 16, st);  // tile counter + fill count of the deferred-pair list
        if (err != cudaSuccess) return err;
        const unsigned grid = (unsigned)std::min<uint64_t>(n_ctas, max_ctas);
        void *cptr = counts;
        void *args[] = {(void *)&A, (void *)&B, (void *)&E, (void *)&S, (void *)&n_ctas, (void *)&counter, (void *)&cptr};
        err = cudaLaunchKernel(kern, dim3(grid), dim3(DENSE_THREADS), args, dsmem, st);
        if (err != cudaSuccess) return err;
        if (defer)
            defer_kernel<<<8 * sm_count, 256, 0, st>>>(S.a_rows, S.b_rows, A.W, E, S.defer_list, counter + 1, S.defer_cap,
                                                      reinterpret_cast<float2 *>(S.out));
        return cudaGetLastError();
    }
#define SKD_LAUNCH(T, C)                                                                                      \
    do {                                                                                                      \
        err = cudaFuncSetAttribute(tile_kernel<T, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (err != cudaSuccess) return err;                                                                   \
        tile_kernel<T, C><<<(unsigned)n_ctas, THREADS, smem, st>>>(A, B, E, S);                                \
    } while (0)
#define SKD_LAUNCH_COLS(T)                          \
    do {                                            \
        if (cols == 8) SKD_LAUNCH(T, 8);            \
        else if (cols == 16) SKD_LAUNCH(T, 16);     \
        else if (cols == 32) SKD_LAUNCH(T, 32);     \
        else SKD_LAUNCH(T, TS);                     \
    } while (0)
    if (count32) SKD_LAUNCH_COLS(uint32_t); else SKD_LAUNCH_COLS(uint16_t);
#undef SKD_LAUNCH_COLS
#undef SKD_LAUNCH
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// kNN merge: one warp per row merges n_split ascending lists of knn composites and writes the final
// sparse row (neighbour index, first distance, second distance).  ANI keys are un-transformed as
// 1.0f - key like the reference (src/distances/mod.rs:183-189).
__global__ void knn_merge_kernel(const unsigned long long *__restrict__ part_key, const float *__restrict__ part_aux,
                                 uint32_t n_split, uint32_t knn, uint32_t cap, uint64_t first_lrow, uint64_t n_rows,
                                 int kind,
                                 unsigned long long *__restrict__ idx_out, float *__restrict__ d0_out,
                                 float *__restrict__ d1_out, unsigned long long *__restrict__ key_out,
                                 float *__restrict__ aux_out) {
    const uint64_t row = (uint64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x & 31;
    if (row >= n_rows) return;
    const uint64_t lrow = first_lrow + row;
    // list of (row block rbl, split s, row r) starts at ((rbl * n_split + s) * TS + r) * cap
    const uint64_t rbl = lrow / TS, r = lrow % TS;
    auto list = [&](uint32_t s) { return ((rbl * n_split + s) * TS + r) * (uint64_t)cap; };
    const uint32_t total = n_split * knn;
    // rank of element e = number of elements smaller than it (composites are unique per row except padding)
    for (uint32_t e = lane; e < total; e += 32) {
        const uint64_t src = list(e / knn) + e % knn;
        unsigned long long v = part_key[src];
        if (v == ~0ull) continue;
        uint32_t rank = 0;
        for (uint32_t s = 0; s < n_split; ++s) {
            // binary search in list s for the number of entries < v
            const unsigned long long *l = part_key + list(s);
            uint32_t lo = 0, hi = knn;
            while (lo < hi) {
                uint32_t mid = (lo + hi) >> 1;
                if (l[mid] < v) lo = mid + 1; else hi = mid;
            }
            rank += lo;
        }
        if (rank < knn) {
            if (key_out) {  // partial result for a later cross-rank merge: keep the composite
                key_out[row * knn + rank] = v;
                if (aux_out) aux_out[row * knn + rank] = part_aux ? part_aux[src] : 0.0f;
            } else {
                float key = float_from_order_bits((uint32_t)(v >> 32));
                idx_out[row * knn + rank] = (unsigned long long)(uint32_t)v;
                d0_out[row * knn + rank] = kind == KIND_ANI ? 1.0f - key : key;
                if (d1_out) d1_out[row * knn + rank] = part_aux ? part_aux[src] : 0.0f;
            }
        }
    }
}

void launch_knn_merge(const unsigned long long *part_key, const float *part_aux, uint32_t n_split, uint32_t knn,
                      uint32_t cap, uint64_t first_lrow, uint64_t n_rows, int kind, unsigned long long *idx_out, float *d0_out,
                      float *d1_out, unsigned long long *key_out, float *aux_out, cudaStream_t st) {
    if (n_rows == 0) return;
    const int warps = 8;
    knn_merge_kernel<<<(unsigned)((n_rows + warps - 1) / warps), warps * 32, 0, st>>>(
        part_key, part_aux, n_split, knn, cap, first_lrow, n_rows, kind, idx_out, d0_out, d1_out, key_out, aux_out);
}

// ------------------------------------------------------------------------------------------------
// Symmetric self-kNN, second stage.  The tile kernel (rect mode) has written the kNN keys of the pairs
// (i, j > i) of a row group as a row-major matrix D in HBM; every such value is a candidate for row i
// (read along rows) and for row j (read along columns), so each unordered pair is computed once.
// select_sym_kernel folds D into the running per-sample lists L[sample][knn] (sorted u64 composites, ~0 padded);
// select_few_kernel does the same for ref-major cross matrices with few query columns.  Both keep 128-slot lists
// in shared memory with threshold rejection, exactly like the tile kernel's kNN epilogue.  (A first version streamed
// 64x64 sub-tiles of D through shared memory with one warp per target and two CTA barriers per sub-tile: 2.7x slower.)

// Cross-mode selection over few target columns (queries).  D is a ref-major matrix of kNN keys (rows = references
// g0 .., columns = the nq queries; float2 key/aux with AUX; ld >= nq): every row is a candidate of every column and
// the stored neighbour index is row + idx_offset.  A CTA takes a group of 8 targets (blockIdx.y) and one slice of
// rows_per_split rows (blockIdx.x; one list set per slice at Lkey + slice * part_stride, carried over from the same
// slice of earlier reference chunks, so thresholds are tight from the second chunk on); the rows of the slice are
// divided over the 8 warps, warp w keeps one list per target (list w * 8 + q), lanes take 32 consecutive rows at a
// time, and the 8 lists of a target are merged at the end.  (Giving each target to one warp leaves one warp per CTA
// busy when there is a single query: 2.3 ms instead of 0.3 ms per 500 000 references.)
template <bool AUX>
__global__ void __launch_bounds__(256) select_few_kernel(const float *__restrict__ D, uint64_t ld, uint64_t g0, uint32_t R,
                                                         uint32_t nq, unsigned long long *__restrict__ Lkey,
                                                         float *__restrict__ Laux, uint32_t knn, uint64_t idx_offset,
                                                         uint32_t rows_per_split, uint64_t part_stride) {
    constexpr int CAP = 128;
    extern __shared__ __align__(16) unsigned char sm_raw[];
    unsigned long long *buf = reinterpret_cast<unsigned long long *>(sm_raw);  // [8 warps][8 targets][CAP]
    float *baux = reinterpret_cast<float *>(buf + TS * CAP);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint64_t lo = (uint64_t)blockIdx.x * rows_per_split;
    const uint64_t hi = lo + rows_per_split < R ? lo + rows_per_split : R;
    const uint32_t t0 = blockIdx.y * 8;  // this CTA's group of (at most) 8 targets
    Lkey += (uint64_t)blockIdx.x * part_stride + (uint64_t)t0 * knn;
    if (AUX) Laux += (uint64_t)blockIdx.x * part_stride + (uint64_t)t0 * knn;
    D += (AUX ? 2 : 1) * (size_t)t0;
    nq = nq - t0 < 8 ? nq - t0 : 8;
    for (int e = lane; e < 8 * CAP; e += 32) buf[warp * 8 * CAP + e] = ~0ull;
    __syncwarp();

    const uint64_t span = hi > lo ? hi - lo : 0;
    const uint64_t per_warp = ((span + 7) / 8 + 31) / 32 * 32;
    const uint64_t w_lo = lo + (uint64_t)warp * per_warp;
    const uint64_t w_hi = w_lo + per_warp < hi ? w_lo + per_warp : hi;
    uint32_t cnt[8];
    unsigned long long thr[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        cnt[q] = 0;
        // the slice's list from earlier reference chunks (~0 padded): nothing at or above its knn-th key can enter
        thr[q] = q < (int)nq ? min(Lkey[(uint64_t)q * knn + knn - 1], useless_from(knn)) : 0ull;
    }
    for (uint64_t r0 = w_lo; r0 < w_hi; r0 += 32) {
        const uint64_t row = r0 + lane;
        const bool in = row < w_hi;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (q >= (int)nq) break;
            unsigned long long *list = buf + (warp * 8 + q) * CAP;
            float key = 0.f, a = 0.f;
            if (in) {
                if (AUX) {
                    const float2 v = __ldcs(reinterpret_cast<const float2 *>(D) + row * ld + q);
                    key = v.x;
                    a = v.y;
                } else {
                    key = __ldcs(D + row * ld + q);
                }
            }
            const unsigned long long comp =
                ((unsigned long long)float_order_bits(key) << 32) | (uint32_t)(g0 + row + idx_offset);
            const bool ok = in && comp < thr[q];
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            if (m == 0) continue;
            if (ok) {
                const uint32_t slot = cnt[q] + __popc(m & ((1u << lane) - 1u));
                list[slot] = comp;
                if (AUX) baux[(warp * 8 + q) * CAP + slot] = a;
            }
            cnt[q] += __popc(m);
            __syncwarp();
            if (cnt[q] + 32 > CAP) {
                warp_bitonic_sort(list, AUX ? baux + (warp * 8 + q) * CAP : nullptr, CAP, lane);
                for (int e = knn + lane; e < CAP; e += 32) list[e] = ~0ull;
                __syncwarp();
                cnt[q] = knn;
                thr[q] = list[knn - 1] < thr[q] ? list[knn - 1] : thr[q];
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q)
        if (q < (int)nq && cnt[q] > 0)
            warp_bitonic_sort(buf + (warp * 8 + q) * CAP, AUX ? baux + (warp * 8 + q) * CAP : nullptr, CAP, lane);
    __syncthreads();
    // warp q merges the stored list of target q and the 8 warp lists (each sorted, best knn <= 64 first)
    if (warp < (int)nq) {
        const int q = warp;
        unsigned long long *acc = buf + q * CAP;  // warp 0's list of this target doubles as the accumulator
        float *acc_a = AUX ? baux + q * CAP : nullptr;
        bool any = false;
        for (int w = 0; w < 8; ++w) any |= buf[(w * 8 + q) * CAP] != ~0ull;
        if (!any) return;  // no candidate of this chunk beat the threshold: the stored list stands
        for (int w = 0; w < 8; ++w) {
            const unsigned long long *src = w == 0 ? Lkey + (uint64_t)q * knn : buf + (w * 8 + q) * CAP;
            const float *src_a = !AUX ? nullptr : (w == 0 ? Laux + (uint64_t)q * knn : baux + (w * 8 + q) * CAP);
            if (src[0] == ~0ull) continue;  // empty
            for (int e = lane; e < CAP - (int)knn; e += 32) {
                acc[knn + e] = e < (int)knn ? src[e] : ~0ull;
                if (AUX) acc_a[knn + e] = e < (int)knn ? src_a[e] : 0.f;
            }
            __syncwarp();
            warp_bitonic_sort(acc, acc_a, CAP, lane);
        }
        for (uint32_t e = lane; e < knn; e += 32) {
            Lkey[(uint64_t)q * knn + e] = acc[e];
            if (AUX) Laux[(uint64_t)q * knn + e] = acc_a[e];
        }
    }
}

cudaError_t launch_select_few(bool aux, const float *D, uint64_t ld, uint64_t g0, uint32_t R, uint32_t nq,
                              unsigned long long *Lkey, float *Laux, uint32_t knn, cudaStream_t st, uint64_t idx_offset,
                              uint32_t n_splits, uint32_t rows_per_split, uint64_t part_stride) {
    const size_t smem = (size_t)TS * 128 * 8 + (aux ? (size_t)TS * 128 * 4 : 0);
    const dim3 grid(n_splits, (nq + 7) / 8);
    cudaError_t err;
    if (aux) {
        err = cudaFuncSetAttribute(select_few_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        select_few_kernel<true><<<grid, 256, smem, st>>>(D, ld, g0, R, nq, Lkey, Laux, knn, idx_offset, rows_per_split,
                                                             part_stride);
    } else {
        err = cudaFuncSetAttribute(select_few_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        select_few_kernel<false><<<grid, 256, smem, st>>>(D, ld, g0, R, nq, Lkey, Laux, knn, idx_offset, rows_per_split,
                                                              part_stride);
    }
    return cudaGetLastError();
}

// Symmetric self-kNN, second stage: a CTA takes 8 targets, its 8
// warps take contiguous slices of the candidate range, every warp keeps one 128-slot list per target, and the 8 lists
// of a target are merged with the list accumulated so far (Lkey / Laux, [n][knn]) at the end.
//   ROWS: targets are rows i of the group [t0, t1) = [g0, g1), candidates are the columns j in (i, n): lanes read 32
//         consecutive keys of row i.
//   COLS: targets are columns j in [t0, t1) = [g0, n), candidates are the rows i in [g0, min(j, g0 + R)): a lane reads
//         the keys of its row for the CTA's 8 consecutive columns (32 or 64 contiguous bytes).
// Every candidate at or above a target's current knn-th key is rejected up front by all warps.
template <bool COLS, bool AUX>
__global__ void __launch_bounds__(256) select_sym_kernel(const float *__restrict__ D, uint64_t ld, uint64_t g0, uint32_t R,
                                                         uint64_t n, uint64_t t0, uint64_t t1,
                                                         unsigned long long *__restrict__ Lkey, float *__restrict__ Laux,
                                                         uint32_t knn) {
    constexpr int CAP = 128;
    extern __shared__ __align__(16) unsigned char sm_raw[];
    unsigned long long *buf = reinterpret_cast<unsigned long long *>(sm_raw);  // [8 warps][8 targets][CAP]
    unsigned long long *gbuf = buf + TS * CAP;                                 // [8 targets][CAP]: gather buffer of the merge
    float *baux = reinterpret_cast<float *>(gbuf + 8 * CAP);
    float *gaux = baux + TS * CAP;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint64_t tbase = t0 + (uint64_t)blockIdx.x * 8;
    const uint32_t nt = (uint32_t)(t1 - tbase < 8 ? t1 - tbase : 8);

    // candidate range of the CTA (union over its targets), cut into 8 warp slices of whole 32-candidate steps
    const uint64_t s_begin = COLS ? g0 : tbase + 1;
    const uint64_t s_end = COLS ? (tbase + nt - 1 < g0 + R ? tbase + nt - 1 : g0 + R) : n;
    const uint64_t span = s_end > s_begin ? s_end - s_begin : 0;
    const uint64_t per_warp = ((span + 7) / 8 + 31) / 32 * 32;
    const uint64_t w_lo = s_begin + (uint64_t)warp * per_warp;
    const uint64_t w_hi = w_lo + per_warp < s_end ? w_lo + per_warp : s_end;

    uint32_t cnt[8];
    unsigned long long thr[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        unsigned long long *list = buf + (warp * 8 + q) * CAP;
        for (int e = lane; e < CAP; e += 32) list[e] = ~0ull;
        cnt[q] = 0;
        thr[q] = q < (int)nt ? min(Lkey[(tbase + q) * knn + knn - 1], useless_from(knn + 1)) : 0ull;
    }
    __syncwarp();

    // One step = 32 candidates (one per lane) against the CTA's 8 targets.  The keys of the next step are requested
    // before the current one is folded: the loop is bound by the latency of these loads (16 warps per SM, one
    // 64-byte request per lane and step), not by the folding.
    auto load_step = [&](uint64_t c0, float (&key)[8], float (&av)[8]) {
        const uint64_t cand = c0 + lane;
        const bool in = cand < w_hi;
        if (COLS) {  // one row, 8 consecutive columns
            const size_t off = (size_t)(cand - g0) * ld + (tbase - g0);
            if (in) {
                if (AUX) {
                    const float4 *p4 = reinterpret_cast<const float4 *>(reinterpret_cast<const float2 *>(D) + off);
#pragma unroll
                    for (int h = 0; h < 4; ++h) {
                        const float4 v = __ldcs(p4 + h);
                        key[2 * h] = v.x;
                        av[2 * h] = v.y;
                        key[2 * h + 1] = v.z;
                        av[2 * h + 1] = v.w;
                    }
                } else {
                    const float4 *p4 = reinterpret_cast<const float4 *>(D + off);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const float4 v = __ldcs(p4 + h);
                        key[4 * h] = v.x;
                        key[4 * h + 1] = v.y;
                        key[4 * h + 2] = v.z;
                        key[4 * h + 3] = v.w;
                    }
                }
            }
        } else {  // 8 rows, one column each
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (in && q < (int)nt && cand > tbase + q) {
                    const size_t off = (size_t)(tbase + q - g0) * ld + (cand - g0);
                    if (AUX) {
                        const float2 v = __ldcs(reinterpret_cast<const float2 *>(D) + off);
                        key[q] = v.x;
                        av[q] = v.y;
                    } else {
                        key[q] = __ldcs(D + off);
                    }
                }
            }
        }
    };
    auto fold_step = [&](uint64_t c0, const float (&key)[8], const float (&av)[8]) {
        const uint64_t cand = c0 + lane;
        const bool in = cand < w_hi;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (q >= (int)nt) break;
            const uint64_t tg = tbase + q;
            const bool valid = in && (COLS ? cand < tg : cand > tg);
            unsigned long long comp = 0;
            if (valid) comp = ((unsigned long long)float_order_bits(key[q]) << 32) | (uint32_t)cand;
            const bool ok = valid && comp < thr[q];
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            if (m == 0) continue;
            unsigned long long *list = buf + (warp * 8 + q) * CAP;
            float *la = AUX ? baux + (warp * 8 + q) * CAP : nullptr;
            if (ok) {
                const uint32_t slot = cnt[q] + __popc(m & ((1u << lane) - 1u));
                list[slot] = comp;
                if (AUX) la[slot] = av[q];
            }
            cnt[q] += __popc(m);
            __syncwarp();
            if (cnt[q] + 32 > CAP) {
                warp_bitonic_sort(list, la, CAP, lane);
                for (int e = knn + lane; e < CAP; e += 32) list[e] = ~0ull;
                __syncwarp();
                cnt[q] = knn;
                thr[q] = list[knn - 1] < thr[q] ? list[knn - 1] : thr[q];
            }
        }
    };
    {
        float keyA[8], avA[8], keyB[8], avB[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) keyA[q] = avA[q] = keyB[q] = avB[q] = 0.0f;
        if (w_lo < w_hi) load_step(w_lo, keyA, avA);
        for (uint64_t c0 = w_lo; c0 < w_hi; c0 += 64) {
            if (c0 + 32 < w_hi) load_step(c0 + 32, keyB, avB);
            fold_step(c0, keyA, avA);
            if (c0 + 64 < w_hi) load_step(c0 + 64, keyA, avA);
            if (c0 + 32 < w_hi) fold_step(c0 + 32, keyB, avB);
        }
    }
    // ---- merge.  Usually only a handful of candidates survived per target (everything at distance 1 beyond the
    // first knn indices is rejected up front): then the 8 warps' entries of a target are gathered unsorted behind the
    // stored list and ONE sort finishes it.  Otherwise every warp sorts its lists and they are merged one by one.
    __shared__ uint32_t wcnt[8][8];
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < 8; ++q) wcnt[warp][q] = cnt[q];
    }
    __syncthreads();
    uint32_t total[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        total[q] = 0;
        for (int w = 0; w < 8; ++w) total[q] += wcnt[w][q];
        if (q < (int)nt && cnt[q] > 0 && total[q] > CAP - knn)
            warp_bitonic_sort(buf + (warp * 8 + q) * CAP, AUX ? baux + (warp * 8 + q) * CAP : nullptr, CAP, lane);
    }
    __syncthreads();
    if (warp < (int)nt) {
        const int q = warp;
        const uint64_t tg = tbase + q;
        if (total[q] == 0) return;  // no candidate beat the threshold: the stored list stands
        if (total[q] <= CAP - knn) {
            unsigned long long *acc = gbuf + q * CAP;
            float *acc_a = AUX ? gaux + q * CAP : nullptr;
            for (int e = lane; e < CAP; e += 32) {
                acc[e] = e < (int)knn ? Lkey[tg * knn + e] : ~0ull;
                if (AUX) acc_a[e] = e < (int)knn ? Laux[tg * knn + e] : 0.f;
            }
            uint32_t at = knn;
            for (int w = 0; w < 8; ++w) {
                const uint32_t c = wcnt[w][q];
                for (uint32_t e = lane; e < c; e += 32) {
                    acc[at + e] = buf[(w * 8 + q) * CAP + e];
                    if (AUX) acc_a[at + e] = baux[(w * 8 + q) * CAP + e];
                }
                at += c;
            }
            __syncwarp();
            warp_bitonic_sort(acc, acc_a, CAP, lane);
            for (uint32_t e = lane; e < knn; e += 32) {
                Lkey[tg * knn + e] = acc[e];
                if (AUX) Laux[tg * knn + e] = acc_a[e];
            }
            return;
        }
        unsigned long long *acc = buf + q * CAP;  // warp 0's list of this target doubles as the accumulator
        float *acc_a = AUX ? baux + q * CAP : nullptr;
        for (int w = 0; w < 8; ++w) {
            // source: w == 0 -> the stored list; otherwise warp w's list
            const unsigned long long *src = w == 0 ? Lkey + tg * knn : buf + (w * 8 + q) * CAP;
            const float *src_a = !AUX ? nullptr : (w == 0 ? Laux + tg * knn : baux + (w * 8 + q) * CAP);
            if (src[0] == ~0ull) continue;  // empty
            for (int e = lane; e < CAP - (int)knn; e += 32) {
                acc[knn + e] = e < (int)knn ? src[e] : ~0ull;
                if (AUX) acc_a[knn + e] = e < (int)knn ? src_a[e] : 0.f;
            }
            __syncwarp();
            warp_bitonic_sort(acc, acc_a, CAP, lane);
        }
        for (uint32_t e = lane; e < knn; e += 32) {
            Lkey[tg * knn + e] = acc[e];
            if (AUX) Laux[tg * knn + e] = acc_a[e];
        }
    }
}

cudaError_t launch_select_sym(bool cols, bool aux, const float *D, uint64_t ld, uint64_t g0, uint32_t R, uint64_t n,
                              uint64_t t0, uint64_t t1, unsigned long long *Lkey, float *Laux, uint32_t knn,
                              cudaStream_t st) {
    if (t1 <= t0) return cudaSuccess;
    const unsigned grid = (unsigned)((t1 - t0 + 7) / 8);
    const size_t smem = (size_t)(TS + 8) * 128 * 8 + (aux ? (size_t)(TS + 8) * 128 * 4 : 0);
    cudaError_t err;
#define SKD_SEL(C, A)                                                                                                \
    do {                                                                                                             \
        err = cudaFuncSetAttribute(select_sym_kernel<C, A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (err != cudaSuccess) return err;                                                                          \
        select_sym_kernel<C, A><<<grid, 256, smem, st>>>(D, ld, g0, R, n, t0, t1, Lkey, Laux, knn);                   \
    } while (0)
    if (cols && aux) SKD_SEL(true, true);
    else if (cols) SKD_SEL(true, false);
    else if (aux) SKD_SEL(false, true);
    else SKD_SEL(false, false);
#undef SKD_SEL
    return cudaGetLastError();
}

// Merges n_parts sorted partial lists per row (keys[part][row][knn], ~0 padded) and either writes the final
// sparse rows (idx, d0, d1; ANI un-transformed as 1.0f - key like the reference, mod.rs:183-189) or, when
// key_out is given, one merged composite list per row.
__global__ void knn_finalize_kernel(const unsigned long long *__restrict__ keys, const float *__restrict__ aux,
                                    uint32_t n_parts, uint64_t part_stride, uint64_t row0, uint64_t n_rows, uint32_t knn,
                                    int kind, unsigned long long *__restrict__ idx_out, float *__restrict__ d0_out,
                                    float *__restrict__ d1_out, unsigned long long *__restrict__ key_out,
                                    float *__restrict__ aux_out) {
    const uint64_t row = (uint64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x & 31;
    if (row >= n_rows) return;
    const uint32_t total = n_parts * knn;
    for (uint32_t e = lane; e < total; e += 32) {
        const uint64_t src = (uint64_t)(e / knn) * part_stride + (row0 + row) * knn + e % knn;
        const unsigned long long v = keys[src];
        uint32_t rank = 0;
        if (v != ~0ull) {
            for (uint32_t s = 0; s < n_parts; ++s) {
                const unsigned long long *l = keys + (uint64_t)s * part_stride + (row0 + row) * knn;
                uint32_t lo = 0, hi = knn;
                while (lo < hi) {
                    uint32_t mid = (lo + hi) >> 1;
                    if (l[mid] < v) lo = mid + 1; else hi = mid;
                }
                rank += lo;
            }
        } else {
            rank = knn;  // padding never ranks
        }
        if (rank < knn) {
            if (key_out) {
                key_out[row * knn + rank] = v;
                if (aux_out) aux_out[row * knn + rank] = aux ? aux[src] : 0.0f;
            } else {
                const float key = float_from_order_bits((uint32_t)(v >> 32));
                idx_out[row * knn + rank] = (unsigned long long)(uint32_t)v;
                d0_out[row * knn + rank] = kind == KIND_ANI ? 1.0f - key : key;
                if (d1_out) d1_out[row * knn + rank] = aux ? aux[src] : 0.0f;
            }
        }
    }
}

void launch_knn_finalize(const unsigned long long *keys, const float *aux, uint32_t n_parts, uint64_t part_stride,
                         uint64_t row0, uint64_t n_rows, uint32_t knn, int kind, unsigned long long *idx_out,
                         float *d0_out, float *d1_out, unsigned long long *key_out, float *aux_out, cudaStream_t st) {
    if (n_rows == 0) return;
    const int warps = 8;
    knn_finalize_kernel<<<(unsigned)((n_rows + warps - 1) / warps), warps * 32, 0, st>>>(
        keys, aux, n_parts, part_stride, row0, n_rows, knn, kind, idx_out, d0_out, d1_out, key_out, aux_out);
}

__global__ void fill_u64_kernel(unsigned long long *p, uint64_t n, unsigned long long v) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = v;
}
void launch_fill_u64(unsigned long long *p, uint64_t n, unsigned long long v, cudaStream_t st) {
    if (n) fill_u64_kernel<<<(unsigned)std::min<uint64_t>(4096, (n + 255) / 256), 256, 0, st>>>(p, n, v);
}

// Mismatching bins of one (sample of A, sample of B) pair at k-mer index k, computed by a whole warp straight
// from the tiled layout in global memory (lanes stride over the 32-bin halves).  Every lane gets the total.
__device__ __forceinline__ uint32_t warp_pair_mismatches(const DevDb &A, const DevDb &B, uint32_t k, uint32_t ba,
                                                         uint32_t sa, uint32_t bb, uint32_t sb, int lane) {
    uint32_t mism = 0;
    for (uint32_t wc = lane; wc < 2 * A.W; wc += 32) {
        const uint32_t *pa = A.tiled + slab_index(A, k, wc >> 1, ba) * STAGE_U32 + (wc & 1) * HALF_U32;
        const uint32_t *pb = B.tiled + slab_index(B, k, wc >> 1, bb) * STAGE_U32 + (wc & 1) * HALF_U32;
        uint32_t d = 0;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const uint4 a = reinterpret_cast<const uint4 *>(pa + q * QUAD_U32)[sa];
            const uint4 b = reinterpret_cast<const uint4 *>(pb + q * QUAD_U32)[sb];
            d |= (a.x ^ b.x) | (a.y ^ b.y) | (a.z ^ b.z) | (a.w ^ b.w);
        }
        {
            const uint2 a = reinterpret_cast<const uint2 *>(pa + TAIL_U32)[sa];
            const uint2 b = reinterpret_cast<const uint2 *>(pb + TAIL_U32)[sb];
            d |= (a.x ^ b.x) | (a.y ^ b.y);
        }
        mism += __popc(d);
    }
    for (int o = 16; o > 0; o >>= 1) mism += __shfl_xor_sync(0xffffffffu, mism, o);
    return mism;
}

// ------------------------------------------------------------------------------------------------
// Candidate-list kNN (the distance half of self_dists_knn_precluster, reference src/distances/mod.rs:399-553):
// one CTA per row walks the row's candidate list (CSR, .skd indices, the row itself is skipped), one pair per
// warp at a time, and keeps the knn best through a threshold + sort-in-rounds buffer in shared memory.
// Rows with fewer than knn neighbours are padded with (row, 1.0); retain == 1 (Singleton) turns a row without
// any neighbour into (row, 0.0), (row, 1.0)...  Single-k modes only, like the reference.
__global__ void __launch_bounds__(128) knn_candidates_kernel(const __grid_constant__ DevDb A,
                                                             const __grid_constant__ EpiParams E,
                                                             const uint64_t *__restrict__ cand_off,
                                                             const uint64_t *__restrict__ cand_idx, uint32_t knn,
                                                             uint32_t cap, int retain,
                                                             unsigned long long *__restrict__ idx_out,
                                                             float *__restrict__ d0_out) {
    extern __shared__ __align__(16) unsigned long long cbuf[];  // [cap]
    __shared__ uint32_t cnt;
    __shared__ unsigned long long thr;
    const uint64_t row = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (uint32_t e = tid; e < cap; e += 128) cbuf[e] = ~0ull;
    if (tid == 0) {
        cnt = 0;
        thr = ~0ull;
    }
    __syncthreads();
    const uint64_t lo = cand_off[row], hi = cand_off[row + 1];
    const uint32_t ba = (uint32_t)(row / TS), sa = (uint32_t)(row % TS);
    const bool comp = E.has_c != 0;
    const double c1 = comp ? E.a_comp[row] : 0.0;
    const uint32_t round = cap - knn;  // new entries a round may add on top of the knn kept
    for (uint64_t base = lo; base < hi; base += round) {
        const uint64_t end = hi < base + round ? hi : base + round;
        const unsigned long long th = thr;
        for (uint64_t c = base + warp; c < end; c += 4) {
            const uint64_t j = cand_idx[c];
            if (j == row || j >= A.n) continue;
            const uint32_t mism = warp_pair_mismatches(A, A, (uint32_t)E.k0, ba, sa, (uint32_t)(j / TS), (uint32_t)(j % TS), lane);
            if (lane == 0) {
                const uint32_t count = E.nbins - mism;
                float key;
                if (!comp) {
                    key = E.key_tab[count];
                } else {
                    const double jac = corrected_jaccard(count, c1, E.b_comp[j], E);
                    key = E.kind == KIND_ANI ? (float)__dsub_rn(1.0, ani_pois_dev(jac, E.k_single)) : (float)__dsub_rn(1.0, jac);
                }
                const unsigned long long v = ((unsigned long long)float_order_bits(key) << 32) | (uint32_t)j;
                if (v < th) cbuf[atomicAdd(&cnt, 1u)] = v;
            }
        }
        __syncthreads();
        if (warp == 0) {
            warp_bitonic_sort(cbuf, nullptr, (int)cap, lane);
            for (uint32_t e = knn + lane; e < cap; e += 32) cbuf[e] = ~0ull;
            __syncwarp();
            if (lane == 0) {
                cnt = cnt < knn ? cnt : knn;
                thr = cnt == knn ? cbuf[knn - 1] : ~0ull;
            }
        }
        __syncthreads();
    }
    const uint32_t have = cnt;
    for (uint32_t t = tid; t < knn; t += 128) {
        unsigned long long oi = row;
        float od = 1.0f;  // padding: (row, 1.0), mod.rs:537-544
        if (t < have) {
            const unsigned long long v = cbuf[t];
            const float key = float_from_order_bits((uint32_t)(v >> 32));
            oi = (uint32_t)v;
            od = E.kind == KIND_ANI ? 1.0f - key : key;
        } else if (have == 0 && retain == 1 && t == 0) {
            od = 0.0f;  // Singleton, mod.rs:491-499
        }
        idx_out[row * knn + t] = oi;
        d0_out[row * knn + t] = od;
    }
}

cudaError_t launch_knn_candidates(const DevDb &A, const EpiParams &E, const uint64_t *cand_off, const uint64_t *cand_idx,
                                  uint64_t n_rows, uint32_t knn, uint32_t cap, int retain, unsigned long long *idx_out,
                                  float *d0_out, cudaStream_t st) {
    if (n_rows == 0) return cudaSuccess;
    const size_t smem = (size_t)cap * 8;
    cudaError_t err = cudaFuncSetAttribute(knn_candidates_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    knn_candidates_kernel<<<(unsigned)n_rows, 128, smem, st>>>(A, E, cand_off, cand_idx, knn, cap, retain, idx_out, d0_out);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Explicit pair list: one warp per pair.  counts_out (nullable): [pair][K]; dist_out (nullable): 1 or 2 floats.
__global__ void pair_kernel(const __grid_constant__ DevDb A, const __grid_constant__ DevDb B,
                            const __grid_constant__ EpiParams E, const uint64_t *__restrict__ pairs, uint64_t n_pairs,
                            uint32_t *__restrict__ counts_out, float *__restrict__ dist_out) {
    const uint64_t p = (uint64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x & 31;
    if (p >= n_pairs) return;
    const uint64_t gi = pairs[2 * p], gj = pairs[2 * p + 1];
    const uint32_t ba = (uint32_t)(gi / TS), sa = (uint32_t)(gi % TS), bb = (uint32_t)(gj / TS), sb = (uint32_t)(gj % TS);
    const bool all_k = counts_out != nullptr || E.kind == KIND_CORE_ACC;
    const int k_lo = all_k ? 0 : E.k0, k_hi = all_k ? (int)A.K : E.k0 + 1;
    double ysum = 0.0, xysum = 0.0, ysq = 0.0;
    int n = 0;
    bool broke = false;
    const bool comp = E.has_c != 0;
    double c1 = 0.0, c2 = 0.0;
    if (comp) {
        c1 = E.a_comp[gi];
        c2 = E.b_comp[gj];
    }
    float d0 = 0.0f, d1 = 0.0f;
    for (int k = k_lo; k < k_hi; ++k) {
        const uint32_t mism = warp_pair_mismatches(A, B, (uint32_t)k, ba, sa, bb, sb, lane);
        const uint32_t c = E.nbins - mism;
        if (counts_out && lane == 0) counts_out[p * A.K + k] = c;
        if (!dist_out) continue;
        if (E.kind == KIND_CORE_ACC) {
            if (!broke) {
                double y = comp ? log(corrected_jaccard(c, c1, c2, E)) : E.lny[c];
                if (y < E.tolerance) {
                    broke = true;
                } else {
                    ysum = __dadd_rn(ysum, y);
                    xysum = __dadd_rn(xysum, __dmul_rn(E.kf[k], y));
                    ysq = __dadd_rn(ysq, __dmul_rn(y, y));
                    ++n;
                }
            }
        } else if (k == E.k0) {
            if (!comp) {
                d0 = E.val_tab[c];
            } else {
                double j = corrected_jaccard(c, c1, c2, E);
                d0 = E.kind == KIND_ANI ? (float)ani_pois_dev(j, E.k_single) : (float)__dsub_rn(1.0, j);
            }
        }
    }
    if (dist_out && lane == 0) {
        if (E.kind == KIND_CORE_ACC) {
            regress(ysum, xysum, ysq, n, E, d0, d1);
            dist_out[2 * p] = d0;
            dist_out[2 * p + 1] = d1;
        } else {
            dist_out[p] = d0;
        }
    }
}

void launch_pairs(const DevDb &A, const DevDb &B, const EpiParams &E, const uint64_t *pairs, uint64_t n_pairs,
                  uint32_t *counts_out, float *dist_out, cudaStream_t st) {
    if (n_pairs == 0) return;
    const int warps = 8;
    pair_kernel<<<(unsigned)((n_pairs + warps - 1) / warps), warps * 32, 0, st>>>(A, B, E, pairs, n_pairs, counts_out,
                                                                                   dist_out);
}

// With lazy module loading (the CUDA 12 default) the first launch of a kernel loads it, and that load can wait for
// the device to go idle.  A peer barrier (peer.cu) spins on the device until the other ranks arrive, so a rank that
// shares the GPU with a spinning peer must never load code at that moment: every kernel is loaded when a context is
// created.
cudaError_t preload_kernels() {
    cudaFuncAttributes a;
    cudaError_t e = cudaSuccess;
#define SKD_PRELOAD(...)                                                              \
    do {                                                                              \
        if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, (const void *)(__VA_ARGS__)); \
    } while (0)
    SKD_PRELOAD(repack_kernel);
    SKD_PRELOAD(tile_kernel<uint16_t, 8>);
    SKD_PRELOAD(tile_kernel<uint16_t, 16>);
    SKD_PRELOAD(tile_kernel<uint16_t, 32>);
    SKD_PRELOAD(tile_kernel<uint16_t, TS>);
    SKD_PRELOAD(tile_kernel<uint32_t, 8>);
    SKD_PRELOAD(tile_kernel<uint32_t, 16>);
    SKD_PRELOAD(tile_kernel<uint32_t, 32>);
    SKD_PRELOAD(tile_kernel<uint32_t, TS>);
    SKD_PRELOAD((dense_kernel<uint16_t, false>));
    SKD_PRELOAD((dense_kernel<uint16_t, true>));
    SKD_PRELOAD((dense_kernel<uint32_t, false>));
    SKD_PRELOAD((dense_kernel<uint32_t, true>));
    SKD_PRELOAD(defer_kernel);
    SKD_PRELOAD(untile_kernel);
    SKD_PRELOAD((stream_kernel<1, false>));
    SKD_PRELOAD((stream_kernel<2, false>));
    SKD_PRELOAD((stream_kernel<4, false>));
    SKD_PRELOAD((stream_kernel<8, false>));
    SKD_PRELOAD((stream_kernel<16, true>));
    SKD_PRELOAD(knn_merge_kernel);
    SKD_PRELOAD(select_few_kernel<true>);
    SKD_PRELOAD(select_few_kernel<false>);
    SKD_PRELOAD(select_sym_kernel<true, true>);
    SKD_PRELOAD(select_sym_kernel<true, false>);
    SKD_PRELOAD(select_sym_kernel<false, true>);
    SKD_PRELOAD(select_sym_kernel<false, false>);
    SKD_PRELOAD(knn_finalize_kernel);
    SKD_PRELOAD(fill_u64_kernel);
    SKD_PRELOAD(knn_candidates_kernel);
    SKD_PRELOAD(pair_kernel);
#undef SKD_PRELOAD
    return e;
}

}  // namespace skd
