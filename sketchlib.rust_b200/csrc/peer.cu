// Peer groups (include/sketchdist.h, skd_comm_*): several GPUs of one box working on one distance problem.
//
// Every rank owns a WINDOW of device memory (cudaMalloc, so that it can be exported with CUDA IPC) that its peers map
// -- cudaIpcOpenMemHandle across processes, plain peer access between threads of one process -- and read with
// ordinary loads over NVLink.  The exchange steps of the path are this library's own kernels working on peer
// addresses, not library collectives:
//   peer_barrier_kernel    flag barrier: st.release.sys of an epoch into every peer's window, ld.acquire.sys spin on
//                          the own flags (bounded by a timeout, so a missing rank is an error, not a hang)
//   peer_gather_db_kernel  distributed loader: every rank re-tiles its own sample slice, then pulls the slabs of the
//                          other slices from the peers' windows (LDG.128 over NVLink) into its full database
//   peer_merge_kernel      sparse kNN: the owner of a row pulls that row's partial lists from all peers into shared
//                          memory (one coalesced round trip) and merges them by rank-by-binary-search
//   peer_copy_kernel       gather of the final sparse rows on the root rank
// Window layout: [flags: 4096 B][data: window_bytes], every collective call lays its buffers out in the data area and
// starts with a barrier, so a call never overwrites memory a peer may still be reading from the previous call.
#include <cuda_runtime.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "internal.h"

using namespace skd;
using namespace skdi;

namespace {
constexpr size_t kFlagBytes = 4096;
constexpr int kMaxWorld = SKD_COMM_MAX_WORLD;

struct HandleBlob {  // what skd_comm_export writes: SKD_COMM_HANDLE_BYTES bytes
    cudaIpcMemHandle_t ipc;   // 64 B
    uint64_t ptr;             // the window's address in the exporting process
    uint64_t window_bytes;
    int32_t device;
    int32_t pid;
    int32_t rank, world;
    unsigned char pad[SKD_COMM_HANDLE_BYTES - 64 - 8 - 8 - 16];
};
static_assert(sizeof(HandleBlob) == SKD_COMM_HANDLE_BYTES, "handle blob size");

struct PeerBases {
    unsigned char *base[kMaxWorld];
};

size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }
}  // namespace

struct skd_comm {
    skd_ctx *ctx = nullptr;
    int rank = 0, world = 1;
    size_t window_bytes = 0;
    unsigned char *window = nullptr;  // flags + data
    PeerBases peers{};                // peers.base[r]: rank r's window as mapped into this process
    bool opened[kMaxWorld] = {};      // mapped with cudaIpcOpenMemHandle (to be closed)
    bool connected = false;
    uint32_t epoch = 0;
    uint32_t *err_host = nullptr, *err_dev = nullptr;  // mapped pinned word: result of the last barrier wait
    double timeout_s = 60.0;
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};  // op begin / local kernels done / op end
    float last_ms = 0.f, last_xchg_ms = 0.f;
    bool adaptive_split = true;  // SKD_ADAPTIVE_SPLIT=0: always the analytic equal-pair triangle split
    uint64_t last_r0 = 0, last_r1 = 0;  // triangle rows this rank evaluated in the last symmetric self-kNN call
    unsigned char *data(int r) const { return peers.base[r] + kFlagBytes; }
};

namespace {

// ---- kernels over peer memory --------------------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint64_t global_timer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Flag barrier.  One thread per peer: (signal) publish `epoch` in the peer's flag slot of this rank, (wait) spin on the
// peer's slot in the own window for at most spin_ns, then report in *status whether every peer has arrived.  The host
// re-launches the wait until it succeeds (enqueue_barrier): the device never spins for long, so another host thread that
// shares this GPU and sits in a device-synchronising driver call (a first kernel load, a memory-pool growth) always
// gets through -- a pure device spin dead-locks against it.  Everything this rank's stream did before the barrier is
// visible to a peer that has passed it.
__global__ void peer_barrier_kernel(PeerBases P, int rank, int world, uint32_t epoch, int do_signal, uint64_t spin_ns,
                                    uint32_t *status) {
    const int p = threadIdx.x;
    bool ok = true;
    if (p < world) {
        if (do_signal) {
            __threadfence_system();
            st_release_sys(reinterpret_cast<uint32_t *>(P.base[p]) + rank, epoch);
        }
        const uint32_t *mine = reinterpret_cast<const uint32_t *>(P.base[rank]) + p;
        const uint64_t t0 = global_timer_ns();
        while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
            if (global_timer_ns() - t0 > spin_ns) {
                ok = false;
                break;
            }
            __nanosleep(100);
        }
        __threadfence_system();
    }
    const unsigned all = __ballot_sync(0xffffffffu, ok);
    if (p == 0) {
        uint32_t missing = 0;  // lowest rank that has not arrived, + 1
        if (all != 0xffffffffu) missing = (uint32_t)__ffs(~all);
        *reinterpret_cast<volatile uint32_t *>(status) = missing;
    }
}

// Distributed loader, second half.  Rank p's window holds its slice re-tiled as [K*W rows][nb_p blocks] slabs
// (contiguous); the full database has nb_total blocks per row.  blockIdx.y = source rank; 16-byte units.
struct GatherDbParams {
    const uint4 *src[kMaxWorld];
    uint32_t nb[kMaxWorld];  // blocks of the slice
    uint32_t b0[kMaxWorld];  // first block of the slice in the full database
    uint32_t rows;           // K * W
    uint32_t nb_total;
};
__global__ void __launch_bounds__(256) peer_gather_db_kernel(const __grid_constant__ GatherDbParams G, uint4 *__restrict__ dst) {
    constexpr uint32_t U = STAGE_BYTES / 16;  // 16-byte units per slab
    const int p = blockIdx.z;                 // source rank
    const uint32_t per_row = G.nb[p] * U;
    const uint4 *src = G.src[p];
    for (uint32_t row = blockIdx.y; row < G.rows; row += gridDim.y) {
        const uint4 *s = src + (uint64_t)row * per_row;
        uint4 *d = dst + ((uint64_t)row * G.nb_total + G.b0[p]) * U;
        for (uint32_t col = blockIdx.x * blockDim.x + threadIdx.x; col < per_row; col += gridDim.x * blockDim.x) d[col] = s[col];
    }
}

struct CopyTable {
    const uint4 *src[3 * kMaxWorld];
    uint4 *dst[3 * kMaxWorld];
    uint64_t n16[3 * kMaxWorld];
};
__global__ void __launch_bounds__(256) peer_copy_kernel(const __grid_constant__ CopyTable T) {
    const int s = blockIdx.y;
    const uint4 *src = T.src[s];
    uint4 *dst = T.dst[s];
    const uint64_t n = T.n16[s];
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (uint64_t)gridDim.x * blockDim.x) dst[e] = src[e];
}

// Sparse kNN merge over peer memory.  part p of row r lives at key[p] + r * knn (ascending composites, ~0 padded;
// aux likewise); parts with first_row[p] > r hold only padding for that row and are not read.  One warp per row:
// the lists are first pulled into shared memory (coalesced loads, one NVLink round trip per part), then every entry
// finds its rank by binary search in each list (composites are unique within a row).  Output rows [row0, row0 + n_rows)
// go to idx/d0/d1 at (row - row0) * knn; ANI keys are un-transformed as 1.0f - key (src/distances/mod.rs:183-189).
struct MergeParts {
    const unsigned long long *key[kMaxWorld];
    const float *aux[kMaxWorld];
    uint64_t first_row[kMaxWorld];
    uint32_t n_parts;
    uint32_t knn;
};
template <bool STAGED>
__global__ void __launch_bounds__(256) peer_merge_kernel(const __grid_constant__ MergeParts M, uint64_t row0, uint64_t n_rows,
                                                         int kind, bool with_aux, unsigned long long *__restrict__ idx_out,
                                                         float *__restrict__ d0_out, float *__restrict__ d1_out) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t knn = M.knn, total = M.n_parts * knn;
    unsigned long long *skey = reinterpret_cast<unsigned long long *>(sm_raw) + (size_t)warp * total;
    float *saux = reinterpret_cast<float *>(sm_raw + (size_t)(blockDim.x >> 5) * total * 8) + (size_t)warp * total;
    const uint64_t lrow = (uint64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (lrow >= n_rows) return;
    const uint64_t row = row0 + lrow;
    if (STAGED) {
        for (uint32_t e = lane; e < total; e += 32) {
            const uint32_t p = e / knn, t = e - p * knn;
            unsigned long long v = ~0ull;
            float a = 0.f;
            if (M.first_row[p] <= row) {
                v = M.key[p][row * knn + t];
                if (with_aux && v != ~0ull) a = M.aux[p][row * knn + t];
            }
            skey[e] = v;
            if (with_aux) saux[e] = a;
        }
        __syncwarp();
    }
    // STAGED = false (lists too long for shared memory: very large knn): the same search straight in peer memory
    for (uint32_t e = lane; e < total; e += 32) {
        const uint32_t pe = e / knn, te = e - pe * knn;
        if (!STAGED && M.first_row[pe] > row) continue;
        const unsigned long long v = STAGED ? skey[e] : M.key[pe][row * knn + te];
        if (v == ~0ull) continue;
        uint32_t rank = 0;
        for (uint32_t s = 0; s < M.n_parts; ++s) {
            if (!STAGED && M.first_row[s] > row) continue;
            const unsigned long long *l = STAGED ? skey + s * knn : M.key[s] + row * knn;
            uint32_t lo = 0, hi = knn;
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (l[mid] < v) lo = mid + 1; else hi = mid;
            }
            rank += lo;
        }
        if (rank < knn) {
            const float key = float_from_order_bits((uint32_t)(v >> 32));
            idx_out[lrow * knn + rank] = (unsigned long long)(uint32_t)v;
            d0_out[lrow * knn + rank] = kind == KIND_ANI ? 1.0f - key : key;
            if (with_aux) d1_out[lrow * knn + rank] = STAGED ? saux[e] : M.aux[pe][row * knn + te];
        }
    }
}

// ---- host side ------------------------------------------------------------------------------------------------
double now_s() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + ts.tv_nsec * 1e-9;
}
bool g_debug = getenv("SKD_PEER_DEBUG") != nullptr;
#define PEER_LOG(c, ...)                                                     \
    do {                                                                     \
        if (g_debug) {                                                       \
            fprintf(stderr, "[peer %d/%d %.6f] ", (c)->rank, (c)->world, now_s()); \
            fprintf(stderr, __VA_ARGS__);                                    \
            fputc('\n', stderr);                                             \
        }                                                                    \
    } while (0)

skd_status comm_fail(skd_comm *c, skd_status st, const char *msg) { return fail(c ? c->ctx : nullptr, st, "%s", msg); }

// What a rank leaves in the flag area of its window (offset kSplitOff) at the end of a symmetric self-kNN call: how long
// its local part took and how many pairs that was.  The next call with the same tag (same problem shape) re-splits the
// triangle so that every rank's predicted time is the same: equal pair counts are not equal times (the ranks' key
// matrices differ in shape -- few long rows at the top of the triangle, many short ones at the bottom -- and so do
// their select kernels and diagonal waste).  Every rank reads the same N records after the entry barrier, so all
// compute the same boundaries; results do not depend on the split.
constexpr size_t kSplitOff = 2048;
struct SplitStat {
    float local_ms;
    uint32_t tag;
    uint64_t pairs;
};
uint64_t triangle_bound(uint64_t n, double f) {  // row after a fraction f of the pairs, the rule of skd_triangle_rows
    if (f <= 0.0) return 0;
    if (f >= 1.0) return n;
    const double x = (double)n * (1.0 - std::sqrt(1.0 - f));
    return std::min<uint64_t>(n, (uint64_t)std::llround(x / TS) * TS);
}
skd_status split_triangle(skd_comm *c, uint64_t n, uint32_t tag, uint64_t *bounds /* world + 1 */) {
    const int W = c->world;
    double share[kMaxWorld];
    for (int p = 0; p < W; ++p) share[p] = 1.0 / W;
    if (c->adaptive_split && W > 1) {
        SplitStat st[kMaxWorld];
        for (int p = 0; p < W; ++p)
            CU(c->ctx, cudaMemcpyAsync(&st[p], c->peers.base[p] + kSplitOff, sizeof(SplitStat), cudaMemcpyDeviceToHost,
                                       c->ctx->compute));
        CU(c->ctx, cudaStreamSynchronize(c->ctx->compute));
        bool ok = true;
        double pairs_all = 0.0, speed_all = 0.0;
        for (int p = 0; p < W; ++p) {
            ok = ok && st[p].tag == tag && st[p].pairs > 0 && st[p].local_ms > 0.f && std::isfinite(st[p].local_ms);
            if (ok) {
                pairs_all += (double)st[p].pairs;
                speed_all += (double)st[p].pairs / st[p].local_ms;
            }
        }
        if (ok) {
            double sum = 0.0;
            for (int p = 0; p < W; ++p) {
                const double prev = (double)st[p].pairs / pairs_all;                    // what the rank had
                const double want = ((double)st[p].pairs / st[p].local_ms) / speed_all;  // what its speed deserves
                share[p] = std::min(2.0 / W, std::max(0.5 / W, 0.25 * prev + 0.75 * want));  // damped, bounded
                sum += share[p];
            }
            for (int p = 0; p < W; ++p) share[p] /= sum;
        }
    }
    skd_triangle_rows_weighted(n, W, share, bounds);
    return SKD_OK;
}
skd_status leave_split_stat(skd_comm *c, uint32_t tag, uint64_t pairs) {
    if (c->world == 1) return SKD_OK;
    // the device time of this rank's own kernels (distance tiles, deferred pairs, list folds): CUDA-event timers around
    // the launches, so neither the waits at the barriers nor host-side skew between the ranks enter the split
    const SplitStat st{c->ctx->last_ms, tag, pairs};
    CU(c->ctx, cudaMemcpyAsync(c->window + kSplitOff, &st, sizeof st, cudaMemcpyHostToDevice, c->ctx->compute));
    CU(c->ctx, cudaStreamSynchronize(c->ctx->compute));  // st is a local; the next call's entry barrier publishes it
    return SKD_OK;
}

// One flag barrier on the compute stream; returns when every rank has reached it (the stream is idle then).
skd_status enqueue_barrier(skd_comm *c) {
    if (c->world == 1) return SKD_OK;
    if (!c->connected) return comm_fail(c, SKD_ERR_INVALID, "peer group is not connected (skd_comm_connect)");
    skd_ctx *ctx = c->ctx;
    c->epoch++;
    PEER_LOG(c, "barrier %u", c->epoch);
    const double t0 = now_s();
    for (int attempt = 0;; ++attempt) {
        // first wait: 2 ms on the device (ranks of a balanced job arrive within that); later ones short, with the host
        // sleeping in between so that the GPU is free for whoever is blocked on it
        const uint64_t spin_ns = attempt == 0 ? 2000000ull : 200000ull;
        peer_barrier_kernel<<<1, 32, 0, ctx->compute>>>(c->peers, c->rank, c->world, c->epoch, attempt == 0, spin_ns, c->err_dev);
        CU(ctx, cudaGetLastError());
        ctx->launches++;
        CU(ctx, cudaStreamSynchronize(ctx->compute));
        const uint32_t missing = *reinterpret_cast<volatile uint32_t *>(c->err_host);
        if (missing == 0) return SKD_OK;
        if (now_s() - t0 > c->timeout_s)
            return fail(ctx, SKD_ERR_CUDA, "peer barrier timed out after %.0f s: rank %u of %d never arrived", c->timeout_s,
                        missing - 1, c->world);
        usleep(attempt < 50 ? 50 : 500);
    }
}

// Ends a collective call: synchronises the compute stream and collects the timings.
skd_status finish_op(skd_comm *c) {
    CU(c->ctx, cudaStreamSynchronize(c->ctx->compute));
    PEER_LOG(c, "op finished (epoch %u)", c->epoch);
    float a = 0.f, b = 0.f;
    CU(c->ctx, cudaEventElapsedTime(&a, c->ev[0], c->ev[2]));
    CU(c->ctx, cudaEventElapsedTime(&b, c->ev[1], c->ev[2]));
    c->last_ms = a;
    c->last_xchg_ms = b;
    return timers_collect(c->ctx);
}

struct KnnLayout {  // offsets into the window's data area
    size_t keys, aux, fin_idx, fin_d0, fin_d1, end;
    uint64_t rows_per;
};
KnnLayout knn_layout(uint64_t n_rows, uint64_t knn, int world) {
    KnnLayout L{};
    L.rows_per = (n_rows + world - 1) / world;
    size_t off = 0;
    L.keys = off;
    off = align_up(off + n_rows * knn * 8);
    L.aux = off;
    off = align_up(off + n_rows * knn * 4);
    L.fin_idx = off;
    off = align_up(off + L.rows_per * knn * 8);
    L.fin_d0 = off;
    off = align_up(off + L.rows_per * knn * 4);
    L.fin_d1 = off;
    off = align_up(off + L.rows_per * knn * 4);
    L.end = off;
    return L;
}

// After the partial lists (or the owned final rows, when `merge` is false) are in the windows: barrier, owner merge,
// barrier, gather on the root.  parts_first_row[p]: first row rank p's lists can hold real entries for.
skd_status exchange_rows(skd_comm *c, const KnnLayout &L, uint64_t n_rows, uint64_t knn, int kind, bool merge,
                         const uint64_t *parts_first_row, int root, uint64_t *idx_out, float *d0_out, float *d1_out,
                         skd_loc out_loc) {
    skd_ctx *ctx = c->ctx;
    const bool with_aux = kind == KIND_CORE_ACC;
    uint64_t o0, o1;
    skd_slice_rows(n_rows, c->rank, c->world, 1, &o0, &o1);
    CU(ctx, cudaEventRecord(c->ev[1], ctx->compute));
    skd_status st;
    if (merge) {
        st = enqueue_barrier(c);  // every rank's partial lists are complete
        if (st) return st;
        if (o1 > o0) {
            MergeParts M{};
            M.n_parts = (uint32_t)c->world;
            M.knn = (uint32_t)knn;
            for (int p = 0; p < c->world; ++p) {
                M.key[p] = reinterpret_cast<const unsigned long long *>(c->data(p) + L.keys);
                M.aux[p] = reinterpret_cast<const float *>(c->data(p) + L.aux);
                M.first_row[p] = parts_first_row ? parts_first_row[p] : 0;
            }
            const size_t per_warp = (size_t)c->world * knn * (with_aux ? 12 : 8);
            const bool staged = per_warp <= 200 * 1024;
            const int warps = staged ? (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / per_warp)) : 8;
            const size_t smem = staged ? (size_t)warps * per_warp : 0;
            unsigned char *mine = c->data(c->rank);
            const unsigned grid = (unsigned)((o1 - o0 + warps - 1) / warps);
            auto *oi = reinterpret_cast<unsigned long long *>(mine + L.fin_idx);
            float *o0p = reinterpret_cast<float *>(mine + L.fin_d0), *o1p = reinterpret_cast<float *>(mine + L.fin_d1);
            if (staged) {
                CU(ctx, cudaFuncSetAttribute(peer_merge_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                peer_merge_kernel<true><<<grid, warps * 32, smem, ctx->compute>>>(M, o0, o1 - o0, kind, with_aux, oi, o0p, o1p);
            } else {
                peer_merge_kernel<false><<<grid, warps * 32, 0, ctx->compute>>>(M, o0, o1 - o0, kind, with_aux, oi, o0p, o1p);
            }
            CU(ctx, cudaGetLastError());
            ctx->launches++;
        }
    }
    st = enqueue_barrier(c);  // every rank's final rows are complete
    if (st) return st;
    if (root < 0 || root == c->rank) {
        if (!idx_out || !d0_out || (with_aux && !d1_out)) return comm_fail(c, SKD_ERR_INVALID, "null output on the root rank");
        unsigned long long *didx = (unsigned long long *)idx_out;
        float *dd0 = d0_out, *dd1 = d1_out;
        if (out_loc == SKD_HOST) {
            void *p;
            st = scratch_get(ctx, "knn_idx", n_rows * knn * 8, &p);
            if (st) return st;
            didx = (unsigned long long *)p;
            st = scratch_get(ctx, "knn_d0", n_rows * knn * 4, &p);
            if (st) return st;
            dd0 = (float *)p;
            if (with_aux) {
                st = scratch_get(ctx, "knn_d1", n_rows * knn * 4, &p);
                if (st) return st;
                dd1 = (float *)p;
            }
        }
        CopyTable T{};
        int ns = 0;
        uint64_t most = 0;
        for (int p = 0; p < c->world; ++p) {
            uint64_t a, b;
            skd_slice_rows(n_rows, p, c->world, 1, &a, &b);
            if (b <= a) continue;
            const uint64_t ents = (b - a) * knn;
            auto add = [&](size_t src_off, void *dst, uint64_t bytes) {
                // rows_per * knn * {8, 4} bytes per rank: starts are 16-byte aligned when rows_per * knn % 4 == 0;
                // otherwise fall back to a plain copy below
                T.src[ns] = reinterpret_cast<const uint4 *>(c->data(p) + src_off);
                T.dst[ns] = reinterpret_cast<uint4 *>(dst);
                T.n16[ns] = bytes / 16;
                most = std::max(most, T.n16[ns]);
                ++ns;
            };
            const bool aligned = (a * knn) % 4 == 0 && ents % 4 == 0 && (uintptr_t)didx % 16 == 0 &&
                                 (uintptr_t)dd0 % 16 == 0 && (!with_aux || (uintptr_t)dd1 % 16 == 0);
            if (aligned) {
                add(L.fin_idx, didx + a * knn, ents * 8);
                add(L.fin_d0, dd0 + a * knn, ents * 4);
                if (with_aux) add(L.fin_d1, dd1 + a * knn, ents * 4);
            } else {
                CU(ctx, cudaMemcpyAsync(didx + a * knn, c->data(p) + L.fin_idx, ents * 8, cudaMemcpyDefault, ctx->compute));
                CU(ctx, cudaMemcpyAsync(dd0 + a * knn, c->data(p) + L.fin_d0, ents * 4, cudaMemcpyDefault, ctx->compute));
                if (with_aux)
                    CU(ctx, cudaMemcpyAsync(dd1 + a * knn, c->data(p) + L.fin_d1, ents * 4, cudaMemcpyDefault, ctx->compute));
            }
        }
        if (ns) {
            const unsigned gx = (unsigned)std::min<uint64_t>(std::max<uint64_t>(1, (most + 255) / 256), 4u * ctx->sm_count / std::max(1, ns) + 1);
            peer_copy_kernel<<<dim3(gx, ns), 256, 0, ctx->compute>>>(T);
            CU(ctx, cudaGetLastError());
            ctx->launches++;
        }
        CU(ctx, cudaEventRecord(c->ev[2], ctx->compute));
        if (out_loc == SKD_HOST) {
            CU(ctx, cudaMemcpyAsync(idx_out, didx, n_rows * knn * 8, cudaMemcpyDeviceToHost, ctx->compute));
            CU(ctx, cudaMemcpyAsync(d0_out, dd0, n_rows * knn * 4, cudaMemcpyDeviceToHost, ctx->compute));
            if (with_aux) CU(ctx, cudaMemcpyAsync(d1_out, dd1, n_rows * knn * 4, cudaMemcpyDeviceToHost, ctx->compute));
        }
    } else {
        CU(ctx, cudaEventRecord(c->ev[2], ctx->compute));
    }
    return finish_op(c);
}

skd_status begin_op(skd_comm *c) {
    if (!c || !c->ctx) return SKD_ERR_INVALID;
    CU(c->ctx, cudaSetDevice(c->ctx->device));
    if (c->world > 1 && !c->connected) return comm_fail(c, SKD_ERR_INVALID, "peer group is not connected (skd_comm_connect)");
    timers_reset(c->ctx);
    CU(c->ctx, cudaEventRecord(c->ev[0], c->ctx->compute));
    return enqueue_barrier(c);  // no peer is still reading this window on behalf of the previous call
}

// Tail of both distributed loaders: `slice` is this rank's part of the database, complete in HBM.
skd_status gather_db(skd_comm *c, skd_db *slice, uint64_t n_total, uint64_t sample_stride, uint64_t kmer_stride, uint64_t W,
                     uint64_t sketch_size, const uint64_t *kmers, uint32_t K, skd_db **out) {
    skd_ctx *ctx = c->ctx;
    auto drop = [&](skd_status st) {
        if (slice) skd_db_free(slice);
        return st;
    };
    const size_t slice_bytes = slice ? slice->tiled_bytes : 0;
    if (slice_bytes > c->window_bytes)
        return drop(fail(ctx, SKD_ERR_INVALID, "peer window of %zu bytes is too small for a database slice of %zu bytes",
                         c->window_bytes, slice_bytes));
    if (slice_bytes)
        CU(ctx, cudaMemcpyAsync(c->data(c->rank), slice->tiled, slice_bytes, cudaMemcpyDeviceToDevice, ctx->compute));
    CU(ctx, cudaEventRecord(c->ev[1], ctx->compute));
    skd_status st = enqueue_barrier(c);  // every rank's slice is in its window
    if (st) return drop(st);
    ChunkSource none;
    none.alloc_only = true;
    skd_db *db = nullptr;
    st = db_build(ctx, none, n_total, sample_stride, kmer_stride, W, sketch_size, kmers, K, &db);
    if (st) return drop(st);
    GatherDbParams G{};
    G.rows = (uint32_t)(K * W);
    G.nb_total = db->dev.nb;
    uint32_t most = 0;
    for (int p = 0; p < c->world; ++p) {
        uint64_t a, b;
        skd_slice_rows(n_total, p, c->world, TS, &a, &b);
        G.src[p] = reinterpret_cast<const uint4 *>(c->data(p));
        G.b0[p] = (uint32_t)(a / TS);
        G.nb[p] = (uint32_t)((b - a + TS - 1) / TS);
        most = std::max(most, G.nb[p]);
    }
    if (most) {
        const uint32_t per_row = most * (STAGE_BYTES / 16);
        const unsigned gx = std::max(1u, std::min(16u, (per_row + 255) / 256));
        const unsigned gy = std::min<uint32_t>(G.rows, std::max<uint32_t>(1, (uint32_t)(16 * ctx->sm_count) / (gx * c->world)));
        peer_gather_db_kernel<<<dim3(gx, gy, c->world), 256, 0, ctx->compute>>>(G, reinterpret_cast<uint4 *>(db->tiled));
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            skd_db_free(db);
            return drop(fail(ctx, SKD_ERR_CUDA, "peer_gather_db_kernel: %s", cudaGetErrorString(e)));
        }
        ctx->launches++;
    }
    CU(ctx, cudaEventRecord(c->ev[2], ctx->compute));
    st = finish_op(c);
    if (slice) skd_db_free(slice);
    if (st) {
        skd_db_free(db);
        return st;
    }
    *out = db;
    return SKD_OK;
}

}  // namespace

cudaError_t skdi::preload_peer_kernels() {
    cudaFuncAttributes a;
    cudaError_t e = cudaFuncGetAttributes(&a, (const void *)peer_barrier_kernel);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, (const void *)peer_gather_db_kernel);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, (const void *)peer_copy_kernel);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, (const void *)peer_merge_kernel<true>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, (const void *)peer_merge_kernel<false>);
    return e;
}

// =================================================================================================================
extern "C" {

void skd_triangle_rows_weighted(uint64_t n, int world, const double *shares, uint64_t *bounds) {
    if (world < 1 || !bounds) return;
    double cum = 0.0;
    bounds[0] = 0;
    for (int p = 0; p < world; ++p) {
        cum += shares ? shares[p] : 1.0 / world;
        bounds[p + 1] = p + 1 == world ? n : std::max(bounds[p], triangle_bound(n, cum));
    }
}

void skd_triangle_rows(uint64_t n, int rank, int world, uint64_t *row_begin, uint64_t *row_end) {
    auto bound = [&](int r) -> uint64_t {
        if (r <= 0) return 0;
        if (r >= world) return n;
        const double x = (double)n * (1.0 - std::sqrt(1.0 - (double)r / world));
        return std::min<uint64_t>(n, (uint64_t)std::llround(x / TS) * TS);
    };
    if (row_begin) *row_begin = bound(rank);
    if (row_end) *row_end = bound(rank + 1);
}

void skd_slice_rows(uint64_t n, int rank, int world, uint64_t align, uint64_t *begin, uint64_t *end) {
    if (world < 1) world = 1;
    if (align < 1) align = 1;
    const uint64_t per = ((n + world - 1) / world + align - 1) / align * align;
    const uint64_t a = std::min<uint64_t>(n, per * (uint64_t)std::max(0, rank));
    const uint64_t b = std::min<uint64_t>(n, a + per);
    if (begin) *begin = a;
    if (end) *end = b;
}

size_t skd_comm_window_bytes_knn(uint64_t n_rows, uint64_t knn, int world) {
    return knn_layout(n_rows, knn, std::max(1, world)).end;
}
size_t skd_comm_window_bytes_db(uint64_t n_samples, uint64_t sketchsize64, uint32_t n_kmers, int world) {
    uint64_t a, b;
    skd_slice_rows(n_samples, 0, std::max(1, world), TS, &a, &b);
    return align_up((size_t)n_kmers * sketchsize64 * ((b - a + TS - 1) / TS) * STAGE_BYTES);
}

skd_status skd_comm_create(skd_ctx *ctx, int rank, int world, size_t window_bytes, skd_comm **out) {
    if (!ctx || !out) return fail(ctx, SKD_ERR_INVALID, "null argument");
    *out = nullptr;
    if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world)
        return fail(ctx, SKD_ERR_INVALID, "rank %d / world %d outside [0, %d]", rank, world, kMaxWorld);
    CU(ctx, cudaSetDevice(ctx->device));
    skd_comm *c = new (std::nothrow) skd_comm;
    if (!c) return fail(ctx, SKD_ERR_NOMEM, "out of host memory");
    c->ctx = ctx;
    c->rank = rank;
    c->world = world;
    c->window_bytes = align_up(std::max<size_t>(window_bytes, 256));
    if (const char *e = getenv("SKD_PEER_TIMEOUT_S")) c->timeout_s = std::max(1.0, atof(e));
    if (const char *e = getenv("SKD_ADAPTIVE_SPLIT")) c->adaptive_split = atoi(e) != 0;
    cudaError_t e = cudaMalloc((void **)&c->window, kFlagBytes + c->window_bytes);  // plain cudaMalloc: IPC-exportable
    if (e == cudaSuccess) e = cudaMemset(c->window, 0, kFlagBytes);
    if (e == cudaSuccess) e = cudaHostAlloc((void **)&c->err_host, 64, cudaHostAllocMapped);
    if (e == cudaSuccess) {
        *c->err_host = 0;
        e = cudaHostGetDevicePointer((void **)&c->err_dev, c->err_host, 0);
    }
    for (int i = 0; i < 3 && e == cudaSuccess; ++i) e = cudaEventCreate(&c->ev[i]);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();  // the flags are zero before any peer can see the handle
    if (e != cudaSuccess) {
        skd_status st = fail(ctx, SKD_ERR_CUDA, "peer window of %zu bytes: %s", c->window_bytes, cudaGetErrorString(e));
        skd_comm_destroy(c);
        return st;
    }
    c->peers.base[rank] = c->window;
    c->connected = world == 1;
    *out = c;
    return SKD_OK;
}

void skd_comm_destroy(skd_comm *c) {
    if (!c) return;
    if (c->ctx) {
        cudaSetDevice(c->ctx->device);
        cudaStreamSynchronize(c->ctx->compute);
    }
    for (int p = 0; p < c->world; ++p)
        if (c->opened[p] && c->peers.base[p]) cudaIpcCloseMemHandle(c->peers.base[p]);
    if (c->window) cudaFree(c->window);
    if (c->err_host) cudaFreeHost(c->err_host);
    for (auto &e : c->ev)
        if (e) cudaEventDestroy(e);
    delete c;
}

skd_status skd_comm_export(const skd_comm *c, void *handle_out) {
    if (!c || !handle_out) return SKD_ERR_INVALID;
    skd_ctx *ctx = c->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    HandleBlob b;
    memset(&b, 0, sizeof b);
    CU(ctx, cudaIpcGetMemHandle(&b.ipc, c->window));
    b.ptr = (uint64_t)(uintptr_t)c->window;
    b.window_bytes = c->window_bytes;
    b.device = ctx->device;
    b.pid = (int32_t)getpid();
    b.rank = c->rank;
    b.world = c->world;
    memcpy(handle_out, &b, sizeof b);
    return SKD_OK;
}

skd_status skd_comm_connect(skd_comm *c, const void *handles) {
    if (!c || !handles) return SKD_ERR_INVALID;
    skd_ctx *ctx = c->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    if (c->connected && c->world > 1) return fail(ctx, SKD_ERR_INVALID, "peer group is already connected");
    const HandleBlob *hb = reinterpret_cast<const HandleBlob *>(handles);
    for (int p = 0; p < c->world; ++p) {
        if (p == c->rank) continue;
        HandleBlob b;
        memcpy(&b, &hb[p], sizeof b);
        if (b.rank != p || b.world != c->world)
            return fail(ctx, SKD_ERR_INVALID, "handle %d describes rank %d of %d", p, b.rank, b.world);
        if (b.window_bytes != c->window_bytes)
            return fail(ctx, SKD_ERR_INVALID, "rank %d has a window of %llu bytes, this rank %zu", p,
                        (unsigned long long)b.window_bytes, c->window_bytes);
        if (b.pid == (int32_t)getpid()) {  // a thread of this process: direct peer access
            if (b.device != ctx->device) {
                int can = 0;
                CU(ctx, cudaDeviceCanAccessPeer(&can, ctx->device, b.device));
                if (!can) return fail(ctx, SKD_ERR_UNSUPPORTED, "device %d cannot access device %d", ctx->device, b.device);
                cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled)
                    (void)cudaGetLastError();
                else if (e != cudaSuccess)
                    return fail(ctx, SKD_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d): %s", b.device, cudaGetErrorString(e));
            }
            c->peers.base[p] = reinterpret_cast<unsigned char *>((uintptr_t)b.ptr);
        } else {
            void *ptr = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&ptr, b.ipc, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess)
                return fail(ctx, SKD_ERR_CUDA, "cudaIpcOpenMemHandle for rank %d (device %d, pid %d): %s", p, b.device, b.pid,
                            cudaGetErrorString(e));
            c->peers.base[p] = reinterpret_cast<unsigned char *>(ptr);
            c->opened[p] = true;
        }
    }
    c->connected = true;
    return SKD_OK;
}

int skd_comm_rank(const skd_comm *c) { return c ? c->rank : -1; }
int skd_comm_world(const skd_comm *c) { return c ? c->world : 0; }
float skd_comm_last_op_ms(const skd_comm *c) { return c ? c->last_ms : 0.f; }
float skd_comm_last_exchange_ms(const skd_comm *c) { return c ? c->last_xchg_ms : 0.f; }
void skd_comm_last_row_block(const skd_comm *c, uint64_t *row_begin, uint64_t *row_end) {
    if (row_begin) *row_begin = c ? c->last_r0 : 0;
    if (row_end) *row_end = c ? c->last_r1 : 0;
}

skd_status skd_comm_barrier(skd_comm *c) {
    skd_status st = begin_op(c);
    if (st) return st;
    CU(c->ctx, cudaEventRecord(c->ev[1], c->ctx->compute));
    CU(c->ctx, cudaEventRecord(c->ev[2], c->ctx->compute));
    return finish_op(c);
}

skd_status skd_comm_db_upload(skd_comm *c, const uint64_t *bins_slice, skd_loc bins_loc, uint64_t n_total,
                              uint64_t sample_stride, uint64_t kmer_stride, uint64_t W, uint64_t sketch_size,
                              const uint64_t *kmers, uint32_t K, skd_db **out) {
    if (!c || !out) return SKD_ERR_INVALID;
    skd_ctx *ctx = c->ctx;
    uint64_t s0, s1;
    skd_slice_rows(n_total, c->rank, c->world, TS, &s0, &s1);
    if (s1 > s0 && !bins_slice) return fail(ctx, SKD_ERR_INVALID, "null bins");
    if (c->world == 1)  // a group of one: the slice is the database
        return skd_db_upload(ctx, bins_slice, bins_loc, n_total, sample_stride, kmer_stride, W, sketch_size, kmers, K, out);
    skd_status st = begin_op(c);
    if (st) return st;
    skd_db *slice = nullptr;
    if (s1 > s0) {
        ChunkSource src;
        src.direct = bins_slice;
        src.loc = bins_loc;
        st = db_build(ctx, src, s1 - s0, sample_stride, kmer_stride, W, sketch_size, kmers, K, &slice);
        if (st) return st;
    }
    return gather_db(c, slice, n_total, sample_stride, kmer_stride, W, sketch_size, kmers, K, out);
}

skd_status skd_comm_db_load_skd(skd_comm *c, const char *skd_path, const uint64_t *sample_indices, uint64_t n_total,
                                uint64_t sample_stride, uint64_t kmer_stride, uint64_t W, uint64_t sketch_size,
                                const uint64_t *kmers, uint32_t K, skd_db **out) {
    if (!c || !out || !skd_path) return SKD_ERR_INVALID;
    skd_ctx *ctx = c->ctx;
    uint64_t s0, s1;
    skd_slice_rows(n_total, c->rank, c->world, TS, &s0, &s1);
    if (c->world == 1)
        return skd_db_load_skd(ctx, skd_path, sample_indices, n_total, sample_stride, kmer_stride, W, sketch_size, kmers, K, out);
    skd_status st = begin_op(c);
    if (st) return st;
    skd_db *slice = nullptr;
    if (s1 > s0) {
        std::vector<uint64_t> idx(s1 - s0);
        for (uint64_t s = s0; s < s1; ++s) idx[s - s0] = sample_indices ? sample_indices[s] : s;
        st = skd_db_load_skd(ctx, skd_path, idx.data(), s1 - s0, sample_stride, kmer_stride, W, sketch_size, kmers, K, &slice);
        if (st) return st;
    }
    return gather_db(c, slice, n_total, sample_stride, kmer_stride, W, sketch_size, kmers, K, out);
}

skd_status skd_comm_self_dists_knn(skd_comm *c, const skd_db *db, uint64_t knn, skd_dist_type dist,
                                   const double *completeness, double cutoff, int root, uint64_t *idx_out, float *d0_out,
                                   float *d1_out, skd_loc out_loc) {
    if (!c) return SKD_ERR_INVALID;
    skd_ctx *ctx = c->ctx;
    skd_status st = check_pair(ctx, db, db);
    if (st) return st;
    const uint64_t n = db->n;
    if (n < 2 || knn == 0 || knn > n - 1)
        return fail(ctx, SKD_ERR_INVALID, "knn must be in [1, n-1] (n = %llu, knn = %llu)", (unsigned long long)n,
                    (unsigned long long)knn);
    if (root >= c->world) return fail(ctx, SKD_ERR_INVALID, "root %d outside the group of %d", root, c->world);
    const KnnLayout L = knn_layout(n, knn, c->world);
    if (L.end > c->window_bytes)
        return fail(ctx, SKD_ERR_INVALID, "peer window of %zu bytes is too small: kNN %llu over %llu rows needs %zu "
                    "(skd_comm_window_bytes_knn)", c->window_bytes, (unsigned long long)knn, (unsigned long long)n, L.end);
    EpiParams E;
    st = build_epi(ctx, db, db, dist, completeness, completeness, cutoff, &E);
    if (st) return st;
    st = begin_op(c);
    if (st) return st;
    unsigned char *mine = c->data(c->rank);
    timers_reset(ctx);
    const bool symmetric = knn <= kSymKnnMax && ctx->knn_symmetric;
    uint64_t first_row[kMaxWorld] = {};
    // problem shape the split statistics belong to (never 0: a fresh window reads as tag 0)
    const uint32_t split_tag = (uint32_t)((n * 2654435761ull) ^ (knn << 20) ^ ((uint64_t)E.kind << 28) ^ ((uint64_t)E.nk << 8)) | 1u;
    uint64_t my_pairs = 0;
    if (symmetric) {
        uint64_t bounds[kMaxWorld + 1];
        st = split_triangle(c, n, split_tag, bounds);
        if (st) return st;
        const uint64_t r0 = bounds[c->rank], r1 = bounds[c->rank + 1];
        for (int p = 0; p < c->world; ++p) first_row[p] = bounds[p];
        my_pairs = skd_self_pairs_in_rows(n, r0, std::max(r0, r1));
        c->last_r0 = r0;
        c->last_r1 = std::max(r0, r1);
        st = run_knn_symmetric(ctx, db, E, knn, r0, std::max(r0, r1), reinterpret_cast<unsigned long long *>(mine + L.keys),
                               reinterpret_cast<float *>(mine + L.aux));
    } else {  // every rank computes its owned rows against all columns: final rows straight into the window
        uint64_t o0, o1;
        skd_slice_rows(n, c->rank, c->world, 1, &o0, &o1);
        if (o1 > o0)
            st = run_knn(ctx, db, db, true, E, knn, o0, o1, reinterpret_cast<uint64_t *>(mine + L.fin_idx),
                         reinterpret_cast<float *>(mine + L.fin_d0), reinterpret_cast<float *>(mine + L.fin_d1), SKD_DEVICE);
    }
    if (st) return st;
    st = exchange_rows(c, L, n, knn, E.kind, symmetric, first_row, root, idx_out, d0_out, d1_out, out_loc);
    if (st || !symmetric) return st;
    return leave_split_stat(c, split_tag, my_pairs);
}

skd_status skd_comm_cross_dists_knn(skd_comm *c, const skd_db *ref_shard, const skd_db *query, uint64_t knn,
                                    skd_dist_type dist, const double *ref_c, const double *query_c, double cutoff,
                                    uint64_t ref_index_offset, uint64_t n_ref_total, int root, uint64_t *idx_out,
                                    float *d0_out, float *d1_out, skd_loc out_loc, uint64_t *knn_used) {
    if (!c) return SKD_ERR_INVALID;
    skd_ctx *ctx = c->ctx;
    skd_status st = check_pair(ctx, ref_shard, query);
    if (st) return st;
    if (n_ref_total == 0) return fail(ctx, SKD_ERR_EMPTY_DB, "Reference database has no loaded samples");
    if (query->n == 0) return fail(ctx, SKD_ERR_EMPTY_DB, "Query database has no loaded samples");
    if (knn == 0) return fail(ctx, SKD_ERR_INVALID, "knn must be at least 1");
    if (ref_index_offset + ref_shard->n > n_ref_total || n_ref_total > 0xffffffffull)
        return fail(ctx, SKD_ERR_INVALID, "reference shard [%llu, %llu) outside the %llu references",
                    (unsigned long long)ref_index_offset, (unsigned long long)(ref_index_offset + ref_shard->n),
                    (unsigned long long)n_ref_total);
    if (root >= c->world) return fail(ctx, SKD_ERR_INVALID, "root %d outside the group of %d", root, c->world);
    knn = std::min<uint64_t>(knn, n_ref_total);  // mod.rs:325
    if (knn_used) *knn_used = knn;
    const uint64_t nq = query->n;
    const KnnLayout L = knn_layout(nq, knn, c->world);
    if (L.end > c->window_bytes)
        return fail(ctx, SKD_ERR_INVALID, "peer window of %zu bytes is too small: kNN %llu over %llu rows needs %zu "
                    "(skd_comm_window_bytes_knn)", c->window_bytes, (unsigned long long)knn, (unsigned long long)nq, L.end);
    EpiParams E;  // tile rows are queries (A side), columns refs (B side)
    st = build_epi(ctx, query, ref_shard, dist, query_c, ref_c, cutoff, &E);
    if (st) return st;
    st = begin_op(c);
    if (st) return st;
    unsigned char *mine = c->data(c->rank);
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(mine + L.keys);
    timers_reset(ctx);
    if (ref_shard->n == 0) {  // an empty shard contributes only padding
        launch_fill_u64(keys, nq * knn, ~0ull, ctx->compute);
        ctx->launches++;
    } else {
        st = run_knn(ctx, ref_shard, query, false, E, knn, 0, nq, nullptr, nullptr, nullptr, SKD_DEVICE,
                     reinterpret_cast<uint64_t *>(keys), reinterpret_cast<float *>(mine + L.aux), ref_index_offset);
        if (st) return st;
    }
    return exchange_rows(c, L, nq, knn, E.kind, true, nullptr, root, idx_out, d0_out, d1_out, out_loc);
}

skd_status skd_comm_knn_merge(skd_comm *c, const uint64_t *keys, const float *aux, uint32_t n_parts_local, skd_loc in_loc,
                              uint64_t n_rows, uint64_t knn, skd_dist_type dist, int root, uint64_t *idx_out, float *d0_out,
                              float *d1_out, skd_loc out_loc) {
    if (!c) return SKD_ERR_INVALID;
    skd_ctx *ctx = c->ctx;
    const bool with_aux = dist.kind == SKD_CORE_ACC;
    if (!keys || n_parts_local == 0 || knn == 0 || n_rows == 0) return fail(ctx, SKD_ERR_INVALID, "bad argument");
    if (with_aux && !aux) return fail(ctx, SKD_ERR_INVALID, "core/accessory needs aux");
    if (root >= c->world) return fail(ctx, SKD_ERR_INVALID, "root %d outside the group of %d", root, c->world);
    const KnnLayout L = knn_layout(n_rows, knn, c->world);
    if (L.end > c->window_bytes)
        return fail(ctx, SKD_ERR_INVALID, "peer window of %zu bytes is too small: kNN %llu over %llu rows needs %zu "
                    "(skd_comm_window_bytes_knn)", c->window_bytes, (unsigned long long)knn, (unsigned long long)n_rows, L.end);
    skd_status st = begin_op(c);
    if (st) return st;
    unsigned char *mine = c->data(c->rank);
    unsigned long long *wkeys = reinterpret_cast<unsigned long long *>(mine + L.keys);
    float *waux = reinterpret_cast<float *>(mine + L.aux);
    const size_t per = (size_t)n_rows * knn;
    if (n_parts_local == 1) {
        CU(ctx, cudaMemcpyAsync(wkeys, keys, per * 8, cudaMemcpyDefault, ctx->compute));
        if (with_aux) CU(ctx, cudaMemcpyAsync(waux, aux, per * 4, cudaMemcpyDefault, ctx->compute));
    } else {  // fold this rank's list sets into one (composites kept)
        const unsigned long long *dk = reinterpret_cast<const unsigned long long *>(keys);
        const float *da = aux;
        if (in_loc == SKD_HOST) {
            void *p;
            st = scratch_get(ctx, "merge_key", per * n_parts_local * 8, &p);
            if (st) return st;
            CU(ctx, cudaMemcpyAsync(p, keys, per * n_parts_local * 8, cudaMemcpyHostToDevice, ctx->compute));
            dk = (const unsigned long long *)p;
            if (with_aux) {
                st = scratch_get(ctx, "merge_aux", per * n_parts_local * 4, &p);
                if (st) return st;
                CU(ctx, cudaMemcpyAsync(p, aux, per * n_parts_local * 4, cudaMemcpyHostToDevice, ctx->compute));
                da = (const float *)p;
            }
        }
        launch_fill_u64(wkeys, per, ~0ull, ctx->compute);
        launch_knn_finalize(dk, with_aux ? da : nullptr, n_parts_local, per, 0, n_rows, (uint32_t)knn, dist.kind, nullptr, nullptr,
                            nullptr, wkeys, with_aux ? waux : nullptr, ctx->compute);
        CU(ctx, cudaGetLastError());
        ctx->launches += 2;
    }
    return exchange_rows(c, L, n_rows, knn, dist.kind, true, nullptr, root, idx_out, d0_out, d1_out, out_loc);
}

}  // extern "C"
