// C ABI of libsketchdist.so (include/sketchdist.h): context, sketch-database loader, and the host side of
// the distance drivers (tile schedules, chunked result streaming).  No CPU fallback exists: every compute
// entry point launches the CUDA kernels of kernels.cu or fails.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "internal.h"

using namespace skd;

namespace {
thread_local std::string g_create_error;
}  // namespace

namespace skdi {

skd_status fail(skd_ctx *ctx, skd_status st, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx)
        ctx->err = buf;
    else
        g_create_error = buf;
    return st;
}

// Named device buffers reused across calls.  Stream-ordered allocation (cudaMallocAsync / cudaFreeAsync on the
// compute stream): growing a buffer never synchronises the device, which matters when several contexts share a GPU
// and one of them sits in a peer barrier (peer.cu) -- a cudaFree there would wait for the spinning kernel.
skd_status scratch_get(skd_ctx *ctx, const char *name, size_t bytes, void **out) {
    auto &slot = ctx->scratch[name];
    if (slot.second < bytes) {
        if (slot.first) CU(ctx, cudaFreeAsync(slot.first, ctx->compute));
        slot = {nullptr, 0};
        // large buffers get 1/8 of headroom: the sizes of the key-matrix chunks follow the row blocks of a call, and a
        // peer group re-splits those slightly from call to call -- regrowing an 8 GB buffer stalls a rank for ~50 ms
        size_t want = std::max<size_t>(bytes, 256);
        if (want >= (64u << 20)) want += want / 8;
        CU(ctx, cudaMallocAsync(&slot.first, want, ctx->compute));
        slot.second = want;
    }
    *out = slot.first;
    return SKD_OK;
}

skd_status pinned_get(skd_ctx *ctx, size_t bytes) {
    if (ctx->pinned_bytes >= bytes) return SKD_OK;
    for (int i = 0; i < 2; ++i) {
        if (ctx->pinned[i]) cudaFreeHost(ctx->pinned[i]);
        ctx->pinned[i] = nullptr;
    }
    ctx->pinned_bytes = 0;
    for (int i = 0; i < 2; ++i) CU(ctx, cudaMallocHost(&ctx->pinned[i], bytes));
    ctx->pinned_bytes = bytes;
    return SKD_OK;
}

void timers_reset(skd_ctx *ctx) { ctx->timers_used = 0; }
skd_status timer_begin(skd_ctx *ctx, cudaStream_t stream, int tag) {
    if (ctx->timers_used == ctx->timers.size()) {
        cudaEvent_t a, b;
        CU(ctx, cudaEventCreate(&a));
        CU(ctx, cudaEventCreate(&b));
        ctx->timers.push_back({a, b});
        ctx->timer_tags.push_back(0);
    }
    ctx->timer_tags[ctx->timers_used] = tag;
    CU(ctx, cudaEventRecord(ctx->timers[ctx->timers_used].first, stream ? stream : ctx->compute));
    return SKD_OK;
}
skd_status timer_end(skd_ctx *ctx, cudaStream_t stream) {
    CU(ctx, cudaEventRecord(ctx->timers[ctx->timers_used].second, stream ? stream : ctx->compute));
    ctx->timers_used++;
    return SKD_OK;
}
// Call after the compute stream is idle.
skd_status timers_collect(skd_ctx *ctx) {
    float total = 0.f, tile = 0.f;
    for (size_t i = 0; i < ctx->timers_used; ++i) {
        float ms = 0.f;
        CU(ctx, cudaEventElapsedTime(&ms, ctx->timers[i].first, ctx->timers[i].second));
        total += ms;
        if (ctx->timer_tags[i] == 1) tile += ms;
    }
    ctx->last_ms = total;
    ctx->last_tile_ms = tile;
    return SKD_OK;
}

uint64_t condensed(uint64_t i, uint64_t j, uint64_t n) {  // square_to_condensed, distance_matrix.rs:11-14
    return n * i - ((i * (i + 1)) >> 1) + j - 1 - i;
}

// ---- epilogue parameter block --------------------------------------------------------------------
// Host side of jaccard_index's float tail (jaccard.rs:26-33) for the tables.
double host_jaccard(uint32_t c, uint32_t nbins) {
    uint32_t expected = nbins >> BBITS;
    uint32_t diff = c > expected ? c - expected : 0;
    double intersize = ((double)diff * (double)nbins) / (double)(nbins - expected);
    return intersize / (double)nbins;
}

skd_status build_epi(skd_ctx *ctx, const skd_db *A, const skd_db *B, skd_dist_type dist, const double *a_comp,
                     const double *b_comp, double cutoff, EpiParams *E) {
    memset(E, 0, sizeof *E);
    const uint32_t nbins = (uint32_t)(64 * A->W);
    const int K = (int)A->kmers.size();
    E->kind = dist.kind;
    E->nbins = nbins;
    E->expected = nbins >> BBITS;
    E->cutoff = cutoff;
    E->tolerance = std::log(2.0 / (double)(A->sketch_size * 64));  // jaccard.rs:75 (sketch_size = bins)
    E->lny = A->lny_dev;
    if (dist.kind == KIND_CORE_ACC) {
        if (K < 2)
            return fail(ctx, SKD_ERR_NEED_TWO_K, "Need at least two k-mer lengths to calculate core/accessory distances");
        if (K > MAXK) return fail(ctx, SKD_ERR_UNSUPPORTED, "core/accessory supports at most %d k-mer lengths", MAXK);
        E->k0 = 0;
        E->nk = K;
        double xsum = 0.0, xsq = 0.0;
        for (int n = 0; n <= K; ++n) {
            if (n > 0) {  // running sums in the reference's order (jaccard.rs:92-97)
                double kf = (double)A->kmers[n - 1];
                E->kf[n - 1] = kf;
                xsum += kf;
                xsq += kf * kf;
            }
            E->xsum[n] = xsum;
            if (n > 0) {  // jaccard.rs:121-125
                double dn = (double)n;
                E->xbar[n] = xsum / dn;
                E->x_diff[n] = xsq - xsum * xsum / dn;
                E->xstddev[n] = std::sqrt((xsq - xsum * xsum / dn) / dn);
                E->rcp_n[n] = 1.0 / dn;
            }
        }
    } else if (dist.kind == KIND_JACCARD || dist.kind == KIND_ANI) {
        if (dist.k_idx < 0 || dist.k_idx >= K) return fail(ctx, SKD_ERR_INVALID, "k_idx %d out of range", dist.k_idx);
        E->k0 = dist.k_idx;
        E->nk = 1;
        E->k_single = (double)A->kmers[dist.k_idx];
        // value/key tables by match count, computed with the host libm like the reference does per pair
        std::vector<float> tab(2 * ((size_t)nbins + 1));
        for (uint32_t c = 0; c <= nbins; ++c) {
            double j = host_jaccard(c, nbins);
            if (dist.kind == KIND_ANI) {
                double ani = std::fmax(0.0, 1.0 + 1.0 / E->k_single * std::log((2.0 * j) / (1.0 + j)));
                tab[c] = (float)ani;                       // mod.rs:97-98
                tab[nbins + 1 + c] = (float)(1.0 - ani);   // mod.rs:175
            } else {
                tab[c] = (float)(1.0 - j);                 // mod.rs:100
                tab[nbins + 1 + c] = (float)(1.0 - j);     // mod.rs:177
            }
        }
        void *dtab;
        skd_status st = scratch_get(ctx, "valtab", tab.size() * sizeof(float), &dtab);
        if (st) return st;
        CU(ctx, cudaMemcpyAsync(dtab, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->compute));
        CU(ctx, cudaStreamSynchronize(ctx->compute));  // `tab` is stack-scoped pageable memory
        E->val_tab = (const float *)dtab;
        E->key_tab = (const float *)dtab + nbins + 1;
    } else {
        return fail(ctx, SKD_ERR_INVALID, "unknown distance kind %d", dist.kind);
    }
    E->kp = counts_per_pair(E->nk, nbins >= 65536);
    // Completeness applies only when both vectors are given (jaccard.rs:36)
    if (a_comp && b_comp) {
        E->has_c = 1;
        void *da, *dbp;
        skd_status st = scratch_get(ctx, "acomp", std::max<uint64_t>(1, A->n) * sizeof(double), &da);
        if (st) return st;
        st = scratch_get(ctx, "bcomp", std::max<uint64_t>(1, B->n) * sizeof(double), &dbp);
        if (st) return st;
        CU(ctx, cudaMemcpyAsync(da, a_comp, A->n * sizeof(double), cudaMemcpyHostToDevice, ctx->compute));
        CU(ctx, cudaMemcpyAsync(dbp, b_comp, B->n * sizeof(double), cudaMemcpyHostToDevice, ctx->compute));
        CU(ctx, cudaStreamSynchronize(ctx->compute));
        E->a_comp = (const double *)da;
        E->b_comp = (const double *)dbp;
    }
    return SKD_OK;
}

skd_status materialize(skd_ctx *ctx, const skd_db *cdb);

skd_status check_pair(skd_ctx *ctx, const skd_db *A, const skd_db *B, bool keep_pending) {
    if (!ctx || !A || !B) return fail(ctx, SKD_ERR_INVALID, "null context or database");
    if (A->ctx != ctx || B->ctx != ctx) return fail(ctx, SKD_ERR_INVALID, "database belongs to another context");
    if (A != B && !skd_db_compatible(A, B))
        return fail(ctx, SKD_ERR_INCOMPATIBLE, "reference and query databases differ in sketch size or k-mer lengths");
    if (!keep_pending) {  // an asynchronously uploaded database is re-tiled before anything but a dense self run
        skd_status st = materialize(ctx, A);
        if (st) return st;
        if (B != A) st = materialize(ctx, B);
        if (st) return st;
    }
    return SKD_OK;
}

// Group offsets for a dense launch over row blocks [rb0, rb0 + n_rb).
// `pairs`: the launch goes to dense_kernel, whose tiles span two adjacent column blocks.
void build_groups(uint32_t group, bool self, uint32_t rb0, uint32_t n_rb, uint32_t nbB, std::vector<uint64_t> *off,
                  bool pairs, uint32_t cb_lo, uint32_t cb_hi) {
    uint32_t n_groups = (n_rb + group - 1) / group;
    off->assign(n_groups + 1, 0);
    const uint32_t hi = cb_hi ? std::min(cb_hi, nbB) : nbB;
    for (uint32_t g = 0; g < n_groups; ++g) {
        uint32_t rows = std::min(group, n_rb - g * group);
        const uint32_t first = std::max(self ? rb0 + g * group : 0u, cb_lo);  // the kernel's cb_first
        uint64_t cols = hi > first ? hi - first : 0;
        if (pairs) cols = (cols + 1) / 2;
        (*off)[g + 1] = (*off)[g] + (uint64_t)rows * cols;
    }
}

// Dense launches (all 64 columns of a tile in use) go to the persistent dense_kernel unless SKD_DENSE_V1 is set.
// Narrow column sides (<= 32 samples) keep the narrow tile_kernel variants -- unless there are so many k-mer lengths
// that a tile's counts do not fit shared memory: dense_kernel keeps them in L2 and takes any number.
bool uses_dense_kernel(const skd_ctx *ctx, int mode, const skd_db *B, const EpiParams &E) {
    if (!ctx->dense_v2 || (mode != MODE_DENSE_SELF && mode != MODE_DENSE_CROSS)) return false;
    if (tile_kernel_smem(E.nk, E.nbins >= 65536, false, 0, false) > kMaxSmem) return true;
    // dense_kernel's tiles span two column blocks: with an odd, small number of them (cross runs against 33..64,
    // 129..192, ... queries) the empty half of the last pair would idle half the warps of 1 / (nb + 1) of the tiles
    const uint32_t nb = B->dev.nb;
    return mode == MODE_DENSE_SELF ? B->n > 64 : (nb >= 2 && (nb % 2 == 0 || nb >= 16));
}

// Re-tile chunk c of an asynchronously uploaded database (compute stream waits for its arrival).
skd_status retile_chunk(skd_ctx *ctx, skd_db *db, size_t c, cudaStream_t stream = nullptr) {
    auto &p = db->pending;
    if (!stream) stream = ctx->compute;
    const uint32_t b0 = c ? p.block_end[c - 1] : 0, b1 = p.block_end[c];
    const uint64_t s0 = (uint64_t)b0 * TS, s1 = std::min<uint64_t>(db->n, (uint64_t)b1 * TS);
    CU(ctx, cudaStreamWaitEvent(stream, p.ready[c], 0));
    launch_repack(p.raw + s0 * p.sample_stride, s1 - s0, p.sample_stride, p.kmer_stride, db->tiled, db->dev, b0, b1 - b0,
                  stream);
    CU(ctx, cudaGetLastError());
    ctx->launches++;
    return SKD_OK;
}
skd_status finish_pending(skd_ctx *ctx, skd_db *db) {
    auto &p = db->pending;
    for (auto e : p.ready) cudaEventDestroy(e);
    p.ready.clear();
    p.block_end.clear();
    if (p.raw) CU(ctx, cudaFreeAsync(p.raw, ctx->compute));
    p.raw = nullptr;
    return SKD_OK;
}
// Every compute entry point that does not interleave calls this first.
skd_status materialize(skd_ctx *ctx, const skd_db *cdb) {
    skd_db *db = const_cast<skd_db *>(cdb);
    if (!db || !db->is_pending()) return SKD_OK;
    for (size_t c = 0; c < db->pending.ready.size(); ++c) {
        skd_status st = retile_chunk(ctx, db, c);
        if (st) return st;
    }
    return finish_pending(ctx, db);
}

// Match-count scratch of the persistent dense kernel, one per stream that can carry dense launches (the async-upload
// path alternates column bands between the compute stream and aux[1]).  Allocated on the compute stream: callers that
// launch on aux[1] make that stream wait for the compute stream first (run_dense does).
skd_status dense_scratch(skd_ctx *ctx, const EpiParams &E, int slot, void **out) {
    *out = nullptr;
    if (!ctx->dense_v2) return SKD_OK;
    return scratch_get(ctx, slot ? "tile_counts1" : "tile_counts0",
                       dense_kernel_scratch_bytes(E.kp, E.nbins >= 65536, ctx->sm_count), out);
}
// The early exit of dense_kernel (core/accessory, more than three k-mer lengths): list for the deferred pairs.
skd_status ensure_rows(skd_ctx *ctx, const skd_db *cdb) {
    skd_db *db = const_cast<skd_db *>(cdb);
    if (db->rows || db->n == 0 || db->is_pending()) return SKD_OK;
    const size_t bytes = (size_t)db->n * db->kmers.size() * db->W * BBITS * sizeof(uint64_t);
    CU(ctx, cudaMallocAsync((void **)&db->rows, bytes, ctx->compute));
    launch_untile(db->dev, db->rows, ctx->compute);
    CU(ctx, cudaGetLastError());
    ctx->launches++;
    return SKD_OK;
}
skd_status defer_setup(skd_ctx *ctx, const skd_db *A, const skd_db *B, const EpiParams &E, int slot, SchedParams *S) {
    S->defer_list = nullptr;
    S->defer_cap = S->defer_max = 0;
    S->a_rows = S->b_rows = nullptr;
    S->work_stats = ctx->work_stats;
    if (!ctx->early_out || !S->dense_pairs || E.kind != KIND_CORE_ACC || E.nk <= 3) return SKD_OK;
    if (A->is_pending() || B->is_pending() || slot != 0) return SKD_OK;  // column bands of an upload still in flight
    skd_status st = ensure_rows(ctx, A);
    if (st) return st;
    if ((st = ensure_rows(ctx, B))) return st;
    void *p;
    st = scratch_get(ctx, "defer_list0", (size_t)ctx->defer_cap * 2 * sizeof(uint4), &p);
    if (st) return st;
    S->defer_list = (uint4 *)p;
    S->defer_cap = ctx->defer_cap;
    S->defer_max = ctx->defer_max;
    S->a_rows = A->rows;
    S->b_rows = B->rows;
    return SKD_OK;
}

// Launch the tile kernel for rows [r0, r1) writing to device memory `dout`.  `d_grp_off` is the device copy
// of `off` (group offsets built by build_groups for exactly this row range).
skd_status launch_dense(skd_ctx *ctx, const skd_db *A, const skd_db *B, const EpiParams &E, int mode, uint64_t r0,
                        uint64_t r1, float *dout, const std::vector<uint64_t> &off, const uint64_t *d_grp_off,
                        uint64_t rect_ld, uint32_t cb_lo, uint32_t cb_hi, uint64_t out_row_begin, cudaStream_t stream) {
    if (!stream) stream = ctx->compute;
    SchedParams S{};
    S.mode = mode;
    S.one = 1;
    S.self_mode = mode == MODE_DENSE_SELF;
    S.rb0 = (uint32_t)(r0 / TS);
    S.n_rb = (uint32_t)((r1 + TS - 1) / TS) - S.rb0;
    S.group = ctx->group;
    S.n_groups = (uint32_t)off.size() - 1;
    S.grp_off = d_grp_off;
    S.row_begin = r0;
    S.row_end = r1;
    // `dout` starts at the first pair of row out_row_begin (default: r0)
    const uint64_t ob = out_row_begin == ~0ull ? r0 : out_row_begin;
    S.out_base = S.self_mode ? condensed(ob, ob + 1, A->n) : 0;
    S.out = dout;
    S.rect_ld = rect_ld;
    S.cb_lo = cb_lo;
    S.cb_hi = cb_hi;
    const bool count32 = E.nbins >= 65536;
    size_t smem = tile_kernel_smem(E.nk, count32, false, 0, false);
    skd_status st = SKD_OK;
    // the count scratch: dense_kernel's tiles, or the one-pass stream kernel for 9..16 queries (kernels.cu)
    S.dense_pairs = uses_dense_kernel(ctx, mode, B, E);
    if (S.dense_pairs || (mode == MODE_DENSE_CROSS && B->n > 8 && B->n <= 16 && !count32)) {
        st = dense_scratch(ctx, E, stream == ctx->aux[1] ? 1 : 0, &S.counts_scratch);
        if (st) return st;
    }
    st = defer_setup(ctx, A, B, E, stream == ctx->compute ? 0 : 1, &S);
    if (st) return st;
    // tile_kernel keeps a tile's counts in shared memory; dense_kernel keeps them in L2
    if (!S.dense_pairs && smem > kMaxSmem)
        return fail(ctx, SKD_ERR_UNSUPPORTED, "%d k-mer lengths at %u bins need %zu B of shared memory per tile", E.nk,
                    E.nbins, smem);
    st = timer_begin(ctx, stream, 1);
    if (st) return st;
    CU(ctx, launch_tiles(A->dev, B->dev, E, S, off.back(), count32, smem, stream));
    ctx->launches++;
    return timer_end(ctx, stream);
}

// Rows [row_begin, row_end) of a dense matrix, streamed to host in chunks or written in place on the device.
skd_status run_dense(skd_ctx *ctx, const skd_db *A, const skd_db *B, bool self, const EpiParams &E, int mode,
                     uint64_t row_begin, uint64_t row_end, float *out, skd_loc out_loc) {
    const uint64_t ncols = mode == MODE_COUNTS ? (uint64_t)E.nk : (E.kind == KIND_CORE_ACC ? 2 : 1);
    auto elems_before = [&](uint64_t r) -> uint64_t {  // output elements (pairs) of rows [row_begin, r)
        if (self) return skd_self_pairs_in_rows(A->n, row_begin, r);
        return (r - row_begin) * B->n;
    };
    timers_reset(ctx);
    // A pinned (page-locked, mapped) host destination can be written by the kernel directly: the result stores
    // stream over PCIe while later tiles are still being computed, with no staging copy at all.
    if (out_loc == SKD_HOST && ctx->zero_copy) {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, out) == cudaSuccess && attr.type == cudaMemoryTypeHost &&
            attr.devicePointer != nullptr) {
            out = (float *)attr.devicePointer;
            out_loc = SKD_DEVICE;
        } else {
            (void)cudaGetLastError();
        }
    }
    // Plan: one chunk when the result stays on the device; otherwise row chunks of about ctx->result_chunk_bytes
    // on two device buffers, so that the D2H copy of chunk c overlaps the kernel of chunk c+1.
    struct Chunk {
        uint64_t r0, r1;
        std::vector<uint64_t> off;
        size_t off_pos;
    };
    std::vector<Chunk> chunks;
    uint64_t rows_per_chunk = row_end - row_begin + TS;
    if (out_loc == SKD_HOST) {
        const uint64_t widest = self ? (A->n - 1 - std::min(row_begin, A->n - 1)) : B->n;
        rows_per_chunk = ctx->result_chunk_bytes / std::max<uint64_t>(1, widest * ncols * 4);
        rows_per_chunk = std::max<uint64_t>(TS, rows_per_chunk / TS * TS);
    }
    size_t chunk_elems = 0, off_total = 0;
    for (uint64_t r = row_begin; r < row_end;) {
        Chunk c;
        c.r0 = r;
        c.r1 = std::min(row_end, (r / TS) * TS + rows_per_chunk);
        build_groups(ctx->group, self, (uint32_t)(c.r0 / TS), (uint32_t)((c.r1 + TS - 1) / TS) - (uint32_t)(c.r0 / TS), B->dev.nb,
                     &c.off, uses_dense_kernel(ctx, mode, B, E));
        c.off_pos = off_total;
        off_total += c.off.size();
        chunk_elems = std::max<size_t>(chunk_elems, elems_before(c.r1) - elems_before(c.r0));
        r = c.r1;
        chunks.push_back(std::move(c));
    }
    std::vector<uint64_t> all_off;
    all_off.reserve(off_total);
    for (auto &c : chunks) all_off.insert(all_off.end(), c.off.begin(), c.off.end());
    void *doff;
    skd_status st = scratch_get(ctx, "grp_off", all_off.size() * sizeof(uint64_t), &doff);
    if (st) return st;
    CU(ctx, cudaMemcpyAsync(doff, all_off.data(), all_off.size() * sizeof(uint64_t), cudaMemcpyHostToDevice,
                            ctx->compute));
    CU(ctx, cudaStreamSynchronize(ctx->compute));  // all_off is pageable
    const uint64_t *d_off = (const uint64_t *)doff;

    if (out_loc == SKD_DEVICE) {
        skd_db *pa = const_cast<skd_db *>(A);
        if (self && mode == MODE_DENSE_SELF && pa->is_pending()) {
            // The database is still arriving: re-tile each chunk as it lands and immediately run the tiles whose
            // column block lies in it (their row blocks arrived earlier), so PCIe and the tile kernel overlap.
            const size_t nc = pa->pending.ready.size();
            std::vector<std::vector<uint64_t>> offs(nc);
            std::vector<size_t> pos(nc);
            std::vector<uint64_t> flat;
            const uint32_t rb0 = (uint32_t)(row_begin / TS);
            for (size_t c = 0; c < nc; ++c) {
                const uint32_t b0 = c ? pa->pending.block_end[c - 1] : 0, b1 = pa->pending.block_end[c];
                const uint64_t r1 = std::min<uint64_t>(row_end, (uint64_t)b1 * TS);
                pos[c] = flat.size();
                if (r1 > row_begin) {
                    build_groups(ctx->group, true, rb0, (uint32_t)((r1 + TS - 1) / TS) - rb0, A->dev.nb, &offs[c],
                                 uses_dense_kernel(ctx, mode, A, E), b0, b1);
                    flat.insert(flat.end(), offs[c].begin(), offs[c].end());
                }
            }
            void *dband;
            st = scratch_get(ctx, "grp_band", std::max<size_t>(1, flat.size()) * sizeof(uint64_t), &dband);
            if (st) return st;
            CU(ctx, cudaMemcpyAsync(dband, flat.data(), flat.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->compute));
            CU(ctx, cudaStreamSynchronize(ctx->compute));  // flat is pageable
            // Streams: aux[0] re-tiles the chunks in arrival order; the column bands alternate between the compute
            // stream and aux[1], each waiting only for "its" chunk to be re-tiled, so that the tiles of band c + 1 fill
            // the SMs while band c drains (on one stream every band boundary costs a partial wave).
            std::vector<cudaEvent_t> retiled(nc, nullptr);
            auto drop_events = [&] {
                for (auto e : retiled)
                    if (e) cudaEventDestroy(e);
            };
            for (int slot = 0; slot < 2; ++slot) {  // both streams' count scratch exists before the helpers start
                void *unused;
                st = dense_scratch(ctx, E, slot, &unused);
                if (st) return st;
            }
            CU(ctx, cudaEventRecord(ctx->ev_a[0], ctx->compute));  // database buffers exist, offsets uploaded
            CU(ctx, cudaStreamWaitEvent(ctx->aux[0], ctx->ev_a[0], 0));
            CU(ctx, cudaStreamWaitEvent(ctx->aux[1], ctx->ev_a[0], 0));
            for (size_t c = 0; c < nc; ++c) {
                st = retile_chunk(ctx, pa, c, ctx->aux[0]);
                if (st == SKD_OK && cudaEventCreateWithFlags(&retiled[c], cudaEventDisableTiming) != cudaSuccess)
                    st = fail(ctx, SKD_ERR_CUDA, "cudaEventCreate failed");
                if (st == SKD_OK && cudaEventRecord(retiled[c], ctx->aux[0]) != cudaSuccess)
                    st = fail(ctx, SKD_ERR_CUDA, "cudaEventRecord failed");
                if (st) {
                    cudaDeviceSynchronize();
                    drop_events();
                    return st;
                }
            }
            for (size_t c = 0; c < nc; ++c) {
                cudaStream_t band = (c & 1) ? ctx->aux[1] : ctx->compute;
                if (cudaError_t e = cudaStreamWaitEvent(band, retiled[c], 0); e != cudaSuccess) {
                    cudaDeviceSynchronize();
                    drop_events();
                    return fail(ctx, SKD_ERR_CUDA, "cudaStreamWaitEvent failed: %s", cudaGetErrorString(e));
                }
                if (offs[c].empty() || offs[c].back() == 0) continue;
                const uint32_t b0 = c ? pa->pending.block_end[c - 1] : 0, b1 = pa->pending.block_end[c];
                const uint64_t r1 = std::min<uint64_t>(row_end, (uint64_t)b1 * TS);
                st = launch_dense(ctx, A, B, E, mode, row_begin, r1, out, offs[c], (const uint64_t *)dband + pos[c], 0, b0, b1,
                                  row_begin, band);
                if (st) {
                    cudaDeviceSynchronize();
                    drop_events();
                    return st;
                }
            }
            for (int i = 0; i < 2; ++i) {  // join: the compute stream continues after both helpers
                cudaError_t e = cudaEventRecord(ctx->ev_b[i], ctx->aux[i]);
                if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->compute, ctx->ev_b[i], 0);
                if (e != cudaSuccess) {
                    cudaDeviceSynchronize();
                    drop_events();
                    return fail(ctx, SKD_ERR_CUDA, "joining the band streams failed: %s", cudaGetErrorString(e));
                }
            }
            CU(ctx, cudaStreamSynchronize(ctx->compute));
            drop_events();
            st = finish_pending(ctx, pa);
            if (st) return st;
            CU(ctx, cudaStreamSynchronize(ctx->compute));
            return timers_collect(ctx);
        }
        for (auto &c : chunks) {
            st = launch_dense(ctx, A, B, E, mode, c.r0, c.r1, out + elems_before(c.r0) * ncols, c.off, d_off + c.off_pos);
            if (st) return st;
        }
        CU(ctx, cudaStreamSynchronize(ctx->compute));
        return timers_collect(ctx);
    }
    {   // staged host path: the database must be complete
        st = materialize(ctx, A);
        if (st) return st;
    }
    void *dbuf[2];
    st = scratch_get(ctx, "dense0", chunk_elems * ncols * 4, &dbuf[0]);
    if (st) return st;
    dbuf[1] = dbuf[0];
    if (chunks.size() > 1) {
        st = scratch_get(ctx, "dense1", chunk_elems * ncols * 4, &dbuf[1]);
        if (st) return st;
    }
    for (size_t ci = 0; ci < chunks.size(); ++ci) {
        const Chunk &c = chunks[ci];
        const int b = (int)(ci & 1);
        if (ci >= 2) CU(ctx, cudaStreamWaitEvent(ctx->compute, ctx->ev_b[b], 0));  // D2H of chunk ci-2 done
        st = launch_dense(ctx, A, B, E, mode, c.r0, c.r1, (float *)dbuf[b], c.off, d_off + c.off_pos);
        if (st) return st;
        CU(ctx, cudaEventRecord(ctx->ev_a[b], ctx->compute));
        CU(ctx, cudaStreamWaitEvent(ctx->copy, ctx->ev_a[b], 0));
        const uint64_t e0 = elems_before(c.r0), e1 = elems_before(c.r1);
        CU(ctx, cudaMemcpyAsync(out + e0 * ncols, dbuf[b], (e1 - e0) * ncols * 4, cudaMemcpyDeviceToHost, ctx->copy));
        CU(ctx, cudaEventRecord(ctx->ev_b[b], ctx->copy));
    }
    CU(ctx, cudaStreamSynchronize(ctx->copy));
    CU(ctx, cudaStreamSynchronize(ctx->compute));
    return timers_collect(ctx);
}

// Whole-range cross kNN (knn <= 64): the dense cross kernel (references as rows; stream kernel up to 16 queries,
// narrow-column tile variants up to 32) writes the kNN keys ref-major into HBM, one chunk of references at a time, and
// select_few_kernel folds each query's column into per-slice lists that are carried from chunk to chunk and merged at
// the end.  With few queries the row-owned kernel would have only a few row blocks (queries are its rows); with many
// it still loses 5 % to its in-kernel candidate lists.
// E is oriented for the row-owned call (a = queries, b = refs).

skd_status run_knn_few_queries(skd_ctx *ctx, const skd_db *R, const skd_db *Q, const EpiParams &E, uint64_t knn,
                               uint64_t *idx_out, float *d0_out, float *d1_out, skd_loc out_loc, uint64_t *key_out,
                               float *aux_out, uint64_t idx_offset) {
    const uint64_t nq = Q->n, n_ref = R->n;
    const bool with_aux = E.kind == KIND_CORE_ACC, partial = key_out != nullptr;
    const uint64_t elt = with_aux ? 8 : 4;
    EpiParams Ed = E;  // dense launch: rows = refs, columns = queries
    std::swap(Ed.a_comp, Ed.b_comp);
    uint64_t rows_chunk = std::max<uint64_t>(TS, ctx->sym_chunk_bytes / (nq * elt) / TS * TS);
    rows_chunk = std::min<uint64_t>(rows_chunk, (n_ref + TS - 1) / TS * TS);
    const uint32_t n_chunks = (uint32_t)((n_ref + rows_chunk - 1) / rows_chunk);
    const uint32_t tgt_groups = (uint32_t)((nq + 7) / 8);  // select_few_kernel: 8 queries per CTA
    uint32_t n_splits = (uint32_t)std::max<uint64_t>(1, (4ull * ctx->sm_count) / (uint64_t)tgt_groups);
    // at most 16 row splits: few, long lists, since the final merge is O(parts^2) per query
    n_splits = std::max(1u, std::min<uint32_t>({n_splits, 16u, (uint32_t)(rows_chunk / TS)}));
    const uint32_t rows_per_split = (uint32_t)(((rows_chunk + n_splits - 1) / n_splits + TS - 1) / TS * TS);
    const uint32_t parts = n_splits;  // slice s of every chunk continues the lists of slice s of the chunk before
    const uint64_t part_stride = nq * knn;
    void *D, *pk, *pa = nullptr;
    skd_status st = scratch_get(ctx, "knn_rect", rows_chunk * nq * elt, &D);
    if (st) return st;
    st = scratch_get(ctx, "few_key", (size_t)parts * part_stride * 8, &pk);
    if (st) return st;
    if (with_aux) {
        st = scratch_get(ctx, "few_aux", (size_t)parts * part_stride * 4, &pa);
        if (st) return st;
    }
    timers_reset(ctx);
    launch_fill_u64((unsigned long long *)pk, (uint64_t)parts * part_stride, ~0ull, ctx->compute);
    ctx->launches++;
    for (uint32_t c = 0; c < n_chunks; ++c) {
        const uint64_t g0 = (uint64_t)c * rows_chunk, g1 = std::min(n_ref, g0 + rows_chunk);
        std::vector<uint64_t> off;
        build_groups(ctx->group, false, (uint32_t)(g0 / TS), (uint32_t)((g1 + TS - 1) / TS) - (uint32_t)(g0 / TS), Q->dev.nb, &off,
                     uses_dense_kernel(ctx, MODE_DENSE_CROSS, Q, Ed));
        void *doff;
        st = scratch_get(ctx, "grp_off", off.size() * sizeof(uint64_t), &doff);
        if (st) return st;
        CU(ctx, cudaMemcpyAsync(doff, off.data(), off.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->compute));
        CU(ctx, cudaStreamSynchronize(ctx->compute));  // off is local; also fences reuse of the key matrix
        st = launch_dense(ctx, R, Q, Ed, MODE_DENSE_CROSS, g0, g1, (float *)D, off, (const uint64_t *)doff, nq);
        if (st) return st;
        st = timer_begin(ctx);
        if (st) return st;
        CU(ctx, launch_select_few(with_aux, (const float *)D, nq, g0, (uint32_t)(g1 - g0), (uint32_t)nq,
                                  (unsigned long long *)pk, with_aux ? (float *)pa : nullptr, (uint32_t)knn, ctx->compute,
                                  idx_offset, n_splits, rows_per_split, part_stride));
        ctx->launches++;
        st = timer_end(ctx);
        if (st) return st;
    }
    if (partial) {
        unsigned long long *dk = (unsigned long long *)key_out;
        float *da = aux_out;
        if (out_loc == SKD_HOST) {
            void *p;
            st = scratch_get(ctx, "knn_idx", nq * knn * 8, &p);
            if (st) return st;
            dk = (unsigned long long *)p;
            st = scratch_get(ctx, "knn_d0", nq * knn * 4, &p);
            if (st) return st;
            da = (float *)p;
        }
        launch_fill_u64(dk, nq * knn, ~0ull, ctx->compute);
        launch_knn_finalize((const unsigned long long *)pk, (const float *)pa, parts, part_stride, 0, nq, (uint32_t)knn,
                            E.kind, nullptr, nullptr, nullptr, dk, with_aux ? da : nullptr, ctx->compute);
        CU(ctx, cudaGetLastError());
        ctx->launches += 2;
        if (out_loc == SKD_HOST) {
            CU(ctx, cudaMemcpyAsync(key_out, dk, nq * knn * 8, cudaMemcpyDeviceToHost, ctx->compute));
            if (aux_out && with_aux) CU(ctx, cudaMemcpyAsync(aux_out, da, nq * knn * 4, cudaMemcpyDeviceToHost, ctx->compute));
        }
        CU(ctx, cudaStreamSynchronize(ctx->compute));
        return timers_collect(ctx);
    }
    st = finalize_lists(ctx, (const unsigned long long *)pk, (const float *)pa, parts, part_stride, nq, knn, E.kind, idx_out,
                        d0_out, d1_out, out_loc);
    if (st) return st;
    return timers_collect(ctx);
}

// Row-owned kNN: the tile kernel scans every column for rows [row_begin, row_end).  Output is either the final
// sparse rows (idx/d0/d1) or, when key_out is set, one sorted composite list per row (a partial result for a
// merge across column shards; idx_offset is added to the stored neighbour indices).
skd_status run_knn(skd_ctx *ctx, const skd_db *R, const skd_db *Q, bool self, const EpiParams &E, uint64_t knn,
                   uint64_t row_begin, uint64_t row_end, uint64_t *idx_out, float *d0_out, float *d1_out,
                   skd_loc out_loc, uint64_t *key_out, float *aux_out, uint64_t idx_offset) {
    // Rows are the A side of the tile kernel (self: the database; cross: queries), columns the refs.
    const skd_db *A = self ? R : Q, *B = R;
    const uint64_t n_rows = row_end - row_begin;
    timers_reset(ctx);
    if (n_rows == 0) return SKD_OK;
    if (!self && ctx->knn_symmetric && Q->n <= ctx->few_queries && knn <= kSymKnnMax && row_begin == 0 && row_end == Q->n)
        return run_knn_few_queries(ctx, R, Q, E, knn, idx_out, d0_out, d1_out, out_loc, key_out, aux_out, idx_offset);
    const bool with_aux = E.kind == KIND_CORE_ACC;
    const bool partial = key_out != nullptr;
    uint32_t cap = 128;
    while (cap < knn + TS) cap <<= 1;
    const bool count32 = E.nbins >= 65536;
    size_t smem = tile_kernel_smem(E.nk, count32, true, cap, with_aux);
    // Beyond what the list-keeping kernel holds per row in shared memory (knn > 1984), or with so many k-mer lengths
    // that a tile's counts do not fit: dense keys + a full sort of every row (sortknn.cu).  The reference has no such
    // limits (--knn is clamped to n - 1, src/lib.rs:379-382; any number of k-mer lengths, jaccard.rs:77).
    if ((size_t)cap * 12 > (size_t)NST * 2 * STAGE_BYTES || smem > kMaxSmem)
        return run_knn_sorted(ctx, R, Q, self, E, knn, row_begin, row_end, idx_out, d0_out, d1_out, out_loc, key_out,
                              aux_out, idx_offset);
    SchedParams S{};
    S.mode = MODE_KNN;
    S.one = 1;
    S.self_mode = self;
    S.rb0 = (uint32_t)(row_begin / TS);
    S.n_rb = (uint32_t)((row_end + TS - 1) / TS) - S.rb0;
    S.row_begin = row_begin;
    S.row_end = row_end;
    S.knn = (uint32_t)knn;
    S.cap = cap;
    S.idx_offset = idx_offset;
    uint32_t want = (uint32_t)((8ull * ctx->sm_count + S.n_rb - 1) / S.n_rb);  // aim for >= 8 waves of CTAs
    want = std::min(want, 32u);  // the rank-by-search merge is quadratic in the number of lists per row
    S.n_split = std::max(1u, std::min(want, std::max(1u, B->dev.nb)));
    S.cb_per_split = (std::max(1u, B->dev.nb) + S.n_split - 1) / S.n_split;
    S.n_split = (std::max(1u, B->dev.nb) + S.cb_per_split - 1) / S.cb_per_split;
    const size_t part = (size_t)S.n_rb * TS * S.n_split * cap;
    void *pk, *pa = nullptr;
    skd_status st = scratch_get(ctx, "knn_key", part * 8, &pk);
    if (st) return st;
    if (with_aux) {
        st = scratch_get(ctx, "knn_aux", part * 4, &pa);
        if (st) return st;
    }
    S.part_key = (unsigned long long *)pk;
    S.part_aux = (float *)pa;
    // device-side destinations
    unsigned long long *didx = (unsigned long long *)idx_out, *dkey = (unsigned long long *)key_out;
    float *dd0 = d0_out, *dd1 = d1_out, *daux = aux_out;
    if (out_loc == SKD_HOST) {
        void *p;
        st = scratch_get(ctx, "knn_idx", n_rows * knn * 8, &p);
        if (st) return st;
        didx = dkey = (unsigned long long *)p;
        st = scratch_get(ctx, "knn_d0", n_rows * knn * 4, &p);
        if (st) return st;
        dd0 = daux = (float *)p;
        if (d1_out) {
            st = scratch_get(ctx, "knn_d1", n_rows * knn * 4, &p);
            if (st) return st;
            dd1 = (float *)p;
        }
    }
    if (partial) launch_fill_u64(dkey, n_rows * knn, ~0ull, ctx->compute);
    st = timer_begin(ctx, nullptr, 1);
    if (st) return st;
    CU(ctx, launch_tiles(A->dev, B->dev, E, S, (uint64_t)S.n_rb * S.n_split, count32, smem, ctx->compute));
    st = timer_end(ctx);
    if (st) return st;
    st = timer_begin(ctx);
    if (st) return st;
    launch_knn_merge(S.part_key, S.part_aux, S.n_split, S.knn, S.cap, row_begin - (uint64_t)S.rb0 * TS, n_rows, E.kind,
                     partial ? nullptr : didx, partial ? nullptr : dd0, partial ? nullptr : dd1, partial ? dkey : nullptr,
                     partial && with_aux ? daux : nullptr, ctx->compute);
    CU(ctx, cudaGetLastError());
    ctx->launches += 2;
    st = timer_end(ctx);
    if (st) return st;
    if (out_loc == SKD_HOST) {
        if (partial) {
            CU(ctx, cudaMemcpyAsync(key_out, dkey, n_rows * knn * 8, cudaMemcpyDeviceToHost, ctx->compute));
            if (aux_out && with_aux)
                CU(ctx, cudaMemcpyAsync(aux_out, daux, n_rows * knn * 4, cudaMemcpyDeviceToHost, ctx->compute));
        } else {
            CU(ctx, cudaMemcpyAsync(idx_out, didx, n_rows * knn * 8, cudaMemcpyDeviceToHost, ctx->compute));
            CU(ctx, cudaMemcpyAsync(d0_out, dd0, n_rows * knn * 4, cudaMemcpyDeviceToHost, ctx->compute));
            if (d1_out) CU(ctx, cudaMemcpyAsync(d1_out, dd1, n_rows * knn * 4, cudaMemcpyDeviceToHost, ctx->compute));
        }
    }
    CU(ctx, cudaStreamSynchronize(ctx->compute));
    return timers_collect(ctx);
}

// Final sparse rows from n_parts device-resident partial lists (keys[part][n_rows][knn]).
skd_status finalize_lists(skd_ctx *ctx, const unsigned long long *dkeys, const float *daux, uint32_t n_parts,
                          uint64_t part_stride, uint64_t n_rows, uint64_t knn, int kind, uint64_t *idx_out, float *d0_out,
                          float *d1_out, skd_loc out_loc) {
    unsigned long long *didx = (unsigned long long *)idx_out;
    float *dd0 = d0_out, *dd1 = d1_out;
    skd_status st;
    if (out_loc == SKD_HOST) {
        void *p;
        st = scratch_get(ctx, "knn_idx", n_rows * knn * 8, &p);
        if (st) return st;
        didx = (unsigned long long *)p;
        st = scratch_get(ctx, "knn_d0", n_rows * knn * 4, &p);
        if (st) return st;
        dd0 = (float *)p;
        if (d1_out) {
            st = scratch_get(ctx, "knn_d1", n_rows * knn * 4, &p);
            if (st) return st;
            dd1 = (float *)p;
        }
    }
    st = timer_begin(ctx);
    if (st) return st;
    launch_knn_finalize(dkeys, daux, n_parts, part_stride, 0, n_rows, (uint32_t)knn, kind, didx, dd0, dd1, nullptr, nullptr,
                        ctx->compute);
    CU(ctx, cudaGetLastError());
    ctx->launches++;
    st = timer_end(ctx);
    if (st) return st;
    if (out_loc == SKD_HOST) {
        CU(ctx, cudaMemcpyAsync(idx_out, didx, n_rows * knn * 8, cudaMemcpyDeviceToHost, ctx->compute));
        CU(ctx, cudaMemcpyAsync(d0_out, dd0, n_rows * knn * 4, cudaMemcpyDeviceToHost, ctx->compute));
        if (d1_out) CU(ctx, cudaMemcpyAsync(d1_out, dd1, n_rows * knn * 4, cudaMemcpyDeviceToHost, ctx->compute));
    }
    CU(ctx, cudaStreamSynchronize(ctx->compute));
    return SKD_OK;
}

// Symmetric self-kNN: every unordered pair is evaluated once.  Row groups of the upper triangle are computed
// by the tile kernel into a row-major key matrix in HBM (rect mode) and folded into per-row lists twice by
// select_sym_kernel: along rows (candidates j for row i) and along columns (candidates i for row j).
// Lists: dkey[n][knn] (+ daux), device memory, covering ALL n rows; only triangle rows [row_begin, row_end)
// are evaluated, so ranks that split the triangle produce partial lists to be merged (skd_knn_merge).

skd_status run_knn_symmetric(skd_ctx *ctx, const skd_db *db, const EpiParams &E, uint64_t knn, uint64_t row_begin,
                             uint64_t row_end, unsigned long long *dkey, float *daux) {
    const uint64_t n = db->n;
    const bool with_aux = E.kind == KIND_CORE_ACC;
    const uint64_t elt = with_aux ? 8 : 4;
    launch_fill_u64(dkey, n * knn, ~0ull, ctx->compute);
    ctx->launches++;
    row_begin = row_begin / TS * TS;  // groups are block aligned; callers pass block-aligned shards
    for (uint64_t g0 = row_begin; g0 < row_end && g0 + 1 < n;) {
        const uint64_t C = n - g0;                     // columns g0 .. n-1
        const uint64_t ld = (C + 3) / 4 * 4;
        uint64_t R = std::max<uint64_t>(TS, ctx->sym_chunk_bytes / (ld * elt) / TS * TS);
        R = std::min<uint64_t>(R, (row_end - g0 + TS - 1) / TS * TS);
        const uint64_t g1 = std::min<uint64_t>({g0 + R, row_end, n});
        void *D;
        skd_status st = scratch_get(ctx, "knn_rect", R * ld * elt + 64, &D);  // + 64: 8-column vector reads of the last row
        if (st) return st;
        // tile kernel, rect mode, rows [g0, g1)
        SchedParams S{};
        S.mode = MODE_DENSE_SELF;
        S.one = 1;
        S.self_mode = 1;
        S.rb0 = (uint32_t)(g0 / TS);
        S.n_rb = (uint32_t)((g1 + TS - 1) / TS) - S.rb0;
        S.group = ctx->group;
        std::vector<uint64_t> off;
        build_groups(ctx->group, true, S.rb0, S.n_rb, db->dev.nb, &off, uses_dense_kernel(ctx, MODE_DENSE_SELF, db, E));
        S.n_groups = (uint32_t)off.size() - 1;
        void *doff;
        st = scratch_get(ctx, "grp_off", off.size() * sizeof(uint64_t), &doff);
        if (st) return st;
        CU(ctx, cudaMemcpyAsync(doff, off.data(), off.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->compute));
        CU(ctx, cudaStreamSynchronize(ctx->compute));  // off is a local vector (also fences reuse of grp_off / knn_rect)
        S.grp_off = (const uint64_t *)doff;
        S.row_begin = g0;
        S.row_end = g1;
        S.out = (float *)D;
        S.rect_ld = ld;
        S.rect_c0 = g0;
        const bool count32 = E.nbins >= 65536;
        const size_t smem = tile_kernel_smem(E.nk, count32, false, 0, false);
        S.dense_pairs = uses_dense_kernel(ctx, MODE_DENSE_SELF, db, E);
        if (S.dense_pairs) {
            st = dense_scratch(ctx, E, 0, &S.counts_scratch);
            if (st) return st;
        }
        st = defer_setup(ctx, db, db, E, 0, &S);
        if (st) return st;
        if (!S.dense_pairs && smem > kMaxSmem)
            return fail(ctx, SKD_ERR_UNSUPPORTED, "%d k-mer lengths at %u bins need %zu B of shared memory per tile", E.nk,
                        E.nbins, smem);
        st = timer_begin(ctx, nullptr, 1);
        if (st) return st;
        CU(ctx, launch_tiles(db->dev, db->dev, E, S, off.back(), count32, smem, ctx->compute));
        st = timer_end(ctx);
        if (st) return st;
        st = timer_begin(ctx);
        if (st) return st;
        // rows of the group: candidates to their right; columns right of g0: candidates from the group's rows
        CU(ctx, launch_select_sym(false, with_aux, (const float *)D, ld, g0, (uint32_t)(g1 - g0), n, g0, g1, dkey, daux,
                                  (uint32_t)knn, ctx->compute));
        CU(ctx, launch_select_sym(true, with_aux, (const float *)D, ld, g0, (uint32_t)(g1 - g0), n, g0, n, dkey, daux,
                                  (uint32_t)knn, ctx->compute));
        ctx->launches += 3;
        st = timer_end(ctx);
        if (st) return st;
        g0 = g1;
    }
    return SKD_OK;
}

skd_status db_build(skd_ctx *ctx, const ChunkSource &src, uint64_t n, uint64_t sample_stride, uint64_t kmer_stride,
                    uint64_t W, uint64_t sketch_size, const uint64_t *kmers, uint32_t K, skd_db **out) {
    if (!ctx || !out) return fail(ctx, SKD_ERR_INVALID, "null argument");
    if (K == 0 || !kmers) return fail(ctx, SKD_ERR_INVALID, "database has no k-mer lengths");
    if (W == 0 || W > 65535) return fail(ctx, SKD_ERR_UNSUPPORTED, "sketchsize64 %llu outside [1, 65535]", (unsigned long long)W);
    if (K > 65535) return fail(ctx, SKD_ERR_UNSUPPORTED, "too many k-mer lengths");
    if (sketch_size == 0) sketch_size = 64 * W;
    if (kmer_stride < W * BBITS || sample_stride < (uint64_t)(K - 1) * kmer_stride + W * BBITS)
        return fail(ctx, SKD_ERR_INVALID, "strides too small for sketchsize64=%llu, %u k-mers", (unsigned long long)W, K);
    if (n >= (1ull << 32) - TS) return fail(ctx, SKD_ERR_UNSUPPORTED, "more than 2^32 samples");
    CU(ctx, cudaSetDevice(ctx->device));
    skd_db *db = new (std::nothrow) skd_db;
    if (!db) return fail(ctx, SKD_ERR_NOMEM, "out of host memory");
    db->ctx = ctx;
    db->kmers.assign(kmers, kmers + K);
    db->sketch_size = sketch_size;
    db->W = W;
    db->n = n;
    db->dev.n = n;
    db->dev.nb = (uint32_t)((n + TS - 1) / TS);
    db->dev.W = (uint32_t)W;
    db->dev.K = K;
    db->tiled_bytes = (size_t)K * W * db->dev.nb * STAGE_BYTES;
    auto bail = [&](skd_status st) {
        skd_db_free(db);
        return st;
    };
#define CUB(expr)                                                                                       \
    do {                                                                                                \
        cudaError_t e__ = (expr);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return bail(fail(ctx, SKD_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, \
                             __LINE__));                                                                \
    } while (0)
    if (db->tiled_bytes) CUB(cudaMallocAsync((void **)&db->tiled, db->tiled_bytes, ctx->compute));
    db->dev.tiled = db->tiled;
    {   // ln J by match count with the host libm (the reference calls f64::ln per pair, jaccard.rs:88)
        const uint32_t nbins = (uint32_t)(64 * W);
        std::vector<double> lny((size_t)nbins + 1);
        for (uint32_t c = 0; c <= nbins; ++c) lny[c] = std::log(host_jaccard(c, nbins));
        CUB(cudaMallocAsync((void **)&db->lny_dev, lny.size() * sizeof(double), ctx->compute));
        CUB(cudaMemcpyAsync(db->lny_dev, lny.data(), lny.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->compute));
        CUB(cudaStreamSynchronize(ctx->compute));  // lny is a local vector
    }
    if (src.alloc_only) {  // peer.cu fills the planes from the peers' slices
        *out = db;
        return SKD_OK;
    }
    bool async_ok = false;
    if (n > 0 && src.async && src.direct && src.loc == SKD_HOST) {
        cudaPointerAttributes attr;
        async_ok = cudaPointerGetAttributes(&attr, src.direct) == cudaSuccess && attr.type == cudaMemoryTypeHost;
        if (!async_ok) (void)cudaGetLastError();
    }
    if (async_ok) {
        // queue the chunked H2D copies on the copy stream and return; the chunks are re-tiled by the first
        // compute call (materialize / the interleaved dense self path)
        auto &p = db->pending;
        p.sample_stride = sample_stride;
        p.kmer_stride = kmer_stride;
        CUB(cudaMallocAsync((void **)&p.raw, (size_t)n * sample_stride * 8, ctx->compute));
        CUB(cudaEventRecord(ctx->ev_a[0], ctx->compute));
        CUB(cudaStreamWaitEvent(ctx->copy, ctx->ev_a[0], 0));  // the allocation is ordered before the copies
        // Geometric chunks (1/8, 1/8, 1/4, 1/2 of the sample blocks): the dense-self tiles enabled by a chunk grow
        // with the square of the columns received while the transfer grows linearly, so early chunks must be small
        // for the kernel to start early and late ones large for few, well-filled launches.
        std::vector<uint32_t> ends;
        for (uint32_t div : {8u, 4u, 2u, 1u}) {
            uint32_t e = std::max(1u, db->dev.nb / div);
            if (ends.empty() || e > ends.back()) ends.push_back(e);
        }
        ends.back() = db->dev.nb;
        uint32_t b0 = 0;
        for (size_t ci = 0; ci < ends.size(); b0 = ends[ci], ++ci) {
            const uint32_t b1 = ends[ci];
            const uint64_t s0 = (uint64_t)b0 * TS, s1 = std::min<uint64_t>(n, (uint64_t)b1 * TS);
            CUB(cudaMemcpyAsync(p.raw + s0 * sample_stride, src.direct + s0 * sample_stride,
                                (size_t)(s1 - s0) * sample_stride * 8, cudaMemcpyHostToDevice, ctx->copy));
            cudaEvent_t ev;
            CUB(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            CUB(cudaEventRecord(ev, ctx->copy));
            p.ready.push_back(ev);
            p.block_end.push_back(b1);
        }
    } else if (n > 0) {
        if (src.direct && src.loc == SKD_DEVICE) {
            // a device array is usually produced on another stream (a framework's): wait for the whole device
            CUB(cudaDeviceSynchronize());
            launch_repack(src.direct, n, sample_stride, kmer_stride, db->tiled, db->dev, 0, db->dev.nb, ctx->compute);
            CUB(cudaGetLastError());
            ctx->launches++;
            CUB(cudaStreamSynchronize(ctx->compute));
        } else {
            // chunks of whole sample blocks: host -> (pinned) -> device raw -> repack, double buffered
            const size_t block_bytes = (size_t)TS * sample_stride * 8;
            const uint32_t blocks_per_chunk = (uint32_t)std::max<size_t>(1, ctx->upload_chunk_bytes / block_bytes);
            const size_t chunk_bytes = (size_t)blocks_per_chunk * block_bytes;
            void *raw[2];
            skd_status st = scratch_get(ctx, "raw0", chunk_bytes, &raw[0]);
            if (st) return bail(st);
            st = scratch_get(ctx, "raw1", chunk_bytes, &raw[1]);
            if (st) return bail(st);
            // the staging buffers come from the stream-ordered allocator of the compute stream: order the copy
            // stream's first use after the allocation
            CUB(cudaEventRecord(ctx->ev_a[0], ctx->compute));
            CUB(cudaStreamWaitEvent(ctx->copy, ctx->ev_a[0], 0));
            const bool staged = src.fd >= 0;
            if (staged) {
                st = pinned_get(ctx, chunk_bytes);
                if (st) return bail(st);
            }
            int c = 0;
            for (uint32_t b0 = 0; b0 < db->dev.nb; b0 += blocks_per_chunk, ++c) {
                const int bi = c & 1;
                const uint32_t nblk = std::min(blocks_per_chunk, db->dev.nb - b0);
                const uint64_t s0 = (uint64_t)b0 * TS, s1 = std::min<uint64_t>(n, s0 + (uint64_t)nblk * TS);
                const size_t bytes = (size_t)(s1 - s0) * sample_stride * 8;
                if (c >= 2) CUB(cudaEventSynchronize(ctx->ev_b[bi]));  // repack of chunk c-2 finished with raw[bi]
                const void *hsrc;
                if (staged) {
                    if (c >= 2) CUB(cudaEventSynchronize(ctx->ev_a[bi]));  // H2D of chunk c-2 finished with pinned[bi]
                    char *dst = (char *)ctx->pinned[bi];
                    const size_t sbytes = sample_stride * 8;
                    // several host threads copy from the page cache / disk at once: one pread stream tops out
                    // near 8 GB/s, far below the PCIe link the staging buffer feeds (B200 box, page cache:
                    // 8 threads 33 GB/s, 16 threads 41.5 GB/s)
                    const int n_thr = (int)std::max<size_t>(1, std::min<size_t>(ctx->read_threads, bytes >> 22));
                    std::vector<int> bad(n_thr, 0);
                    auto work = [&](int t) {
                        const uint64_t cnt = s1 - s0, a = s0 + cnt * t / n_thr, b = s0 + cnt * (t + 1) / n_thr;
                        for (uint64_t s = a; s < b;) {
                            // contiguous file range when there is no subset list, else runs of consecutive positions
                            uint64_t run = b - s;
                            if (src.indices) {
                                run = 1;
                                while (s + run < b && src.indices[s + run] == src.indices[s] + run) ++run;
                            }
                            const uint64_t fpos = (src.indices ? src.indices[s] : s) * sbytes;
                            size_t want = (size_t)run * sbytes, got = 0;
                            while (got < want) {
                                ssize_t r = pread(src.fd, dst + (s - s0) * sbytes + got, want - got, (off_t)(fpos + got));
                                if (r <= 0) {
                                    bad[t] = 1;
                                    return;
                                }
                                got += (size_t)r;
                            }
                            s += run;
                        }
                    };
                    if (n_thr == 1) {
                        work(0);
                    } else {
                        std::vector<std::thread> pool;
                        for (int t = 0; t < n_thr; ++t) pool.emplace_back(work, t);
                        for (auto &th : pool) th.join();
                    }
                    for (int t = 0; t < n_thr; ++t)
                        if (bad[t]) return bail(fail(ctx, SKD_ERR_IO, "short read from .skd file"));
                    hsrc = dst;
                } else {
                    hsrc = src.direct + s0 * sample_stride;
                }
                CUB(cudaMemcpyAsync(raw[bi], hsrc, bytes, cudaMemcpyHostToDevice, ctx->copy));
                CUB(cudaEventRecord(ctx->ev_a[bi], ctx->copy));
                CUB(cudaStreamWaitEvent(ctx->compute, ctx->ev_a[bi], 0));
                launch_repack((const uint64_t *)raw[bi], s1 - s0, sample_stride, kmer_stride, db->tiled, db->dev, b0, nblk,
                              ctx->compute);
                CUB(cudaGetLastError());
                ctx->launches++;
                CUB(cudaEventRecord(ctx->ev_b[bi], ctx->compute));
            }
            CUB(cudaStreamSynchronize(ctx->copy));
            CUB(cudaStreamSynchronize(ctx->compute));
        }
    }
#undef CUB
    *out = db;
    return SKD_OK;
}

}  // namespace skdi
using namespace skdi;

// =================================================================================================
extern "C" {

int skd_device_count(void) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return count;
}

// SKD_BIND_CPU=1: bind the calling thread to the CPUs next to the GPU (NVML's ideal affinity) before the context
// allocates anything, so that the pinned host buffers it hands out (first touch) and the thread that fills them sit on
// the GPU's NUMA node.  One process (or thread) per GPU is the intended arrangement; opt-in because it changes the
// caller's affinity.  Best effort: without NVML nothing happens.
static void bind_cpu_to_device(int device) {
    char bus[32];
    if (cudaDeviceGetPCIBusId(bus, (int)sizeof bus, device) != cudaSuccess) return;
    void *lib = dlopen("libnvidia-ml.so.1", RTLD_NOW);
    if (!lib) return;
    auto init = (int (*)())dlsym(lib, "nvmlInit_v2");
    auto by_pci = (int (*)(const char *, void **))dlsym(lib, "nvmlDeviceGetHandleByPciBusId_v2");
    auto set_aff = (int (*)(void *))dlsym(lib, "nvmlDeviceSetCpuAffinity");
    void *h = nullptr;
    if (init && by_pci && set_aff && init() == 0 && by_pci(bus, &h) == 0) (void)set_aff(h);
}

skd_status skd_ctx_create(int device, skd_ctx **out) {
    if (!out) return fail(nullptr, SKD_ERR_INVALID, "null output pointer");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, SKD_ERR_CUDA, "no CUDA device available (%s); libsketchdist has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(nullptr, SKD_ERR_INVALID, "device %d out of range [0, %d)", device, count);
    if (const char *b = getenv("SKD_BIND_CPU"))
        if (atoi(b) != 0) bind_cpu_to_device(device);
    skd_ctx *ctx = new (std::nothrow) skd_ctx;
    if (!ctx) return fail(nullptr, SKD_ERR_NOMEM, "out of host memory");
    ctx->device = device;
    cudaDeviceProp prop;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&ctx->compute, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&ctx->copy, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&ctx->aux[0], cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&ctx->aux[1], cudaStreamNonBlocking)) != cudaSuccess) {
        delete ctx;
        return fail(nullptr, SKD_ERR_CUDA, "context setup failed: %s", cudaGetErrorString(e));
    }
    if (prop.major != 10) {
        cudaStreamDestroy(ctx->compute);
        cudaStreamDestroy(ctx->copy);
        cudaStreamDestroy(ctx->aux[0]);
        cudaStreamDestroy(ctx->aux[1]);
        delete ctx;
        return fail(nullptr, SKD_ERR_UNSUPPORTED, "device %d is sm_%d%d; libsketchdist is built for sm_100a only", device,
                    prop.major, prop.minor);
    }
    ctx->sm_count = prop.multiProcessorCount;
    if ((e = preload_kernels()) != cudaSuccess || (e = preload_peer_kernels()) != cudaSuccess) {
        cudaStreamDestroy(ctx->compute);
        cudaStreamDestroy(ctx->copy);
        cudaStreamDestroy(ctx->aux[0]);
        cudaStreamDestroy(ctx->aux[1]);
        delete ctx;
        return fail(nullptr, SKD_ERR_CUDA, "loading the kernels failed: %s", cudaGetErrorString(e));
    }
    {   // keep freed database buffers in the device's default pool instead of returning them to the driver
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    if (const char *e = getenv("SKD_ZEROCOPY")) ctx->zero_copy = atoi(e) != 0;
    if (const char *e = getenv("SKD_DENSE_V1")) ctx->dense_v2 = atoi(e) == 0;
    if (const char *e = getenv("SKD_EARLY_OUT")) ctx->early_out = atoi(e) != 0;
    if (const char *e = getenv("SKD_DEFER_MAX")) ctx->defer_max = (uint32_t)std::max(0, atoi(e));
    if (const char *e = getenv("SKD_DEFER_CAP")) ctx->defer_cap = (uint32_t)std::max(1, atoi(e));
    if (const char *e = getenv("SKD_GROUP")) ctx->group = (uint32_t)std::max(1, atoi(e));
    if (const char *e = getenv("SKD_UPLOAD_CHUNK_MB")) ctx->upload_chunk_bytes = (size_t)std::max(1, atoi(e)) << 20;
    if (const char *e = getenv("SKD_KNN_SYMMETRIC")) ctx->knn_symmetric = atoi(e) != 0;
    if (const char *e = getenv("SKD_SYM_CHUNK_MB")) ctx->sym_chunk_bytes = (size_t)std::max(1, atoi(e)) << 20;
    ctx->read_threads = std::min<size_t>(16, std::max<size_t>(1, std::thread::hardware_concurrency()));
    if (const char *e = getenv("SKD_FEW_QUERIES")) ctx->few_queries = (uint64_t)std::max(0ll, atoll(e));
    if (const char *e = getenv("SKD_READ_THREADS")) ctx->read_threads = (size_t)std::max(1, atoi(e));
    if (const char *e = getenv("SKD_RESULT_CHUNK_MB")) ctx->result_chunk_bytes = (size_t)std::max(1, atoi(e)) << 20;
    for (int i = 0; i < 2; ++i) {
        cudaEventCreateWithFlags(&ctx->ev_a[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ctx->ev_b[i], cudaEventDisableTiming);
    }
    if (cudaMalloc(&ctx->work_stats, 4 * sizeof(unsigned long long)) == cudaSuccess)
        cudaMemset(ctx->work_stats, 0, 4 * sizeof(unsigned long long));
    else
        ctx->work_stats = nullptr;
    *out = ctx;
    return SKD_OK;
}

skd_status skd_ctx_set_early_exit(skd_ctx *ctx, int enabled) {
    if (!ctx) return SKD_ERR_INVALID;
    ctx->early_out = enabled != 0;
    return SKD_OK;
}

skd_status skd_ctx_work_counters(skd_ctx *ctx, uint64_t *steps_computed, uint64_t *steps_total, uint64_t *pairs_deferred,
                                 int reset) {
    if (!ctx) return SKD_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    unsigned long long v[4] = {0, 0, 0, 0};
    if (ctx->work_stats) {
        CU(ctx, cudaStreamSynchronize(ctx->compute));
        CU(ctx, cudaStreamSynchronize(ctx->aux[1]));
        CU(ctx, cudaMemcpy(v, ctx->work_stats, sizeof v, cudaMemcpyDeviceToHost));
        if (reset) CU(ctx, cudaMemset(ctx->work_stats, 0, sizeof v));
    }
    if (steps_computed) *steps_computed = v[1] - v[3];
    if (steps_total) *steps_total = v[1];
    if (pairs_deferred) *pairs_deferred = v[2];
    return SKD_OK;
}

void skd_ctx_destroy(skd_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (auto &kv : ctx->scratch)
        if (kv.second.first) cudaFreeAsync(kv.second.first, ctx->compute);
    if (ctx->compute) cudaStreamSynchronize(ctx->compute);
    for (int i = 0; i < 2; ++i) {
        if (ctx->pinned[i]) cudaFreeHost(ctx->pinned[i]);
        if (ctx->ev_a[i]) cudaEventDestroy(ctx->ev_a[i]);
        if (ctx->ev_b[i]) cudaEventDestroy(ctx->ev_b[i]);
    }
    for (auto &t : ctx->timers) {
        cudaEventDestroy(t.first);
        cudaEventDestroy(t.second);
    }
    if (ctx->work_stats) cudaFree(ctx->work_stats);
    if (ctx->compute) cudaStreamDestroy(ctx->compute);
    if (ctx->copy) cudaStreamDestroy(ctx->copy);
    for (int i = 0; i < 2; ++i)
        if (ctx->aux[i]) cudaStreamDestroy(ctx->aux[i]);
    delete ctx;
}

const char *skd_last_error(const skd_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
int skd_ctx_device(const skd_ctx *ctx) { return ctx ? ctx->device : -1; }
uint64_t skd_ctx_launch_count(const skd_ctx *ctx) { return ctx ? ctx->launches : 0; }
float skd_ctx_last_kernel_ms(const skd_ctx *ctx) { return ctx ? ctx->last_ms : 0.f; }
float skd_ctx_last_tile_kernel_ms(const skd_ctx *ctx) { return ctx ? ctx->last_tile_ms : 0.f; }
void *skd_ctx_stream(const skd_ctx *ctx) { return ctx ? (void *)ctx->compute : nullptr; }

skd_status skd_ctx_synchronize(skd_ctx *ctx) {
    if (!ctx) return SKD_ERR_INVALID;
    CU(ctx, cudaStreamSynchronize(ctx->copy));
    CU(ctx, cudaStreamSynchronize(ctx->compute));
    return SKD_OK;
}

skd_status skd_host_alloc(skd_ctx *ctx, size_t bytes, void **out) {
    if (!ctx || !out) return SKD_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaMallocHost(out, std::max<size_t>(bytes, 1)));
    return SKD_OK;
}
void skd_host_free(skd_ctx *ctx, void *ptr) {
    (void)ctx;
    if (ptr) cudaFreeHost(ptr);
}
skd_status skd_device_alloc(skd_ctx *ctx, size_t bytes, void **out) {
    if (!ctx || !out) return SKD_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaMalloc(out, std::max<size_t>(bytes, 1)));
    return SKD_OK;
}
void skd_device_free(skd_ctx *ctx, void *ptr) {
    (void)ctx;
    if (ptr) cudaFree(ptr);
}
skd_status skd_memcpy(skd_ctx *ctx, void *dst, skd_loc dst_loc, const void *src, skd_loc src_loc, size_t bytes) {
    if (!ctx) return SKD_ERR_INVALID;
    (void)dst_loc;
    (void)src_loc;
    CU(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, ctx->compute));
    CU(ctx, cudaStreamSynchronize(ctx->compute));
    return SKD_OK;
}

skd_status skd_db_upload(skd_ctx *ctx, const uint64_t *bins, skd_loc bins_loc, uint64_t n_samples,
                         uint64_t sample_stride, uint64_t kmer_stride, uint64_t sketchsize64, uint64_t sketch_size,
                         const uint64_t *kmer_lengths, uint32_t n_kmers, skd_db **out) {
    if (!ctx) return fail(nullptr, SKD_ERR_INVALID, "null context");
    if (n_samples > 0 && !bins) return fail(ctx, SKD_ERR_INVALID, "null bins");
    ChunkSource src;
    src.direct = bins;
    src.loc = bins_loc;
    return db_build(ctx, src, n_samples, sample_stride, kmer_stride, sketchsize64, sketch_size, kmer_lengths, n_kmers,
                    out);
}

skd_status skd_db_upload_async(skd_ctx *ctx, const uint64_t *bins, uint64_t n_samples, uint64_t sample_stride,
                               uint64_t kmer_stride, uint64_t sketchsize64, uint64_t sketch_size,
                               const uint64_t *kmer_lengths, uint32_t n_kmers, skd_db **out) {
    if (!ctx) return fail(nullptr, SKD_ERR_INVALID, "null context");
    if (n_samples > 0 && !bins) return fail(ctx, SKD_ERR_INVALID, "null bins");
    ChunkSource src;
    src.direct = bins;
    src.loc = SKD_HOST;
    src.async = true;
    return db_build(ctx, src, n_samples, sample_stride, kmer_stride, sketchsize64, sketch_size, kmer_lengths, n_kmers,
                    out);
}

skd_status skd_db_load_skd(skd_ctx *ctx, const char *skd_path, const uint64_t *sample_indices, uint64_t n_samples,
                           uint64_t sample_stride, uint64_t kmer_stride, uint64_t sketchsize64, uint64_t sketch_size,
                           const uint64_t *kmer_lengths, uint32_t n_kmers, skd_db **out) {
    if (!ctx) return fail(nullptr, SKD_ERR_INVALID, "null context");
    if (!skd_path) return fail(ctx, SKD_ERR_INVALID, "null path");
    int fd = open(skd_path, O_RDONLY);
    if (fd < 0) return fail(ctx, SKD_ERR_IO, "Could not read from %s", skd_path);  // sketch_datafile.rs:123-125
    struct stat sb;
    if (fstat(fd, &sb) != 0) {
        close(fd);
        return fail(ctx, SKD_ERR_IO, "Could not stat %s", skd_path);
    }
    const uint64_t file_samples = sample_stride ? (uint64_t)sb.st_size / (sample_stride * 8) : 0;
    for (uint64_t s = 0; s < n_samples; ++s) {
        uint64_t fi = sample_indices ? sample_indices[s] : s;
        if (fi >= file_samples) {
            close(fd);
            return fail(ctx, SKD_ERR_IO, "%s holds %llu samples; sample index %llu requested", skd_path,
                        (unsigned long long)file_samples, (unsigned long long)fi);
        }
    }
    ChunkSource src;
    src.fd = fd;
    src.indices = sample_indices;
    skd_status st = db_build(ctx, src, n_samples, sample_stride, kmer_stride, sketchsize64, sketch_size, kmer_lengths,
                             n_kmers, out);
    close(fd);
    return st;
}

void skd_db_free(skd_db *db) {
    if (!db) return;
    if (db->ctx) cudaSetDevice(db->ctx->device);
    cudaStream_t st = db->ctx ? db->ctx->compute : nullptr;
    if (db->is_pending()) {  // never used: wait for the copies still reading the caller's buffer
        if (db->ctx) cudaStreamSynchronize(db->ctx->copy);
        for (auto e : db->pending.ready) cudaEventDestroy(e);
        cudaFreeAsync(db->pending.raw, st);
    }
    if (db->tiled) cudaFreeAsync(db->tiled, st);  // back to the device pool: the next upload reuses it
    if (db->lny_dev) cudaFreeAsync(db->lny_dev, st);
    if (db->rows) cudaFreeAsync(db->rows, st);
    delete db;
}
skd_status skd_db_adopt(skd_ctx *ctx, skd_db *db) {
    if (!ctx || !db || !db->ctx) return fail(ctx, SKD_ERR_INVALID, "null context or database");
    if (db->ctx == ctx) return SKD_OK;
    if (db->ctx->device != ctx->device)
        return fail(ctx, SKD_ERR_INVALID, "database lives on device %d, context on device %d", db->ctx->device, ctx->device);
    CU(ctx, cudaSetDevice(ctx->device));
    skd_status st = materialize(db->ctx, db);  // an upload still in flight is completed by its own context
    if (st) return fail(ctx, st, "%s", db->ctx->err.c_str());
    CU(ctx, cudaStreamSynchronize(db->ctx->compute));
    db->ctx = ctx;
    return SKD_OK;
}
uint64_t skd_db_n_samples(const skd_db *db) { return db ? db->n : 0; }
uint32_t skd_db_n_kmers(const skd_db *db) { return db ? (uint32_t)db->kmers.size() : 0; }
uint64_t skd_db_sketch_size(const skd_db *db) { return db ? db->sketch_size : 0; }
uint64_t skd_db_hbm_bytes(const skd_db *db) { return db ? db->tiled_bytes : 0; }
int skd_db_compatible(const skd_db *a, const skd_db *b) {
    return a && b && a->kmers == b->kmers && a->sketch_size == b->sketch_size && a->W == b->W;
}

skd_status skd_set_k(skd_ctx *ctx, const skd_db *db, int64_t kmer, int ani, skd_dist_type *out) {
    if (!ctx || !db || !out) return fail(ctx, SKD_ERR_INVALID, "null argument");
    if (kmer < 0) {
        out->kind = SKD_CORE_ACC;
        out->k_idx = 0;
        return SKD_OK;
    }
    for (size_t i = 0; i < db->kmers.size(); ++i)
        if (db->kmers[i] == (uint64_t)kmer) {  // get_k_idx: first match, multisketch.rs:115-121
            out->kind = ani ? SKD_ANI : SKD_JACCARD;
            out->k_idx = (int32_t)i;
            return SKD_OK;
        }
    return fail(ctx, SKD_ERR_K_NOT_FOUND, "K-mer size %lld not found in file", (long long)kmer);
}

uint64_t skd_self_pairs_in_rows(uint64_t n, uint64_t row_begin, uint64_t row_end) {
    if (n < 2 || row_begin >= row_end) return 0;
    row_end = std::min(row_end, n);
    if (row_begin >= row_end) return 0;
    // sum_{i=row_begin}^{row_end-1} (n - 1 - i)
    uint64_t cnt = row_end - row_begin;
    return cnt * (n - 1) - (row_begin + row_end - 1) * cnt / 2;
}

skd_status skd_self_dists_all(skd_ctx *ctx, const skd_db *db, skd_dist_type dist, const double *completeness,
                              double cutoff, uint64_t row_begin, uint64_t row_end, float *out, skd_loc out_loc) {
    skd_status st = check_pair(ctx, db, db, /*keep_pending=*/true);  // run_dense overlaps an upload in flight
    if (st) return st;
    CU(ctx, cudaSetDevice(ctx->device));
    EpiParams E;
    st = build_epi(ctx, db, db, dist, completeness, completeness, cutoff, &E);
    if (st) return st;
    row_end = std::min(row_end, db->n);
    if (db->n < 2 || row_begin >= row_end) {
        timers_reset(ctx);
        ctx->last_ms = 0.f;
        return SKD_OK;
    }
    if (!out) return fail(ctx, SKD_ERR_INVALID, "null output");
    return run_dense(ctx, db, db, true, E, MODE_DENSE_SELF, row_begin, row_end, out, out_loc);
}

skd_status skd_cross_dists_all(skd_ctx *ctx, const skd_db *ref, const skd_db *query, skd_dist_type dist,
                               const double *ref_c, const double *query_c, double cutoff, uint64_t ref_begin,
                               uint64_t ref_end, float *out, skd_loc out_loc) {
    skd_status st = check_pair(ctx, ref, query);
    if (st) return st;
    CU(ctx, cudaSetDevice(ctx->device));
    EpiParams E;
    st = build_epi(ctx, ref, query, dist, ref_c, query_c, cutoff, &E);
    if (st) return st;
    ref_end = std::min(ref_end, ref->n);
    if (ref_begin >= ref_end || query->n == 0) {
        timers_reset(ctx);
        ctx->last_ms = 0.f;
        return SKD_OK;
    }
    if (!out) return fail(ctx, SKD_ERR_INVALID, "null output");
    return run_dense(ctx, ref, query, false, E, MODE_DENSE_CROSS, ref_begin, ref_end, out, out_loc);
}

skd_status skd_self_dists_knn(skd_ctx *ctx, const skd_db *db, uint64_t knn, skd_dist_type dist,
                              const double *completeness, double cutoff, uint64_t row_begin, uint64_t row_end,
                              uint64_t *idx_out, float *d0_out, float *d1_out, skd_loc out_loc) {
    skd_status st = check_pair(ctx, db, db);
    if (st) return st;
    CU(ctx, cudaSetDevice(ctx->device));
    if (db->n < 2 || knn == 0 || knn > db->n - 1)
        return fail(ctx, SKD_ERR_INVALID, "knn must be in [1, n-1] (n = %llu, knn = %llu)", (unsigned long long)db->n,
                    (unsigned long long)knn);
    EpiParams E;
    st = build_epi(ctx, db, db, dist, completeness, completeness, cutoff, &E);
    if (st) return st;
    row_end = std::min(row_end, db->n);
    if (row_begin > row_end) row_begin = row_end;
    if (!idx_out || !d0_out) return fail(ctx, SKD_ERR_INVALID, "null output");
    if (ctx->knn_symmetric && row_begin == 0 && row_end == db->n && knn <= kSymKnnMax) {
        // whole matrix on one GPU: evaluate each unordered pair once (the reference evaluates both orders)
        const uint64_t n = db->n;
        void *lk, *la = nullptr;
        st = scratch_get(ctx, "sym_key", n * knn * 8, &lk);
        if (st) return st;
        if (E.kind == KIND_CORE_ACC) {
            st = scratch_get(ctx, "sym_aux", n * knn * 4, &la);
            if (st) return st;
        }
        timers_reset(ctx);
        st = run_knn_symmetric(ctx, db, E, knn, 0, n, (unsigned long long *)lk, (float *)la);
        if (st) return st;
        st = finalize_lists(ctx, (const unsigned long long *)lk, (const float *)la, 1, 0, n, knn, E.kind, idx_out, d0_out,
                            d1_out, out_loc);
        if (st) return st;
        return timers_collect(ctx);
    }
    return run_knn(ctx, db, db, true, E, knn, row_begin, row_end, idx_out, d0_out, d1_out, out_loc);
}

skd_status skd_cross_dists_knn(skd_ctx *ctx, const skd_db *ref, const skd_db *query, uint64_t knn, skd_dist_type dist,
                               const double *ref_c, const double *query_c, double cutoff, uint64_t query_begin,
                               uint64_t query_end, uint64_t *idx_out, float *d0_out, float *d1_out, skd_loc out_loc,
                               uint64_t *knn_used) {
    skd_status st = check_pair(ctx, ref, query);
    if (st) return st;
    CU(ctx, cudaSetDevice(ctx->device));
    if (ref->n == 0) return fail(ctx, SKD_ERR_EMPTY_DB, "Reference database has no loaded samples");
    if (query->n == 0) return fail(ctx, SKD_ERR_EMPTY_DB, "Query database has no loaded samples");
    if (knn == 0) return fail(ctx, SKD_ERR_INVALID, "knn must be at least 1");
    knn = std::min<uint64_t>(knn, ref->n);  // mod.rs:325
    if (knn_used) *knn_used = knn;
    EpiParams E;
    // tile rows are queries (A side), columns refs (B side): completeness sides swap accordingly
    st = build_epi(ctx, query, ref, dist, query_c, ref_c, cutoff, &E);
    if (st) return st;
    query_end = std::min(query_end, query->n);
    if (query_begin > query_end) query_begin = query_end;
    if (!idx_out || !d0_out) return fail(ctx, SKD_ERR_INVALID, "null output");
    return run_knn(ctx, ref, query, false, E, knn, query_begin, query_end, idx_out, d0_out, d1_out, out_loc);
}

skd_status skd_self_knn_partial(skd_ctx *ctx, const skd_db *db, uint64_t knn, skd_dist_type dist,
                                const double *completeness, double cutoff, uint64_t row_begin, uint64_t row_end,
                                uint64_t *keys_out, float *aux_out, skd_loc out_loc) {
    skd_status st = check_pair(ctx, db, db);
    if (st) return st;
    CU(ctx, cudaSetDevice(ctx->device));
    if (db->n < 2 || knn == 0 || knn > db->n - 1)
        return fail(ctx, SKD_ERR_INVALID, "knn must be in [1, n-1] (n = %llu, knn = %llu)", (unsigned long long)db->n,
                    (unsigned long long)knn);
    if (row_begin % TS != 0) return fail(ctx, SKD_ERR_INVALID, "row_begin must be a multiple of %d", TS);
    EpiParams E;
    st = build_epi(ctx, db, db, dist, completeness, completeness, cutoff, &E);
    if (st) return st;
    row_end = std::min(row_end, db->n);
    if (!keys_out) return fail(ctx, SKD_ERR_INVALID, "null output");
    const uint64_t n = db->n;
    const bool with_aux = E.kind == KIND_CORE_ACC;
    unsigned long long *dk = (unsigned long long *)keys_out;
    float *da = aux_out;
    if (out_loc == SKD_HOST) {
        void *p;
        st = scratch_get(ctx, "sym_key", n * knn * 8, &p);
        if (st) return st;
        dk = (unsigned long long *)p;
        st = scratch_get(ctx, "sym_aux", n * knn * 4, &p);
        if (st) return st;
        da = (float *)p;
    } else if (with_aux && !aux_out) {
        return fail(ctx, SKD_ERR_INVALID, "aux_out is required for core/accessory");
    }
    timers_reset(ctx);
    if (knn <= kSymKnnMax) {
        st = run_knn_symmetric(ctx, db, E, knn, row_begin, std::max(row_begin, row_end), dk, da);
    } else {
        // The symmetric fold keeps knn + 64 candidates per row and step; beyond that the rows [row_begin, row_end) are
        // evaluated against ALL columns (both orders of a pair, like the reference) and only they get lists: the row
        // blocks of a triangle partition cover every row once, so the merged result is the same.
        launch_fill_u64(dk, n * knn, ~0ull, ctx->compute);
        ctx->launches++;
        row_end = std::max(row_begin, row_end);
        st = run_knn(ctx, db, db, true, E, knn, row_begin, row_end, nullptr, nullptr, nullptr, SKD_DEVICE,
                     (uint64_t *)(dk + row_begin * knn), with_aux ? da + row_begin * knn : nullptr, 0);
    }
    if (st) return st;
    if (out_loc == SKD_HOST) {
        CU(ctx, cudaMemcpyAsync(keys_out, dk, n * knn * 8, cudaMemcpyDeviceToHost, ctx->compute));
        if (aux_out && with_aux) CU(ctx, cudaMemcpyAsync(aux_out, da, n * knn * 4, cudaMemcpyDeviceToHost, ctx->compute));
    }
    CU(ctx, cudaStreamSynchronize(ctx->compute));
    return timers_collect(ctx);
}

skd_status skd_cross_knn_partial(skd_ctx *ctx, const skd_db *ref_shard, const skd_db *query, uint64_t knn,
                                 skd_dist_type dist, const double *ref_c, const double *query_c, double cutoff,
                                 uint64_t ref_index_offset, uint64_t *keys_out, float *aux_out, skd_loc out_loc) {
    skd_status st = check_pair(ctx, ref_shard, query);
    if (st) return st;
    CU(ctx, cudaSetDevice(ctx->device));
    if (query->n == 0) return fail(ctx, SKD_ERR_EMPTY_DB, "Query database has no loaded samples");
    if (knn == 0) return fail(ctx, SKD_ERR_INVALID, "knn must be at least 1");
    if (ref_index_offset + ref_shard->n > 0xffffffffull) return fail(ctx, SKD_ERR_UNSUPPORTED, "more than 2^32 references");
    if (!keys_out) return fail(ctx, SKD_ERR_INVALID, "null output");
    EpiParams E;
    st = build_epi(ctx, query, ref_shard, dist, query_c, ref_c, cutoff, &E);
    if (st) return st;
    if (E.kind == KIND_CORE_ACC && out_loc == SKD_DEVICE && !aux_out)
        return fail(ctx, SKD_ERR_INVALID, "aux_out is required for core/accessory");
    if (ref_shard->n == 0) {  // an empty shard contributes only padding
        if (out_loc == SKD_HOST) {
            for (uint64_t i = 0; i < query->n * knn; ++i) keys_out[i] = ~0ull;
        } else {
            launch_fill_u64((unsigned long long *)keys_out, query->n * knn, ~0ull, ctx->compute);
            CU(ctx, cudaStreamSynchronize(ctx->compute));
        }
        return SKD_OK;
    }
    return run_knn(ctx, ref_shard, query, false, E, knn, 0, query->n, nullptr, nullptr, nullptr, out_loc, keys_out, aux_out,
                   ref_index_offset);
}

skd_status skd_knn_merge(skd_ctx *ctx, const uint64_t *keys, const float *aux, uint32_t n_parts, uint64_t n_rows,
                         uint64_t knn, skd_dist_type dist, skd_loc in_loc, uint64_t *idx_out, float *d0_out, float *d1_out,
                         skd_loc out_loc) {
    if (!ctx) return SKD_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    if (!keys || !idx_out || !d0_out || n_parts == 0 || knn == 0) return fail(ctx, SKD_ERR_INVALID, "bad argument");
    if (dist.kind == SKD_CORE_ACC && (!aux || !d1_out)) return fail(ctx, SKD_ERR_INVALID, "core/accessory needs aux and d1_out");
    const size_t total = (size_t)n_parts * n_rows * knn;
    const unsigned long long *dk = (const unsigned long long *)keys;
    const float *da = aux;
    timers_reset(ctx);
    if (in_loc == SKD_HOST) {
        void *p;
        skd_status st = scratch_get(ctx, "merge_key", total * 8, &p);
        if (st) return st;
        CU(ctx, cudaMemcpyAsync(p, keys, total * 8, cudaMemcpyHostToDevice, ctx->compute));
        dk = (const unsigned long long *)p;
        if (aux) {
            st = scratch_get(ctx, "merge_aux", total * 4, &p);
            if (st) return st;
            CU(ctx, cudaMemcpyAsync(p, aux, total * 4, cudaMemcpyHostToDevice, ctx->compute));
            da = (const float *)p;
        }
    }
    skd_status st = finalize_lists(ctx, dk, dist.kind == SKD_CORE_ACC ? da : nullptr, n_parts, n_rows * knn, n_rows, knn,
                                   dist.kind, idx_out, d0_out, d1_out, out_loc);
    if (st) return st;
    return timers_collect(ctx);
}

skd_status skd_self_dists_knn_candidates(skd_ctx *ctx, const skd_db *db, uint64_t knn, skd_dist_type dist,
                                         const double *completeness, double cutoff, const uint64_t *cand_offsets,
                                         const uint64_t *cand_idx, int retain_unmatched, uint64_t *idx_out,
                                         float *d0_out) {
    skd_status st = check_pair(ctx, db, db);
    if (st) return st;
    CU(ctx, cudaSetDevice(ctx->device));
    if (dist.kind == SKD_CORE_ACC)
        return fail(ctx, SKD_ERR_UNSUPPORTED, "Prefilter only available for single k-mer distances");  // mod.rs:548-550
    if (knn == 0 || knn > 2048) return fail(ctx, SKD_ERR_UNSUPPORTED, "knn must be in [1, 2048] for candidate lists");
    if (!cand_offsets || !idx_out || !d0_out) return fail(ctx, SKD_ERR_INVALID, "null argument");
    EpiParams E;
    st = build_epi(ctx, db, db, dist, completeness, completeness, cutoff, &E);
    if (st) return st;
    const uint64_t n = db->n;
    timers_reset(ctx);
    if (n == 0) return SKD_OK;
    const uint64_t total = cand_offsets[n];
    for (uint64_t i = 0; i < n; ++i)
        if (cand_offsets[i] > cand_offsets[i + 1]) return fail(ctx, SKD_ERR_INVALID, "candidate offsets must not decrease");
    for (uint64_t c = 0; c < total; ++c)
        if (cand_idx[c] >= n) return fail(ctx, SKD_ERR_INVALID, "candidate index %llu out of range", (unsigned long long)cand_idx[c]);
    uint32_t cap = 128;
    while (cap < knn + 64) cap <<= 1;
    void *doff, *didx, *oi, *od;
    st = scratch_get(ctx, "cand_off", (n + 1) * 8, &doff);
    if (st) return st;
    st = scratch_get(ctx, "cand_idx", std::max<uint64_t>(1, total) * 8, &didx);
    if (st) return st;
    st = scratch_get(ctx, "knn_idx", n * knn * 8, &oi);
    if (st) return st;
    st = scratch_get(ctx, "knn_d0", n * knn * 4, &od);
    if (st) return st;
    CU(ctx, cudaMemcpyAsync(doff, cand_offsets, (n + 1) * 8, cudaMemcpyHostToDevice, ctx->compute));
    if (total) CU(ctx, cudaMemcpyAsync(didx, cand_idx, total * 8, cudaMemcpyHostToDevice, ctx->compute));
    st = timer_begin(ctx);
    if (st) return st;
    CU(ctx, launch_knn_candidates(db->dev, E, (const uint64_t *)doff, (const uint64_t *)didx, n, (uint32_t)knn, cap,
                                  retain_unmatched, (unsigned long long *)oi, (float *)od, ctx->compute));
    ctx->launches++;
    st = timer_end(ctx);
    if (st) return st;
    CU(ctx, cudaMemcpyAsync(idx_out, oi, n * knn * 8, cudaMemcpyDeviceToHost, ctx->compute));
    CU(ctx, cudaMemcpyAsync(d0_out, od, n * knn * 4, cudaMemcpyDeviceToHost, ctx->compute));
    CU(ctx, cudaStreamSynchronize(ctx->compute));
    return timers_collect(ctx);
}

static skd_status pairs_common(skd_ctx *ctx, const skd_db *ref, const skd_db *query, const EpiParams &E,
                               const uint64_t *pairs, uint64_t n_pairs, uint32_t *counts, float *dists, int ncols) {
    timers_reset(ctx);
    if (n_pairs == 0) return SKD_OK;
    for (uint64_t p = 0; p < n_pairs; ++p)
        if (pairs[2 * p] >= ref->n || pairs[2 * p + 1] >= query->n)
            return fail(ctx, SKD_ERR_INVALID, "pair %llu out of range", (unsigned long long)p);
    void *dp, *dc = nullptr, *dd = nullptr;
    skd_status st = scratch_get(ctx, "pairs", n_pairs * 16, &dp);
    if (st) return st;
    if (counts) {
        st = scratch_get(ctx, "pair_counts", n_pairs * ref->kmers.size() * 4, &dc);
        if (st) return st;
    }
    if (dists) {
        st = scratch_get(ctx, "pair_dists", n_pairs * ncols * 4, &dd);
        if (st) return st;
    }
    CU(ctx, cudaMemcpyAsync(dp, pairs, n_pairs * 16, cudaMemcpyHostToDevice, ctx->compute));
    st = timer_begin(ctx);
    if (st) return st;
    launch_pairs(ref->dev, query->dev, E, (const uint64_t *)dp, n_pairs, (uint32_t *)dc, (float *)dd, ctx->compute);
    CU(ctx, cudaGetLastError());
    ctx->launches++;
    st = timer_end(ctx);
    if (st) return st;
    if (counts)
        CU(ctx, cudaMemcpyAsync(counts, dc, n_pairs * ref->kmers.size() * 4, cudaMemcpyDeviceToHost, ctx->compute));
    if (dists) CU(ctx, cudaMemcpyAsync(dists, dd, n_pairs * ncols * 4, cudaMemcpyDeviceToHost, ctx->compute));
    CU(ctx, cudaStreamSynchronize(ctx->compute));
    return timers_collect(ctx);
}

skd_status skd_pair_dists(skd_ctx *ctx, const skd_db *ref, const skd_db *query, skd_dist_type dist, const double *ref_c,
                          const double *query_c, double cutoff, const uint64_t *pairs, uint64_t n_pairs, float *out) {
    skd_status st = check_pair(ctx, ref, query);
    if (st) return st;
    CU(ctx, cudaSetDevice(ctx->device));
    EpiParams E;
    st = build_epi(ctx, ref, query, dist, ref_c, query_c, cutoff, &E);
    if (st) return st;
    if (n_pairs && (!pairs || !out)) return fail(ctx, SKD_ERR_INVALID, "null argument");
    return pairs_common(ctx, ref, query, E, pairs, n_pairs, nullptr, out, dist.kind == SKD_CORE_ACC ? 2 : 1);
}

skd_status skd_match_counts(skd_ctx *ctx, const skd_db *ref, const skd_db *query, const uint64_t *pairs,
                            uint64_t n_pairs, uint32_t *counts) {
    skd_status st = check_pair(ctx, ref, query);
    if (st) return st;
    CU(ctx, cudaSetDevice(ctx->device));
    EpiParams E;
    memset(&E, 0, sizeof E);
    E.kind = KIND_JACCARD;
    E.nbins = (uint32_t)(64 * ref->W);
    E.nk = 1;
    if (n_pairs && (!pairs || !counts)) return fail(ctx, SKD_ERR_INVALID, "null argument");
    return pairs_common(ctx, ref, query, E, pairs, n_pairs, counts, nullptr, 0);
}

skd_status skd_match_counts_dense(skd_ctx *ctx, const skd_db *ref, const skd_db *query, uint64_t ref_begin,
                                  uint64_t ref_end, uint32_t *counts) {
    skd_status st = check_pair(ctx, ref, query);
    if (st) return st;
    CU(ctx, cudaSetDevice(ctx->device));
    EpiParams E;
    memset(&E, 0, sizeof E);
    E.kind = KIND_CORE_ACC;  // visit every k
    E.nbins = (uint32_t)(64 * ref->W);
    E.k0 = 0;
    E.nk = (int)ref->kmers.size();
    E.kp = counts_per_pair(E.nk, E.nbins >= 65536);
    ref_end = std::min(ref_end, ref->n);
    if (ref_begin >= ref_end || query->n == 0) return SKD_OK;
    if (!counts) return fail(ctx, SKD_ERR_INVALID, "null output");
    return run_dense(ctx, ref, query, false, E, MODE_COUNTS, ref_begin, ref_end, (float *)counts, SKD_HOST);
}

}  // extern "C"
