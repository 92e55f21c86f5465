// kNN for the corners of the envelope the list-keeping kernels do not cover: a very large knn (the reference clamps
// --knn to n - 1 and runs, src/lib.rs:379-382; the row-owned tile kernel keeps at most 1984 candidates per row in
// shared memory) or so many k-mer lengths that a tile's counts do not fit shared memory (the reference loops over any
// number, src/distances/jaccard.rs:77).  Row blocks of the key matrix are computed by the dense kernel (rows = the
// kNN rows, columns = all references; counts live in L2, so any number of k-mer lengths fits), every row is sorted
// whole, and its first knn entries become the row's list.  Not a hot path: with knn close to n the output alone is a
// dense matrix, and the full sort costs about as much as the distances.
//
// The segmented sort is CUB's (cub::DeviceSegmentedSort, shipped with the CUDA toolkit); everything else is this
// library's kernels.
#include <cub/device/device_segmented_sort.cuh>

#include <algorithm>
#include <vector>

#include "internal.h"

using namespace skd;

namespace skd {

// keys[r][c] = (order bits of the row's kNN key << 32) | (c + idx_offset); the row's own column (self mode) and the
// padding columns [n_cols, ld) sort last.  Core/accessory key matrices hold (core, accessory) pairs: aux = accessory.
template <bool AUX>
__global__ void sort_keys_kernel(const float *__restrict__ D, uint64_t ld, uint64_t row0, uint64_t n_rows, uint64_t n_cols,
                                 bool self, uint64_t idx_offset, unsigned long long *__restrict__ keys,
                                 float *__restrict__ aux) {
    const uint64_t total = n_rows * ld;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = e / ld, c = e % ld;
        unsigned long long k = ~0ull;
        float a = 0.0f;
        if (c < n_cols && !(self && row0 + r == c)) {
            float d0;
            if (AUX) {
                const float2 v = reinterpret_cast<const float2 *>(D)[e];
                d0 = v.x;
                a = v.y;
            } else {
                d0 = D[e];
            }
            k = ((unsigned long long)float_order_bits(d0) << 32) | (uint32_t)(c + idx_offset);
        }
        keys[e] = k;
        if (AUX) aux[e] = a;
    }
}

// First knn entries of every sorted row -> list [row][knn].
__global__ void take_lists_kernel(const unsigned long long *__restrict__ keys, const float *__restrict__ aux, uint64_t ld,
                                  uint64_t n_rows, uint32_t knn, unsigned long long *__restrict__ key_out,
                                  float *__restrict__ aux_out) {
    const uint64_t total = n_rows * knn;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = e / knn, t = e % knn;
        key_out[e] = t < ld ? keys[r * ld + t] : ~0ull;  // a list longer than the row (knn > references of a shard)
        if (aux_out) aux_out[e] = t < ld ? aux[r * ld + t] : 0.0f;
    }
}

__global__ void segment_offsets_kernel(uint64_t n_rows, uint64_t ld, long long *__restrict__ off) {
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= n_rows; r += (uint64_t)gridDim.x * blockDim.x)
        off[r] = (long long)(r * ld);
}

}  // namespace skd

namespace skdi {

skd_status run_knn_sorted(skd_ctx *ctx, const skd_db *R, const skd_db *Q, bool self, const EpiParams &E, uint64_t knn,
                          uint64_t row_begin, uint64_t row_end, uint64_t *idx_out, float *d0_out, float *d1_out,
                          skd_loc out_loc, uint64_t *key_out, float *aux_out, uint64_t idx_offset) {
    const skd_db *A = self ? R : Q, *B = R;  // rows = the kNN rows (self: samples, cross: queries), columns = references
    const uint64_t n_rows = row_end - row_begin, n_cols = B->n;
    const bool with_aux = E.kind == KIND_CORE_ACC, partial = key_out != nullptr;
    const uint64_t elt = with_aux ? 8 : 4;
    const uint64_t ld = (n_cols + 3) / 4 * 4;
    timers_reset(ctx);
    if (n_rows == 0) return SKD_OK;
    if (!ctx->dense_v2 && tile_kernel_smem(E.nk, E.nbins >= 65536, false, 0, false) > kMaxSmem)
        return fail(ctx, SKD_ERR_UNSUPPORTED, "%d k-mer lengths need the dense kernel (unset SKD_DENSE_V1)", E.nk);
    // per row of a chunk: key matrix + sort keys and values, each twice (the sort ping-pongs between two buffers)
    const uint64_t per_row = ld * (elt + 2 * 8 + (with_aux ? 2 * 4 : 0));
    uint64_t rows_chunk = std::max<uint64_t>(TS, (2 * ctx->sym_chunk_bytes) / per_row / TS * TS);
    rows_chunk = std::min<uint64_t>(rows_chunk, (n_rows + TS - 1) / TS * TS);
    // the row lists (device): final destination for device outputs of partial calls, else scratch
    unsigned long long *lkey = (unsigned long long *)key_out;
    float *laux = aux_out;
    void *p;
    skd_status st;
    if (!partial || out_loc == SKD_HOST) {
        st = scratch_get(ctx, "sorted_lkey", n_rows * knn * 8, &p);
        if (st) return st;
        lkey = (unsigned long long *)p;
        laux = nullptr;
        if (with_aux) {
            st = scratch_get(ctx, "sorted_laux", n_rows * knn * 4, &p);
            if (st) return st;
            laux = (float *)p;
        }
    }
    void *D, *k0, *k1, *a0 = nullptr, *a1 = nullptr, *offs;
    if ((st = scratch_get(ctx, "knn_rect", rows_chunk * ld * elt + 64, &D))) return st;
    if ((st = scratch_get(ctx, "sorted_k0", rows_chunk * ld * 8, &k0))) return st;
    if ((st = scratch_get(ctx, "sorted_k1", rows_chunk * ld * 8, &k1))) return st;
    if (with_aux) {
        if ((st = scratch_get(ctx, "sorted_a0", rows_chunk * ld * 4, &a0))) return st;
        if ((st = scratch_get(ctx, "sorted_a1", rows_chunk * ld * 4, &a1))) return st;
    }
    if ((st = scratch_get(ctx, "sorted_off", (rows_chunk + 1) * 8, &offs))) return st;
    segment_offsets_kernel<<<64, 256, 0, ctx->compute>>>(rows_chunk, ld, (long long *)offs);
    ctx->launches++;
    size_t tmp_bytes = 0;
    {
        cub::DoubleBuffer<unsigned long long> kb((unsigned long long *)k0, (unsigned long long *)k1);
        cub::DoubleBuffer<float> vb((float *)a0, (float *)a1);
        if (with_aux)
            CU(ctx, cub::DeviceSegmentedSort::SortPairs(nullptr, tmp_bytes, kb, vb, (long long)(rows_chunk * ld),
                                                        (long long)rows_chunk, (const long long *)offs,
                                                        (const long long *)offs + 1, ctx->compute));
        else
            CU(ctx, cub::DeviceSegmentedSort::SortKeys(nullptr, tmp_bytes, kb, (long long)(rows_chunk * ld),
                                                       (long long)rows_chunk, (const long long *)offs,
                                                       (const long long *)offs + 1, ctx->compute));
    }
    void *tmp;
    if ((st = scratch_get(ctx, "sorted_tmp", tmp_bytes + 256, &tmp))) return st;

    for (uint64_t g0 = row_begin; g0 < row_end; g0 += rows_chunk) {
        const uint64_t g1 = std::min(row_end, g0 + rows_chunk), rows = g1 - g0;
        std::vector<uint64_t> off;
        const uint32_t rb0 = (uint32_t)(g0 / TS), n_rb = (uint32_t)((g1 + TS - 1) / TS) - rb0;
        build_groups(ctx->group, false, rb0, n_rb, B->dev.nb, &off, uses_dense_kernel(ctx, MODE_DENSE_CROSS, B, E));
        void *doff;
        if ((st = scratch_get(ctx, "grp_off", off.size() * sizeof(uint64_t), &doff))) return st;
        CU(ctx, cudaMemcpyAsync(doff, off.data(), off.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->compute));
        CU(ctx, cudaStreamSynchronize(ctx->compute));  // off is local; also fences reuse of the buffers
        if ((st = launch_dense(ctx, A, B, E, MODE_DENSE_CROSS, g0, g1, (float *)D, off, (const uint64_t *)doff, ld))) return st;
        if ((st = timer_begin(ctx))) return st;
        if (with_aux)
            sort_keys_kernel<true><<<4 * ctx->sm_count, 256, 0, ctx->compute>>>(
                (const float *)D, ld, g0, rows, n_cols, self, idx_offset, (unsigned long long *)k0, (float *)a0);
        else
            sort_keys_kernel<false><<<4 * ctx->sm_count, 256, 0, ctx->compute>>>(
                (const float *)D, ld, g0, rows, n_cols, self, idx_offset, (unsigned long long *)k0, nullptr);
        cub::DoubleBuffer<unsigned long long> kb((unsigned long long *)k0, (unsigned long long *)k1);
        cub::DoubleBuffer<float> vb((float *)a0, (float *)a1);
        size_t tb = tmp_bytes;
        if (with_aux)
            CU(ctx, cub::DeviceSegmentedSort::SortPairs(tmp, tb, kb, vb, (long long)(rows * ld), (long long)rows,
                                                        (const long long *)offs, (const long long *)offs + 1, ctx->compute));
        else
            CU(ctx, cub::DeviceSegmentedSort::SortKeys(tmp, tb, kb, (long long)(rows * ld), (long long)rows,
                                                       (const long long *)offs, (const long long *)offs + 1, ctx->compute));
        take_lists_kernel<<<4 * ctx->sm_count, 256, 0, ctx->compute>>>(kb.Current(), with_aux ? vb.Current() : nullptr, ld, rows,
                                                                      (uint32_t)knn, lkey + (g0 - row_begin) * knn,
                                                                      laux ? laux + (g0 - row_begin) * knn : nullptr);
        CU(ctx, cudaGetLastError());
        ctx->launches += 3;
        if ((st = timer_end(ctx))) return st;
    }
    if (partial) {
        if (out_loc == SKD_HOST) {
            CU(ctx, cudaMemcpyAsync(key_out, lkey, n_rows * knn * 8, cudaMemcpyDeviceToHost, ctx->compute));
            if (aux_out && with_aux) CU(ctx, cudaMemcpyAsync(aux_out, laux, n_rows * knn * 4, cudaMemcpyDeviceToHost, ctx->compute));
        }
        CU(ctx, cudaStreamSynchronize(ctx->compute));
        return timers_collect(ctx);
    }
    st = finalize_lists(ctx, lkey, laux, 1, n_rows * knn, n_rows, knn, E.kind, idx_out, d0_out, d1_out, out_loc);
    if (st) return st;
    return timers_collect(ctx);
}

}  // namespace skdi
