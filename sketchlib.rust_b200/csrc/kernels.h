// Host-callable launchers of the kernels in kernels.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "layout.h"

namespace skd {

void launch_repack(const uint64_t *raw, uint64_t raw_samples, uint64_t sample_stride, uint64_t kmer_stride,
                   uint32_t *tiled, const DevDb &db, uint32_t b0, uint32_t n_blocks, cudaStream_t st);

void launch_untile(const DevDb &db, uint64_t *rows, cudaStream_t st);  // tiled -> sample-major u64 [sample][k][w][plane]

int counts_per_pair(int nk, bool count32);
size_t tile_kernel_smem(int nk, bool count32, bool knn_mode, uint32_t cap, bool with_aux);

size_t dense_kernel_smem();
size_t dense_kernel_scratch_bytes(int kp, bool count32, int sm_count);
cudaError_t launch_tiles(const DevDb &A, const DevDb &B, const EpiParams &E, const SchedParams &S, uint64_t n_ctas,
                         bool count32, size_t smem, cudaStream_t st);

void launch_knn_merge(const unsigned long long *part_key, const float *part_aux, uint32_t n_split, uint32_t knn,
                      uint32_t cap, uint64_t first_lrow, uint64_t n_rows, int kind, unsigned long long *idx_out, float *d0_out,
                      float *d1_out, unsigned long long *key_out, float *aux_out, cudaStream_t st);

// Symmetric self-kNN second stage: folds the rect key matrix D of row group [g0, g0 + R) (leading dimension ld,
// first column g0) into the per-sample lists Lkey/Laux [n][knn].  cols = false: targets are the rows [t0, t1) of the
// group; cols = true: targets are the columns [t0, t1).
cudaError_t launch_select_sym(bool cols, bool aux, const float *D, uint64_t ld, uint64_t g0, uint32_t R, uint64_t n,
                              uint64_t t0, uint64_t t1, unsigned long long *Lkey, float *Laux, uint32_t knn,
                              cudaStream_t st);
// Cross mode (few query columns): 8 targets per CTA, rows of every split divided over its warps (select_few_kernel).
cudaError_t launch_select_few(bool aux, const float *D, uint64_t ld, uint64_t g0, uint32_t R, uint32_t nq,
                              unsigned long long *Lkey, float *Laux, uint32_t knn, cudaStream_t st, uint64_t idx_offset,
                              uint32_t n_splits, uint32_t rows_per_split, uint64_t part_stride);

void launch_knn_finalize(const unsigned long long *keys, const float *aux, uint32_t n_parts, uint64_t part_stride,
                         uint64_t row0, uint64_t n_rows, uint32_t knn, int kind, unsigned long long *idx_out,
                         float *d0_out, float *d1_out, unsigned long long *key_out, float *aux_out, cudaStream_t st);

void launch_fill_u64(unsigned long long *p, uint64_t n, unsigned long long v, cudaStream_t st);

cudaError_t launch_knn_candidates(const DevDb &A, const EpiParams &E, const uint64_t *cand_off, const uint64_t *cand_idx,
                                  uint64_t n_rows, uint32_t knn, uint32_t cap, int retain, unsigned long long *idx_out,
                                  float *d0_out, cudaStream_t st);

// Loads every kernel of kernels.cu into the current context (see the comment at the definition).
cudaError_t preload_kernels();

void launch_pairs(const DevDb &A, const DevDb &B, const EpiParams &E, const uint64_t *pairs, uint64_t n_pairs,
                  uint32_t *counts_out, float *dist_out, cudaStream_t st);

}  // namespace skd
