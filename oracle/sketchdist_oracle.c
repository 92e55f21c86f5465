/*
 * sketchdist_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C CPU restatement of the pairwise sketch-distance path of
 * bacpop/sketchlib.rust (v0.3.0).  It exists so that the CUDA path can be
 * checked against the reference's arithmetic without a Rust toolchain (none
 * exists in this image).  Nothing in the product (sketchlib.rust_b200/, the
 * C-ABI library) may include, link or call this file; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * Parity pin: tests/test_oracle_golden.py checks this file against every golden
 * stdout the reference's own tests hold for the path (dists_knn_ca/jaccard/ani,
 * dists_subset, inverted_precluster[_ani]) and the known-answer match counts
 * recorded in SURVEY.md.
 *
 * Each function cites the reference file:line it restates (paths relative to
 * the reference checkout).  Floating point is f64 in the reference's operation
 * order; build with -ffp-contract=off so no FMA contraction changes rounding.
 */
#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define BBITS 14 /* src/sketch/mod.rs:34 */
#define CHUNK_SIZE 1000 /* src/distances/mod.rs:20 */

/* A loaded MultiSketch as the hot path sees it: src/sketch/multisketch.rs:23-43, 213-219 */
typedef struct {
    const uint64_t *bins;   /* sketch_bins */
    uint64_t n_samples;     /* number_samples_loaded() */
    uint64_t sample_stride; /* in u64 */
    uint64_t kmer_stride;   /* in u64 = sketchsize64 * BBITS */
    uint64_t sketchsize64;
    uint64_t sketch_size;   /* number of bins */
    const uint64_t *kmers;  /* kmer_lengths */
    uint64_t n_k;
} oracle_db;

/* get_sketch_slice: src/sketch/multisketch.rs:213-219 */
static inline const uint64_t *slice(const oracle_db *db, uint64_t s, uint64_t k_idx) {
    return db->bins + s * db->sample_stride + k_idx * db->kmer_stride;
}

/* The integer core of jaccard_index: src/distances/jaccard.rs:15-25 */
uint32_t oracle_samebits(const uint64_t *s1, const uint64_t *s2, uint64_t sketchsize64) {
    uint32_t samebits = 0;
    for (uint64_t w = 0; w < sketchsize64; ++w) {
        uint64_t bits = ~(uint64_t)0;
        for (int p = 0; p < BBITS; ++p) bits &= ~(s1[w * BBITS + p] ^ s2[w * BBITS + p]);
        samebits += (uint32_t)__builtin_popcountll(bits);
    }
    return samebits;
}

/* completeness_correction: src/distances/jaccard.rs:55-57 */
static inline double completeness_correction(double j, double c1, double c2) {
    return j / (c1 * c2 / (c1 + c2 - c1 * c2));
}

/* The float tail of jaccard_index: src/distances/jaccard.rs:14, 26-44 */
double oracle_jaccard_from_samebits(uint32_t samebits, uint64_t sketchsize64, int has_c, double c1,
                                    double c2, double cutoff) {
    double unionsize = (double)(64 * sketchsize64);
    uint32_t maxnbits = (uint32_t)sketchsize64 * 64u;
    uint32_t expected = maxnbits >> BBITS;
    uint32_t diff = samebits > expected ? samebits - expected : 0; /* saturating_sub */
    double intersize = ((double)diff * (double)maxnbits) / (double)(maxnbits - expected);
    double j = intersize / unionsize;
    if (has_c) {
        if (c1 * c2 >= cutoff) {
            j = completeness_correction(j, c1, c2);
            j = j < 1.0 ? j : 1.0; /* f64::min(1.0) */
        }
    }
    return j;
}

/* jaccard_index: src/distances/jaccard.rs:6-45 */
double oracle_jaccard_index(const uint64_t *s1, const uint64_t *s2, uint64_t sketchsize64, int has_c,
                            double c1, double c2, double cutoff) {
    return oracle_jaccard_from_samebits(oracle_samebits(s1, s2, sketchsize64), sketchsize64, has_c, c1,
                                        c2, cutoff);
}

/* ani_pois: src/distances/jaccard.rs:49-51.  f64::max(0.0, NaN) = 0.0 like fmax. */
double oracle_ani_pois(double jaccard, double k) {
    return fmax(0.0, 1.0 + 1.0 / k * log((2.0 * jaccard) / (1.0 + jaccard)));
}

/* simple_linear_regression: src/distances/jaccard.rs:105-142 */
void oracle_simple_linear_regression(double xsum, double ysum, double xysum, double xsquaresum,
                                     double ysquaresum, double n, float *core_out, float *acc_out) {
    if (isnan(ysum) || ysum == -INFINITY || n < 3.0) {
        *core_out = 1.0f;
        *acc_out = 1.0f;
        return;
    }
    double xbar = xsum / n;
    double ybar = ysum / n;
    double x_diff = xsquaresum - xsum * xsum / n;
    double y_diff = ysquaresum - ysum * ysum / n;
    double xstddev = sqrt((xsquaresum - xsum * xsum / n) / n);
    double ystddev = sqrt((ysquaresum - ysum * ysum / n) / n);
    double r = (xysum - xsum * ysum / n) / sqrt(x_diff * y_diff);
    double beta = r * ystddev / xstddev;
    double alpha = -beta * xbar + ybar;
    double core = 0.0, acc = 0.0;
    if (beta < 0.0) {
        core = 1.0 - exp(beta);
    } else if (r > 0.0) {
        core = 1.0;
    }
    if (alpha < 0.0) acc = 1.0 - exp(alpha);
    *core_out = (float)core;
    *acc_out = (float)acc;
}

/* Regression from per-k match counts (the part of core_acc_dist after the bit counting);
 * used by the tests to check the CUDA epilogue in isolation.  src/distances/jaccard.rs:73-100 */
void oracle_core_acc_from_counts(const uint32_t *samebits, const uint64_t *kmers, uint64_t n_k,
                                 uint64_t sketchsize64, uint64_t sketch_size, int has_c, double c1,
                                 double c2, double cutoff, float *core_out, float *acc_out) {
    double xsum = 0, ysum = 0, xysum = 0, xsquaresum = 0, ysquaresum = 0, n = 0;
    double tolerance = log(2.0 / (double)(sketch_size * 64));
    for (uint64_t k_idx = 0; k_idx < n_k; ++k_idx) {
        double y = log(oracle_jaccard_from_samebits(samebits[k_idx], sketchsize64, has_c, c1, c2, cutoff));
        if (y < tolerance) break;
        double k_fl = (double)kmers[k_idx];
        xsum += k_fl;
        ysum += y;
        xysum += k_fl * y;
        xsquaresum += k_fl * k_fl;
        ysquaresum += y * y;
        n += 1.0;
    }
    oracle_simple_linear_regression(xsum, ysum, xysum, xsquaresum, ysquaresum, n, core_out, acc_out);
}

/* core_acc_dist: src/distances/jaccard.rs:61-101.  Returns -1 where the reference panics (K<2). */
int oracle_core_acc_dist(const oracle_db *ref, const oracle_db *query, uint64_t ref_idx, uint64_t query_idx,
                         const double *ref_c, const double *query_c, double cutoff, float *core_out,
                         float *acc_out) {
    if (ref->n_k < 2) return -1;
    double xsum = 0, ysum = 0, xysum = 0, xsquaresum = 0, ysquaresum = 0, n = 0;
    double tolerance = log(2.0 / (double)(ref->sketch_size * 64));
    for (uint64_t k_idx = 0; k_idx < ref->n_k; ++k_idx) {
        int has_c = ref_c != NULL && query_c != NULL;
        double c1 = ref_c ? ref_c[ref_idx] : 0.0, c2 = query_c ? query_c[query_idx] : 0.0;
        double y = log(oracle_jaccard_index(slice(ref, ref_idx, k_idx), slice(query, query_idx, k_idx),
                                            ref->sketchsize64, has_c, c1, c2, cutoff));
        if (y < tolerance) break;
        double k_fl = (double)ref->kmers[k_idx];
        xsum += k_fl;
        ysum += y;
        xysum += k_fl * y;
        xsquaresum += k_fl * k_fl;
        ysquaresum += y * y;
        n += 1.0;
    }
    oracle_simple_linear_regression(xsum, ysum, xysum, xsquaresum, ysquaresum, n, core_out, acc_out);
    return 0;
}

/* Index maths: src/distances/distance_matrix.rs:11-51 */
uint64_t oracle_square_to_condensed(uint64_t i, uint64_t j, uint64_t n) {
    return n * i - ((i * (i + 1)) >> 1) + j - 1 - i;
}
uint64_t oracle_calc_row_idx(uint64_t k, uint64_t n) {
    int64_t k_i = (int64_t)k, n_i = (int64_t)n;
    return n - 2 - (uint64_t)floor(sqrt((double)(-8 * k_i + 4 * n_i * (n_i - 1) - 7)) / 2.0 - 0.5);
}
uint64_t oracle_calc_col_idx(uint64_t k, uint64_t i, uint64_t n) {
    int64_t k_i = (int64_t)k, i_i = (int64_t)i, n_i = (int64_t)n;
    return (uint64_t)(k_i + i_i + 1 - n_i * (n_i - 1) / 2 + (n_i - i_i) * ((n_i - i_i) - 1) / 2);
}

/* ------------------------------------------------------------------ */
/* A tiny "rayon": threads pull task indices from an atomic counter.  */
typedef void (*task_fn)(void *ctx, uint64_t task);
typedef struct {
    task_fn fn;
    void *ctx;
    atomic_ullong *next;
    uint64_t n_tasks;
} pool_arg;
static void *pool_worker(void *p) {
    pool_arg *a = (pool_arg *)p;
    for (;;) {
        uint64_t t = atomic_fetch_add(a->next, 1);
        if (t >= a->n_tasks) break;
        a->fn(a->ctx, t);
    }
    return NULL;
}
static void run_tasks(task_fn fn, void *ctx, uint64_t n_tasks, int threads) {
    atomic_ullong next = 0;
    pool_arg a = {fn, ctx, &next, n_tasks};
    if (threads <= 1) {
        pool_worker(&a);
        return;
    }
    pthread_t *tids = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    for (int t = 0; t < threads; ++t) pthread_create(&tids[t], NULL, pool_worker, &a);
    for (int t = 0; t < threads; ++t) pthread_join(tids[t], NULL);
    free(tids);
}

/* mode: k_idx < 0 => CoreAcc, else Jaccard(k_idx, k, ani)  (DistType, distance_matrix.rs:55-60) */
typedef struct {
    const oracle_db *ref, *query; /* query == NULL: self mode */
    int64_t k_idx;
    int ani;
    const double *ref_c, *query_c;
    double cutoff;
    float *out;
    uint64_t n_dist;      /* total distances */
    uint64_t first_dist;  /* offset of out[0] in the full condensed / rectangular index */
    /* knn */
    uint64_t knn;
    uint64_t *idx_out;
    float *d0_out, *d1_out;
    uint64_t row_begin;
} dist_job;

/* One dense distance, value rules of src/distances/mod.rs:82-114 and :251-283 */
static void dense_one(const dist_job *jb, uint64_t i, uint64_t j, float *dst) {
    const oracle_db *q = jb->query ? jb->query : jb->ref;
    if (jb->k_idx >= 0) {
        int has_c = jb->ref_c != NULL && jb->query_c != NULL;
        double c1 = jb->ref_c ? jb->ref_c[i] : 0.0, c2 = jb->query_c ? jb->query_c[j] : 0.0;
        double jac = oracle_jaccard_index(slice(jb->ref, i, (uint64_t)jb->k_idx), slice(q, j, (uint64_t)jb->k_idx),
                                          jb->ref->sketchsize64, has_c, c1, c2, jb->cutoff);
        double kf = (double)jb->ref->kmers[jb->k_idx];
        dst[0] = jb->ani ? (float)oracle_ani_pois(jac, kf) : (float)(1.0 - jac);
    } else {
        oracle_core_acc_dist(jb->ref, q, i, j, jb->ref_c, jb->query_c, jb->cutoff, &dst[0], &dst[1]);
    }
}

/* self_dists_all chunk body: src/distances/mod.rs:77-128 */
static void self_all_chunk(void *ctx, uint64_t chunk_idx) {
    const dist_job *jb = (const dist_job *)ctx;
    uint64_t n = jb->ref->n_samples;
    int ncols = jb->k_idx >= 0 ? 1 : 2;
    uint64_t start = jb->first_dist + chunk_idx * CHUNK_SIZE;
    uint64_t i = oracle_calc_row_idx(start, n);
    uint64_t j = oracle_calc_col_idx(start, i, n);
    for (uint64_t d = 0; d < CHUNK_SIZE; ++d) {
        uint64_t local = chunk_idx * CHUNK_SIZE + d;
        if (local >= jb->n_dist) break;
        dense_one(jb, i, j, jb->out + local * (uint64_t)ncols);
        j += 1;
        if (j >= n) {
            i += 1;
            j = i + 1;
            if (i >= n - 1) break;
        }
    }
}

/* cross_dists_all chunk body: src/distances/mod.rs:249-295 (ref-major: idx = i*n_query + j) */
static void cross_all_chunk(void *ctx, uint64_t chunk_idx) {
    const dist_job *jb = (const dist_job *)ctx;
    uint64_t n = jb->ref->n_samples, nq = jb->query->n_samples;
    int ncols = jb->k_idx >= 0 ? 1 : 2;
    uint64_t start = jb->first_dist + chunk_idx * CHUNK_SIZE;
    uint64_t i = start / nq, j = start % nq; /* calc_query_indices */
    for (uint64_t d = 0; d < CHUNK_SIZE; ++d) {
        uint64_t local = chunk_idx * CHUNK_SIZE + d;
        if (local >= jb->n_dist) break;
        dense_one(jb, i, j, jb->out + local * (uint64_t)ncols);
        j += 1;
        if (j >= nq) {
            i += 1;
            j = 0;
            if (i >= n) break;
        }
    }
}

/* Dense self mode over condensed indices [first_dist, first_dist + n_dist).
 * first_dist = 0, n_dist = n(n-1)/2 is the reference's self_dists_all (mod.rs:58-130);
 * first_dist must be a multiple of CHUNK_SIZE to keep the reference's chunking. */
int oracle_self_dists_all(const oracle_db *db, int64_t k_idx, int ani, const double *completeness,
                          double cutoff, uint64_t first_dist, uint64_t n_dist, float *out, int threads) {
    if (k_idx < 0 && db->n_k < 2) return -1;
    dist_job jb = {db, NULL, k_idx, ani, completeness, completeness, cutoff, out, n_dist, first_dist, 0, 0, 0, 0, 0};
    run_tasks(self_all_chunk, &jb, (n_dist + CHUNK_SIZE - 1) / CHUNK_SIZE, threads);
    return 0;
}

int oracle_cross_dists_all(const oracle_db *ref, const oracle_db *query, int64_t k_idx, int ani,
                           const double *ref_c, const double *query_c, double cutoff, uint64_t first_dist,
                           uint64_t n_dist, float *out, int threads) {
    if (k_idx < 0 && ref->n_k < 2) return -1;
    dist_job jb = {ref, query, k_idx, ani, ref_c, query_c, cutoff, out, n_dist, first_dist, 0, 0, 0, 0, 0};
    run_tasks(cross_all_chunk, &jb, (n_dist + CHUNK_SIZE - 1) / CHUNK_SIZE, threads);
    return 0;
}

/* ------------------------------------------------------------------ */
/* Bounded max-heap with the algorithm of Rust's std BinaryHeap (push = sift_up,
 * pop = swap-remove + sift_down_to_bottom, into_sorted_vec = repeated swap +
 * sift_down_range), so that even tie handling follows the reference
 * (push_heap: src/distances/mod.rs:41-48; Ord on field 1 only:
 * src/distances/distance_matrix.rs:214-266). */
typedef struct {
    uint64_t idx;
    float key, aux;
} hitem;

static void sift_up(hitem *h, size_t start, size_t pos) {
    hitem e = h[pos];
    while (pos > start) {
        size_t parent = (pos - 1) / 2;
        if (e.key <= h[parent].key) break;
        h[pos] = h[parent];
        pos = parent;
    }
    h[pos] = e;
}
static void sift_down_range(hitem *h, size_t pos, size_t end) {
    hitem e = h[pos];
    size_t child = 2 * pos + 1;
    while (child + 1 < end) { /* child <= end.saturating_sub(2) */
        if (h[child].key <= h[child + 1].key) child += 1;
        if (e.key >= h[child].key) {
            h[pos] = e;
            return;
        }
        h[pos] = h[child];
        pos = child;
        child = 2 * pos + 1;
    }
    if (child == end - 1 && e.key < h[child].key) {
        h[pos] = h[child];
        pos = child;
    }
    h[pos] = e;
}
static void sift_down_to_bottom(hitem *h, size_t pos, size_t end) {
    size_t start = pos;
    hitem e = h[pos];
    size_t child = 2 * pos + 1;
    while (child + 1 < end) {
        if (h[child].key <= h[child + 1].key) child += 1;
        h[pos] = h[child];
        pos = child;
        child = 2 * pos + 1;
    }
    if (child == end - 1) {
        h[pos] = h[child];
        pos = child;
    }
    h[pos] = e;
    sift_up(h, start, pos);
}
static void heap_push_bounded(hitem *h, size_t *len, hitem it, size_t knn) {
    if (*len < knn || it.key < h[0].key) {
        h[*len] = it;
        sift_up(h, 0, *len);
        *len += 1;
        if (*len > knn) { /* pop */
            *len -= 1;
            hitem last = h[*len];
            if (*len > 0) {
                h[0] = last; /* swap(&mut item, &mut data[0]); the popped max is dropped */
                sift_down_to_bottom(h, 0, *len);
            }
        }
    }
}
static void heap_into_sorted(hitem *h, size_t len) {
    size_t end = len;
    while (end > 1) {
        end -= 1;
        hitem t = h[0];
        h[0] = h[end];
        h[end] = t;
        sift_down_range(h, 0, end);
    }
}

/* One kNN row.  Self mode: src/distances/mod.rs:154-192 / 201-220; cross mode: :339-367 / :377-391 */
static void knn_row(void *ctx, uint64_t task) {
    const dist_job *jb = (const dist_job *)ctx;
    uint64_t row = jb->row_begin + task;
    int self_mode = jb->query == NULL;
    uint64_t n = jb->ref->n_samples;
    size_t knn = (size_t)jb->knn;
    hitem *heap = (hitem *)malloc(sizeof(hitem) * (knn + 1));
    size_t len = 0;
    for (uint64_t j = 0; j < n; ++j) {
        if (self_mode && j == row) continue;
        hitem it;
        it.idx = j;
        it.aux = 0.0f;
        if (jb->k_idx >= 0) {
            const uint64_t *row_sketch = slice(self_mode ? jb->ref : jb->query, row, (uint64_t)jb->k_idx);
            int has_c = jb->ref_c != NULL && jb->query_c != NULL;
            /* self: c1 = cv[i], c2 = cv[j]; cross: c1 = query_c[qi], c2 = ref_c[ri] */
            double c1 = self_mode ? (jb->ref_c ? jb->ref_c[row] : 0.0) : (jb->query_c ? jb->query_c[row] : 0.0);
            double c2 = jb->ref_c ? jb->ref_c[j] : 0.0;
            double jac = oracle_jaccard_index(row_sketch, slice(jb->ref, j, (uint64_t)jb->k_idx),
                                              jb->ref->sketchsize64, has_c, c1, c2, jb->cutoff);
            double kf = (double)jb->ref->kmers[jb->k_idx];
            it.key = jb->ani ? (float)(1.0 - oracle_ani_pois(jac, kf)) : (float)(1.0 - jac);
        } else if (self_mode) {
            oracle_core_acc_dist(jb->ref, jb->ref, row, j, jb->ref_c, jb->ref_c, jb->cutoff, &it.key, &it.aux);
        } else {
            oracle_core_acc_dist(jb->ref, jb->query, j, row, jb->ref_c, jb->query_c, jb->cutoff, &it.key, &it.aux);
        }
        heap_push_bounded(heap, &len, it, knn);
    }
    heap_into_sorted(heap, len);
    for (size_t t = 0; t < len; ++t) {
        uint64_t o = task * jb->knn + t;
        jb->idx_out[o] = heap[t].idx;
        jb->d0_out[o] = (jb->k_idx >= 0 && jb->ani) ? 1.0f - heap[t].key : heap[t].key;
        if (jb->d1_out) jb->d1_out[o] = heap[t].aux;
    }
    free(heap);
}

/* self_dists_knn (src/distances/mod.rs:133-224) for rows [row_begin, row_end).
 * The caller applies the CLI's clamp knn = n-1 when knn >= n (src/lib.rs:379-382). */
int oracle_self_dists_knn(const oracle_db *db, uint64_t knn, int64_t k_idx, int ani, const double *completeness,
                          double cutoff, uint64_t row_begin, uint64_t row_end, uint64_t *idx_out, float *d0_out,
                          float *d1_out, int threads) {
    if (k_idx < 0 && db->n_k < 2) return -1;
    if (knn == 0 || knn > db->n_samples - 1) return -2;
    dist_job jb = {db, NULL, k_idx, ani, completeness, completeness, cutoff, NULL, 0, 0, knn, idx_out, d0_out, d1_out, row_begin};
    run_tasks(knn_row, &jb, row_end - row_begin, threads);
    return 0;
}

/* cross_dists_knn (src/distances/mod.rs:306-395): one row per query; knn = min(knn, n_ref) is
 * applied here like the reference (:325); returns the clamped knn through *knn_used. */
int oracle_cross_dists_knn(const oracle_db *ref, const oracle_db *query, uint64_t knn, int64_t k_idx, int ani,
                           const double *ref_c, const double *query_c, double cutoff, uint64_t row_begin,
                           uint64_t row_end, uint64_t *idx_out, float *d0_out, float *d1_out, uint64_t *knn_used,
                           int threads) {
    if (ref->n_samples == 0 || query->n_samples == 0) return -3; /* reference panics: mod.rs:318-323 */
    if (k_idx < 0 && ref->n_k < 2) return -1;
    if (knn > ref->n_samples) knn = ref->n_samples;
    if (knn_used) *knn_used = knn;
    dist_job jb = {ref, query, k_idx, ani, ref_c, query_c, cutoff, NULL, 0, 0, knn, idx_out, d0_out, d1_out, row_begin};
    run_tasks(knn_row, &jb, row_end - row_begin, threads);
    return 0;
}

/* Raw per-k match counts for explicit (ref_idx, query_idx) pairs; the integer parity target. */
void oracle_match_counts(const oracle_db *ref, const oracle_db *query, const uint64_t *pairs, uint64_t n_pairs,
                         uint32_t *counts) {
    for (uint64_t p = 0; p < n_pairs; ++p)
        for (uint64_t k = 0; k < ref->n_k; ++k)
            counts[p * ref->n_k + k] =
                oracle_samebits(slice(ref, pairs[2 * p], k), slice(query, pairs[2 * p + 1], k), ref->sketchsize64);
}

/* self_dists_knn_precluster: src/distances/mod.rs:399-553 (single-k modes only; CoreAcc is unimplemented!
 * there, :548-550).  The inverted index is represented by its effect on this function: for row i the
 * candidates cand_idx[cand_off[i] .. cand_off[i+1]) are the samples `any_shared_bins` returns
 * (src/inverted.rs:259-268), already mapped to .skd indices and in the reference's iteration order (ascending
 * .ski index); the row itself may be listed and is skipped (:459-461).
 * retain: 0 = None, 1 = Singleton, 2 = Bruteforce (src/cli.rs RetainUnmatched; mod.rs:489-527). */
int oracle_self_dists_knn_precluster(const oracle_db *db, const uint64_t *cand_off, const uint64_t *cand_idx,
                                     uint64_t knn, int64_t k_idx, int ani, const double *completeness, double cutoff,
                                     int retain, uint64_t *idx_out, float *d0_out) {
    if (k_idx < 0) return -4;
    uint64_t n = db->n_samples;
    double kf = (double)db->kmers[k_idx];
    hitem *heap = (hitem *)malloc(sizeof(hitem) * (knn + 1));
    for (uint64_t i = 0; i < n; ++i) {
        size_t len = 0;
        const uint64_t *i_sketch = slice(db, i, (uint64_t)k_idx);
        for (int pass = 0; pass < 2; ++pass) {
            if (pass == 1 && !(len == 0 && retain == 2)) break; /* bruteforce only when nothing matched */
            uint64_t lo = pass == 0 ? cand_off[i] : 0, hi = pass == 0 ? cand_off[i + 1] : n;
            for (uint64_t c = lo; c < hi; ++c) {
                uint64_t j = pass == 0 ? cand_idx[c] : c;
                if (j == i) continue;
                int has_c = completeness != NULL;
                double jac = oracle_jaccard_index(i_sketch, slice(db, j, (uint64_t)k_idx), db->sketchsize64, has_c,
                                                  has_c ? completeness[i] : 0.0, has_c ? completeness[j] : 0.0, cutoff);
                hitem it;
                it.idx = j;
                it.aux = 0.0f;
                it.key = ani ? (float)(1.0 - oracle_ani_pois(jac, kf)) : (float)(1.0 - jac);
                heap_push_bounded(heap, &len, it, (size_t)knn);
            }
        }
        uint64_t *ri = idx_out + i * knn;
        float *rd = d0_out + i * knn;
        if (len == 0 && retain == 1) { /* Singleton: (i, 0.0) then (i, 1.0) padding, no ANI transform */
            for (uint64_t t = 0; t < knn; ++t) {
                ri[t] = i;
                rd[t] = t == 0 ? 0.0f : 1.0f;
            }
            continue;
        }
        heap_into_sorted(heap, len);
        for (uint64_t t = 0; t < knn; ++t) {
            if (t < len) {
                ri[t] = heap[t].idx;
                rd[t] = ani ? 1.0f - heap[t].key : heap[t].key;
            } else { /* null values at the end: (i, 1.0) */
                ri[t] = i;
                rd[t] = 1.0f;
            }
        }
    }
    free(heap);
    return 0;
}
