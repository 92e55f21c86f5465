// libsketchhost.so -- C ABI over the host-side C++ readers (no CUDA): `.skm` metadata (MultiSketch), `--subset`
// selection, completeness files, `.ski` inverted indexes and `.skq` bins.  The Python layer
// (sketchlib.rust_b200/host.py) binds it so that Python callers load databases the way the reference does
// (MultiSketch::load_metadata + read_sketch_data[_block], src/sketch/multisketch.rs:90-210; Inverted::load,
// src/inverted.rs:200-214) with the same code the C++ CLI runs.  Errors: functions return 0 on success and copy the
// message (the reference's panic / Err text) into `err`.
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "inverted.hpp"
#include "multisketch.hpp"

using namespace skdhost;

namespace {
int fail(char *err, size_t errlen, const std::string &msg) {
    if (err && errlen) snprintf(err, errlen, "%s", msg.c_str());
    return 1;
}

inline void append_line(std::string &buf, const char *a, const char *b, float d0, const float *d1) {
    buf += a;
    buf += '\t';
    buf += b;
    buf += '\t';
    buf += fmt_f32(d0);
    if (d1) {
        buf += '\t';
        buf += fmt_f32(*d1);
    }
    buf += '\n';
}

// Formats rows [0, n_rows) on `threads` host threads (row_fn appends the text of one row) and writes the blocks in
// row order.  The reference's Display impls are serial; the text is the same.
int write_rows(const char *path, int append, uint64_t n_rows, int threads,
               const std::function<void(uint64_t, std::string &)> &row_fn, char *err, size_t errlen) {
    FILE *f = fopen(path, append ? "ab" : "wb");
    if (!f) return fail(err, errlen, std::string("could not create ") + path);
    const uint64_t block = 32, wave = (uint64_t)std::max(1, threads) * 8;  // blocks formatted per write
    bool ok = true;
    for (uint64_t b0 = 0; b0 * block < n_rows && ok; b0 += wave) {
        const uint64_t nb = std::min<uint64_t>(wave, (n_rows + block - 1) / block - b0);
        std::vector<std::string> text(nb);
        std::atomic<uint64_t> next{0};
        auto work = [&] {
            for (;;) {
                const uint64_t b = next.fetch_add(1);
                if (b >= nb) break;
                for (uint64_t r = (b0 + b) * block; r < std::min(n_rows, (b0 + b + 1) * block); ++r) row_fn(r, text[b]);
            }
        };
        if (threads <= 1 || nb == 1) {
            work();
        } else {
            std::vector<std::thread> pool;
            for (int t = 0; t < threads; ++t) pool.emplace_back(work);
            for (auto &th : pool) th.join();
        }
        for (auto &t : text) ok = ok && fwrite(t.data(), 1, t.size(), f) == t.size();
    }
    ok = fclose(f) == 0 && ok;
    return ok ? 0 : fail(err, errlen, "Error writing output distances");
}
}  // namespace

extern "C" {

struct skh_multisketch {
    MultiSketch m;
};
struct skh_inverted {
    Inverted inv;
};

int skh_skm_load(const char *file_prefix, skh_multisketch **out, char *err, size_t errlen) {
    try {
        auto *h = new skh_multisketch{MultiSketch::load_metadata(strip_sketch_extension(file_prefix))};
        *out = h;
        return 0;
    } catch (const std::exception &e) {
        return fail(err, errlen, e.what());
    }
}
void skh_skm_free(skh_multisketch *h) { delete h; }
uint64_t skh_skm_n_loaded(const skh_multisketch *h) { return h->m.number_samples_loaded(); }
uint64_t skh_skm_n_total(const skh_multisketch *h) { return h->m.sketch_metadata.size(); }
uint32_t skh_skm_n_kmers(const skh_multisketch *h) { return (uint32_t)h->m.kmer_lengths.size(); }
void skh_skm_kmers(const skh_multisketch *h, uint64_t *out) {
    for (size_t i = 0; i < h->m.kmer_lengths.size(); ++i) out[i] = h->m.kmer_lengths[i];
}
uint64_t skh_skm_sketch_size(const skh_multisketch *h) { return h->m.sketch_size; }
uint64_t skh_skm_sketchsize64(const skh_multisketch *h) { return h->m.sketchsize64; }
uint64_t skh_skm_sample_stride(const skh_multisketch *h) { return h->m.sample_stride; }
uint64_t skh_skm_kmer_stride(const skh_multisketch *h) { return h->m.kmer_stride; }
const char *skh_skm_sketch_version(const skh_multisketch *h) { return h->m.sketch_version.c_str(); }
const char *skh_skm_hash_type(const skh_multisketch *h) { return h->m.hash_type.c_str(); }
// name of loaded sample idx (subset order after skh_skm_select_subset), multisketch.rs:139-144
const char *skh_skm_sketch_name(const skh_multisketch *h, uint64_t idx) {
    return idx < h->m.number_samples_loaded() ? h->m.sketch_name(idx).c_str() : "";
}
// positions in the .skd file (Sketch::index) of the loaded samples, in loaded order: what skd_db_load_skd takes
void skh_skm_file_indices(const skh_multisketch *h, uint64_t *out) {
    const MultiSketch &m = h->m;
    // full load: file POSITION like read_sketch_data (multisketch.rs:167-184); Sketch::index only after a subset
    for (size_t i = 0; i < m.number_samples_loaded(); ++i)
        out[i] = m.block_reindex ? m.sketch_metadata[(*m.block_reindex)[i]].index : (uint64_t)i;
}
// read_sketch_data_block's name lookup (multisketch.rs:187-200): afterwards the loaded samples are `names`, in order
int skh_skm_select_subset(skh_multisketch *h, const char *const *names, uint64_t n, char *err, size_t errlen) {
    try {
        std::vector<std::string> v(names, names + n);
        h->m.select_subset(v);
        return 0;
    } catch (const std::exception &e) {
        return fail(err, errlen, e.what());
    }
}
// read_completeness_file (src/io.rs:240-324): one value per loaded sample; warnings joined by '\n' into `warn`
int skh_read_completeness_file(const char *path, const skh_multisketch *h, double *out, char *warn, size_t warnlen,
                               char *err, size_t errlen) {
    try {
        std::vector<std::string> w;
        std::vector<double> c = read_completeness_file(path, h->m, &w);
        for (size_t i = 0; i < c.size(); ++i) out[i] = c[i];
        std::string joined;
        for (auto &s : w) joined += (joined.empty() ? "" : "\n") + s;
        if (warn && warnlen) snprintf(warn, warnlen, "%s", joined.c_str());
        return 0;
    } catch (const std::exception &e) {
        return fail(err, errlen, e.what());
    }
}

int skh_ski_load(const char *file_prefix, skh_inverted **out, char *err, size_t errlen) {
    try {
        auto *h = new skh_inverted{Inverted::load(strip_sketch_extension(file_prefix))};
        *out = h;
        return 0;
    } catch (const std::exception &e) {
        return fail(err, errlen, e.what());
    }
}
void skh_ski_free(skh_inverted *h) { delete h; }
uint64_t skh_ski_n_samples(const skh_inverted *h) { return h->inv.n_samples(); }
uint64_t skh_ski_sketch_size(const skh_inverted *h) { return h->inv.sketch_size(); }
uint64_t skh_ski_kmer(const skh_inverted *h) { return h->inv.kmer_size; }
int skh_ski_rc(const skh_inverted *h) { return h->inv.rc ? 1 : 0; }
const char *skh_ski_sample_name(const skh_inverted *h, uint64_t idx) {
    return idx < h->inv.n_samples() ? h->inv.sample_names[idx].c_str() : "";
}
// Inverted::any_shared_bins (src/inverted.rs:259-268): ascending sample indices; returns the count (at most cap
// are written)
uint64_t skh_ski_any_shared_bins(const skh_inverted *h, const uint16_t *sigs, uint32_t *out, uint64_t cap) {
    std::vector<uint32_t> v = h->inv.any_shared_bins(sigs);
    for (size_t i = 0; i < v.size() && i < cap; ++i) out[i] = v[i];
    return v.size();
}
// Inverted::query_against_inverted_index (src/inverted.rs:229-240): match counts for all n_samples
void skh_ski_match_counts(const skh_inverted *h, const uint16_t *sigs, uint32_t *out) {
    std::vector<uint32_t> v = h->inv.query_against_inverted_index(sigs);
    for (size_t i = 0; i < v.size(); ++i) out[i] = v[i];
}
uint64_t skh_ski_any_shared_bin_pairs(const skh_inverted *h) { return h->inv.any_shared_bin_pairs(); }

// ---- Display of DistanceMatrix (distance_matrix.rs:175-209) and SparseDistanceMatrix (:360-401), straight to a file.
// d: the f32 vector of rows [row_begin, row_end) (ncols values per pair).  Self mode (query_names == NULL): pairs
// (i, j > i) in condensed order; cross mode: ref-major, n_query values per reference row.
int skh_write_dense(const char *path, int append, const char *const *ref_names, uint64_t n_ref,
                    const char *const *query_names, uint64_t n_query, const float *d, int ncols, uint64_t row_begin,
                    uint64_t row_end, int threads, char *err, size_t errlen) {
    if (row_end > n_ref) row_end = n_ref;
    if (row_begin > row_end) row_begin = row_end;
    auto self_offset = [&](uint64_t r) {  // pairs of rows [row_begin, r) of the upper triangle
        auto before = [&](uint64_t x) { return x * n_ref - x * (x + 1) / 2; };
        return before(r) - before(row_begin);
    };
    return write_rows(path, append, row_end - row_begin, threads, [&](uint64_t r, std::string &txt) {
        const uint64_t i = row_begin + r;
        if (query_names) {
            const float *p = d + r * n_query * ncols;
            for (uint64_t j = 0; j < n_query; ++j, p += ncols)
                append_line(txt, ref_names[i], query_names[j], p[0], ncols == 2 ? p + 1 : nullptr);
        } else {
            const float *p = d + self_offset(i) * ncols;
            for (uint64_t j = i + 1; j < n_ref; ++j, p += ncols)
                append_line(txt, ref_names[i], ref_names[j], p[0], ncols == 2 ? p + 1 : nullptr);
        }
    }, err, errlen);
}
// row r of the matrix is sample row_names[row_begin + r]; neighbours index col_names.  Single-k rows skip padding
// entries (distance >= 1 to the row's own name, distance_matrix.rs:376-381).
int skh_write_sparse(const char *path, int append, const char *const *row_names, const char *const *col_names,
                     uint64_t row_begin, uint64_t n_rows, uint64_t knn, const uint64_t *idx, const float *d0,
                     const float *d1, int threads, char *err, size_t errlen) {
    return write_rows(path, append, n_rows, threads, [&](uint64_t r, std::string &txt) {
        const char *rn = row_names[row_begin + r];
        for (uint64_t t = 0; t < knn; ++t) {
            const char *cn = col_names[idx[r * knn + t]];
            if (d1)
                append_line(txt, rn, cn, d0[r * knn + t], &d1[r * knn + t]);
            else if (d0[r * knn + t] < 1.0f || strcmp(cn, rn) != 0)
                append_line(txt, rn, cn, d0[r * knn + t], nullptr);
        }
    }, err, errlen);
}

}  // extern "C"
