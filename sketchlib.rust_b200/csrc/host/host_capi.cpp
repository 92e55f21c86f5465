// libsketchhost.so -- C ABI over the host-side C++ readers (no CUDA): `.skm` metadata (MultiSketch), `--subset`
// selection, completeness files, `.ski` inverted indexes and `.skq` bins.  The Python layer
// (sketchlib.rust_b200/host.py) binds it so that Python callers load databases the way the reference does
// (MultiSketch::load_metadata + read_sketch_data[_block], src/sketch/multisketch.rs:90-210; Inverted::load,
// src/inverted.rs:200-214) with the same code the C++ CLI runs.  Errors: functions return 0 on success and copy the
// message (the reference's panic / Err text) into `err`.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "inverted.hpp"
#include "multisketch.hpp"

using namespace skdhost;

namespace {
int fail(char *err, size_t errlen, const std::string &msg) {
    if (err && errlen) snprintf(err, errlen, "%s", msg.c_str());
    return 1;
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
    for (size_t i = 0; i < m.number_samples_loaded(); ++i)
        out[i] = m.sketch_metadata[m.block_reindex ? (*m.block_reindex)[i] : i].index;
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

}  // extern "C"
