// `sketchlib dist` -- drop-in front end for the B200 distance path (host-only C++17 above the C ABI).
//
// Mirrors the reference's Dist sub-command: flags of src/cli.rs:186-231 (+ the global -v/--verbose, --quiet
// of src/cli.rs:88-104), control flow of src/lib.rs:303-453 (extension stripping, --subset via
// read_sketch_data_block, completeness files, set_k, knn clamping :379-382 / :411-414), and the text
// writers of src/distances/distance_matrix.rs:175-209 (dense) and :360-401 (sparse), byte for byte.
// `dist` and `inverted precluster` are the accelerated paths; `sketch` and `inverted build` (DNA FASTA only, CPU,
// sketcher.cpp / inverted.cpp) are shipped so that inputs can be made without the Rust binary.  The other
// sub-commands are not implemented.
// Extra flag: --gpu <id> (default 0).  --threads is accepted for compatibility (the work runs on the GPU).
#include <algorithm>
#include <array>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <fstream>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/sketchdist.h"
#include "inverted.hpp"
#include "multisketch.hpp"

using namespace skdhost;

namespace skdhost {
void sketch_files_dna(const std::string &output_prefix,
                      const std::vector<std::pair<std::string, std::vector<std::string>>> &inputs,
                      std::vector<uint64_t> kmers, uint64_t sketch_size, bool rc, int threads);
}

namespace {

int g_level = 1;  // 0 = error (--quiet), 1 = warn (default), 2 = info (-v)
void log_info(const std::string &m) {
    if (g_level >= 2) fprintf(stderr, "INFO  [sketchlib] %s\n", m.c_str());
}
void log_warn(const std::string &m) {
    if (g_level >= 1) fprintf(stderr, "WARN  [sketchlib] %s\n", m.c_str());
}

struct Panic : std::runtime_error {  // where the reference panics (exit code 101)
    using std::runtime_error::runtime_error;
};
struct UsageError : std::runtime_error {  // clap errors (exit code 2)
    using std::runtime_error::runtime_error;
};

struct DistArgs {
    std::string ref_db;
    std::string query_db;
    bool has_query = false;
    std::string output;
    bool has_output = false;
    long long knn = -1;
    std::string subset;
    long long kmer = -1;
    bool ani = false;
    long long threads = 1;
    std::string ref_completeness_file, query_completeness_file;
    double completeness_cutoff = 0.64;
    int gpu = 0;
    int gpus = 1;
    long long ref_shard_mb = 4096;  // cross --knn: references are loaded and compared in shards of this size
    bool verbose = false, quiet = false;
};

const char *USAGE =
    "Usage: sketchlib dist [OPTIONS] <REF_DB> [QUERY_DB]\n\n"
    "Arguments:\n"
    "  <REF_DB>    The .skm file used as the reference\n"
    "  [QUERY_DB]  The .skm file used as the query (omit for ref v ref)\n\n"
    "Options:\n"
    "  -o <OUTPUT>                        Output filename (omit to output to stdout)\n"
    "      --knn <KNN>                    Calculate sparse distances with k nearest-neighbours\n"
    "      --subset <SUBSET>              Sample names to analyse\n"
    "  -k <KMER>                          K-mer length (if provided only calculate Jaccard distance)\n"
    "      --ani                          Calculate ANI rather than Jaccard dists, using Poisson model\n"
    "      --threads <THREADS>            Number of CPU threads [default: 1] (formats the dense text output; distances run on the GPU)\n"
    "      --ref-completeness-file <F>    Reference completeness estimates (genome_id<tab>completeness)\n"
    "      --query-completeness-file <F>  Query completeness estimates (cross-query mode only)\n"
    "      --completeness-cutoff <X>      Minimum c1*c2 for the completeness correction [default: 0.64]\n"
    "      --gpu <ID>                     CUDA device [default: 0]\n"
    "      --gpus <N>                     Use devices 0..N-1 (one host thread and context per device, joined in a peer group)\n"
    "      --ref-shard-mb <MB>            Cross --knn: load and compare the references in shards of this size, the next\n"
    "                                     shard loading while the previous one is compared [default: 4096]\n"
    "  -v, --verbose                      Show progress messages\n"
    "      --quiet                        Don't show any messages\n";

long long parse_uint(const std::string &s, const char *what) {
    if (s.empty() || s.find_first_not_of("0123456789") != std::string::npos)
        throw UsageError(std::string("invalid value '") + s + "' for '" + what + "'");
    return atoll(s.c_str());
}

DistArgs parse_args(int argc, char **argv) {
    DistArgs a;
    std::vector<std::string> pos;
    bool saw_dist = false;
    for (int i = 1; i < argc; ++i) {
        std::string t = argv[i];
        auto value = [&](const std::string &name) -> std::string {  // "--opt v", "--opt=v", "-ov", "-o v", "-o=v"
            if (t.rfind("--", 0) == 0) {
                size_t eq = t.find('=');
                if (eq != std::string::npos) return t.substr(eq + 1);
            } else if (t.size() > 2) {
                return t[2] == '=' ? t.substr(3) : t.substr(2);
            }
            if (i + 1 >= argc) throw UsageError("a value is required for '" + name + "' but none was supplied");
            return argv[++i];
        };
        auto is = [&](const char *name) { return t == name || t.rfind(std::string(name) + "=", 0) == 0; };
        if (t == "-h" || t == "--help") {
            fputs(USAGE, stdout);
            exit(0);
        } else if (t == "-v" || t == "--verbose") {
            a.verbose = true;
        } else if (t == "--quiet") {
            a.quiet = true;
        } else if (t == "--ani") {
            a.ani = true;
        } else if (is("--knn")) {
            a.knn = parse_uint(value("--knn <KNN>"), "--knn <KNN>");
        } else if (is("--subset")) {
            a.subset = value("--subset <SUBSET>");
        } else if (is("--threads")) {
            a.threads = parse_uint(value("--threads <THREADS>"), "--threads <THREADS>");
            if (a.threads < 1) throw UsageError("invalid value for '--threads <THREADS>': Threads must be one or higher");
        } else if (is("--ref-completeness-file")) {
            a.ref_completeness_file = value("--ref-completeness-file");
        } else if (is("--query-completeness-file")) {
            a.query_completeness_file = value("--query-completeness-file");
        } else if (is("--completeness-cutoff")) {
            std::string v = value("--completeness-cutoff");
            char *end = nullptr;
            a.completeness_cutoff = strtod(v.c_str(), &end);
            if (v.empty() || end != v.c_str() + v.size())
                throw UsageError("invalid value '" + v + "' for '--completeness-cutoff <COMPLETENESS_CUTOFF>'");
        } else if (is("--ref-shard-mb")) {
            a.ref_shard_mb = parse_uint(value("--ref-shard-mb <MB>"), "--ref-shard-mb <MB>");
            if (a.ref_shard_mb < 1) throw UsageError("invalid value for '--ref-shard-mb <MB>': at least 1");
        } else if (is("--gpus")) {
            a.gpus = (int)parse_uint(value("--gpus <N>"), "--gpus <N>");
            if (a.gpus < 1) throw UsageError("invalid value for '--gpus <N>': at least one device is needed");
        } else if (is("--gpu")) {
            a.gpu = (int)parse_uint(value("--gpu <ID>"), "--gpu <ID>");
        } else if (t.rfind("-o", 0) == 0 && t.rfind("--", 0) != 0) {
            a.output = value("-o <OUTPUT>");
            a.has_output = true;
        } else if (t.rfind("-k", 0) == 0 && t.rfind("--", 0) != 0) {
            a.kmer = parse_uint(value("-k <KMER>"), "-k <KMER>");
        } else if (t.size() > 1 && t[0] == '-') {
            throw UsageError("unexpected argument '" + t + "' found");
        } else if (!saw_dist) {
            if (t != "dist")
                throw UsageError("unrecognized subcommand '" + t +
                                 "': this build implements only `dist`, `inverted` (the B200-accelerated paths), `sketch` (DNA) and `info`");
            saw_dist = true;
        } else {
            pos.push_back(t);
        }
    }
    if (!saw_dist) throw UsageError("a subcommand is required: sketchlib dist ...");
    if (pos.empty()) throw UsageError("the following required arguments were not provided:\n  <REF_DB>");
    if (pos.size() > 2) throw UsageError("unexpected argument '" + pos[2] + "' found");
    a.ref_db = pos[0];
    if (pos.size() == 2) {
        a.query_db = pos[1];
        a.has_query = true;
    }
    if (a.ani && a.kmer < 0)
        throw UsageError("the following required arguments were not provided:\n  -k <KMER>");  // requires("kmer")
    return a;
}

struct Gpu {
    skd_ctx *ctx = nullptr;
    ~Gpu() {
        if (ctx) skd_ctx_destroy(ctx);
    }
    void check(skd_status st) const {
        if (st == SKD_OK) return;
        std::string msg = skd_last_error(ctx);
        if (st == SKD_ERR_NEED_TWO_K || st == SKD_ERR_EMPTY_DB || st == SKD_ERR_IO) throw Panic(msg);
        throw std::runtime_error(msg);
    }
};

struct Loaded {
    MultiSketch meta;
    skd_db *db = nullptr;
    ~Loaded() {
        if (db) skd_db_free(db);
    }
};

void load_db(const Gpu &gpu, Loaded &L, const std::string &prefix, const std::vector<uint64_t> *file_indices) {
    const MultiSketch &m = L.meta;
    const std::string skd = prefix + ".skd";
    const uint64_t n = file_indices ? file_indices->size() : m.sketch_metadata.size();
    gpu.check(skd_db_load_skd(gpu.ctx, skd.c_str(), file_indices ? file_indices->data() : nullptr, n, m.sample_stride,
                              m.kmer_stride, m.sketchsize64, m.sketch_size, m.kmer_lengths.data(),
                              (uint32_t)m.kmer_lengths.size(), &L.db));
}

struct Out {
    FILE *f = stdout;
    std::string buf;
    explicit Out(const DistArgs &a) {
        if (a.has_output) {
            f = fopen(a.output.c_str(), "wb");
            if (!f) throw Panic("could not create " + a.output);
        }
        buf.reserve(1 << 22);
    }
    ~Out() {
        flush();
        if (f != stdout) fclose(f);
    }
    void flush() {
        if (!buf.empty() && fwrite(buf.data(), 1, buf.size(), f) != buf.size()) {
            fprintf(stderr, "Error writing output distances\n");
            exit(101);
        }
        buf.clear();
    }
    void line(const std::string &a, const std::string &b, float d0, const float *d1) {
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
        if (buf.size() > (1 << 22) - 4096) flush();
    }
};

// Formats rows [r0, r1) of a dense result on `threads` host threads (row_fn appends the text of one row) and
// writes the blocks in row order.  The reference's Display is serial; --threads is reused for this tail.
void write_rows(Out &out, uint64_t r0, uint64_t r1, int threads, const std::function<void(uint64_t, std::string &)> &row_fn) {
    if (threads <= 1 || r1 - r0 < 64) {
        for (uint64_t i = r0; i < r1; ++i) {
            row_fn(i, out.buf);
            if (out.buf.size() > (1 << 22)) out.flush();
        }
        return;
    }
    const uint64_t block = 32;
    const uint64_t n_blocks = (r1 - r0 + block - 1) / block;
    std::vector<std::string> text(n_blocks);
    std::atomic<uint64_t> next{0};
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&] {
            for (;;) {
                uint64_t b = next.fetch_add(1);
                if (b >= n_blocks) break;
                for (uint64_t i = r0 + b * block; i < std::min(r1, r0 + (b + 1) * block); ++i) row_fn(i, text[b]);
            }
        });
    for (auto &th : pool) th.join();
    out.flush();
    for (auto &t : text) {
        out.buf.swap(t);
        out.flush();
    }
}

inline void append_line(std::string &buf, const std::string &a, const std::string &b, float d0, const float *d1) {
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

struct Pinned {
    skd_ctx *ctx;
    void *p = nullptr;
    Pinned(const Gpu &g, size_t bytes) : ctx(g.ctx) { g.check(skd_host_alloc(ctx, bytes, &p)); }
    ~Pinned() { skd_host_free(ctx, p); }
};

// ---- --gpus N: one host thread + context per device, joined in a peer group of the library (skd_comm_*).
// Every device reads only its 1/N of the .skd (skd_comm_db_load_skd: the other slices arrive over NVLink).
// Sparse self mode: every device evaluates an equal-pair block of upper-triangle rows, the partial lists are merged
// out of peer memory and the final rows gathered on device 0 (skd_comm_self_dists_knn).  Sparse cross mode: the
// references are sharded over the devices and, per device, streamed through HBM in shards of --ref-shard-mb (the next
// shard loads while the previous one is compared); the per-shard lists are folded, merged across the devices and
// gathered on device 0 (skd_comm_knn_merge).  Dense outputs: see dense_multi below.
struct KnnOut {
    std::vector<uint64_t> idx;
    std::vector<float> d0, d1;
    uint64_t rows = 0, knn = 0;
};

// Peer group of `world` worker threads: creates the contexts and windows, exchanges the handles, runs
// body(rank, ctx, comm) on every thread and tears everything down once all threads are done.
struct Rendezvous {
    std::mutex m;
    std::condition_variable cv;
    int arrived = 0, generation = 0;
    bool broken = false;
    bool wait(int world) {  // false if a rank failed before everybody arrived
        std::unique_lock<std::mutex> lk(m);
        if (broken) return false;
        const int gen = generation;
        if (++arrived == world) {
            arrived = 0;
            ++generation;
            cv.notify_all();
            return true;
        }
        cv.wait(lk, [&] { return generation != gen || broken; });
        return !broken;
    }
    void abort() {
        std::lock_guard<std::mutex> lk(m);
        broken = true;
        cv.notify_all();
    }
};

void run_peer_group(int world, int first_device, size_t window_bytes,
                    const std::function<void(int, skd_ctx *, skd_comm *)> &body) {
    std::vector<skd_ctx *> ctxs(world, nullptr);
    std::vector<skd_comm *> comms(world, nullptr);
    std::vector<unsigned char> handles((size_t)world * SKD_COMM_HANDLE_BYTES);
    std::vector<std::string> errs(world);
    Rendezvous gate;
    auto worker = [&](int d) {
        try {
            // more ranks than devices (--gpus N on a smaller box): the ranks share the devices round robin
            const int n_dev = std::max(1, skd_device_count());
            const int dev = world == 1 ? first_device : d % n_dev;
            if (skd_ctx_create(dev, &ctxs[d]) != SKD_OK)
                throw std::runtime_error("device " + std::to_string(dev) + ": " + skd_last_error(nullptr));
            auto chk = [&](skd_status st, const char *what) {
                if (st != SKD_OK) throw std::runtime_error(std::string(what) + ": " + skd_last_error(ctxs[d]));
            };
            chk(skd_comm_create(ctxs[d], d, world, window_bytes, &comms[d]), "peer window");
            chk(skd_comm_export(comms[d], handles.data() + (size_t)d * SKD_COMM_HANDLE_BYTES), "peer handle");
            if (!gate.wait(world)) return;
            chk(skd_comm_connect(comms[d], handles.data()), "peer connect");
            if (!gate.wait(world)) return;
            body(d, ctxs[d], comms[d]);
        } catch (const std::exception &e) {
            errs[d] = e.what();
            gate.abort();
        }
    };
    std::vector<std::thread> pool;
    for (int d = 0; d < world; ++d) pool.emplace_back(worker, d);
    for (auto &t : pool) t.join();
    for (int d = 0; d < world; ++d) {  // only now: a failing rank must not pull its window from under its peers
        if (comms[d]) skd_comm_destroy(comms[d]);
        if (ctxs[d]) skd_ctx_destroy(ctxs[d]);
    }
    for (auto &e : errs)
        if (!e.empty()) throw std::runtime_error(e);
}

KnnOut multi_gpu_knn(const DistArgs &a, const MultiSketch &refs, const std::string &ref_prefix,
                     const std::vector<uint64_t> *subset_idx, const MultiSketch *queries, const std::string &q_prefix,
                     skd_dist_type dt, uint64_t nn, const double *ref_c, const double *query_c) {
    const int world = a.gpus;
    const uint64_t n = refs.number_samples_loaded();
    const bool cross = queries != nullptr;
    const uint64_t rows = cross ? queries->number_samples_loaded() : n;
    const bool with_aux = dt.kind == SKD_CORE_ACC;
    const std::string ref_skd = ref_prefix + ".skd", q_skd = q_prefix + ".skd";
    KnnOut out;
    out.rows = rows;
    out.knn = nn;
    out.idx.resize(rows * nn);
    out.d0.resize(rows * nn);
    if (with_aux) out.d1.resize(rows * nn);
    size_t window = skd_comm_window_bytes_knn(rows, nn, world);
    window = std::max(window, skd_comm_window_bytes_db(cross ? rows : n, refs.sketchsize64, (uint32_t)refs.kmer_lengths.size(), world));
    auto file_indices = [&](uint64_t lo, uint64_t hi) {
        std::vector<uint64_t> v(hi - lo);
        // no subset: file POSITION, like the reference's read_sketch_data (multisketch.rs:167-184); after `delete`
        // Sketch::index keeps the old numbering while the .skd is rewritten, so index != position
        for (uint64_t s = lo; s < hi; ++s) v[s - lo] = subset_idx ? (*subset_idx)[s] : s;
        return v;
    };
    run_peer_group(world, a.gpu, window, [&](int d, skd_ctx *ctx, skd_comm *comm) {
        auto chk = [&](skd_status st, const char *what) {
            if (st != SKD_OK) throw std::runtime_error(std::string(what) + ": " + skd_last_error(ctx));
        };
        struct Db {
            skd_db *p = nullptr;
            ~Db() {
                if (p) skd_db_free(p);
            }
        };
        uint64_t *oi = d == 0 ? out.idx.data() : nullptr;
        float *o0 = d == 0 ? out.d0.data() : nullptr, *o1 = d == 0 && with_aux ? out.d1.data() : nullptr;
        if (!cross) {
            Db db;
            chk(skd_comm_db_load_skd(comm, ref_skd.c_str(), subset_idx ? subset_idx->data() : nullptr, n, refs.sample_stride,
                                     refs.kmer_stride, refs.sketchsize64, refs.sketch_size, refs.kmer_lengths.data(),
                                     (uint32_t)refs.kmer_lengths.size(), &db.p),
                "loading references");
            chk(skd_comm_self_dists_knn(comm, db.p, nn, dt, ref_c, a.completeness_cutoff, 0, oi, o0, o1, SKD_HOST), "self kNN");
            return;
        }
        Db qdb;
        chk(skd_comm_db_load_skd(comm, q_skd.c_str(), nullptr, rows, queries->sample_stride, queries->kmer_stride,
                                 queries->sketchsize64, queries->sketch_size, queries->kmer_lengths.data(),
                                 (uint32_t)queries->kmer_lengths.size(), &qdb.p),
            "loading queries");
        // references of this device, streamed through HBM in shards
        uint64_t lo, hi;
        skd_slice_rows(n, d, world, 64, &lo, &hi);
        auto load_refs = [&](skd_ctx *c, uint64_t s0, uint64_t s1, skd_db **o) {
            std::vector<uint64_t> fi = file_indices(s0, s1);
            return skd_db_load_skd(c, ref_skd.c_str(), fi.data(), s1 - s0, refs.sample_stride, refs.kmer_stride,
                                   refs.sketchsize64, refs.sketch_size, refs.kmer_lengths.data(),
                                   (uint32_t)refs.kmer_lengths.size(), o) == SKD_OK;
        };
        uint64_t per_shard = std::max<uint64_t>(64, ((uint64_t)a.ref_shard_mb << 20) / (refs.sample_stride * 8) / 64 * 64);
        const uint64_t max_parts = 32;  // lists folded per device
        per_shard = std::max(per_shard, ((hi - lo + max_parts - 1) / max_parts + 63) / 64 * 64);
        const uint64_t ns = hi > lo ? (hi - lo + per_shard - 1) / per_shard : 0;
        const uint64_t np = std::max<uint64_t>(1, ns);
        std::vector<uint64_t> keys(np * rows * nn, ~0ull);  // a device without references contributes only padding
        std::vector<float> aux(with_aux ? np * rows * nn : 0, 0.f);
        skd_ctx *lctx = nullptr;
        if (ns > 1 && skd_ctx_create(skd_ctx_device(ctx), &lctx) != SKD_OK)
            throw std::runtime_error(std::string("loader context: ") + skd_last_error(nullptr));
        struct LoaderCtx {
            skd_ctx *&c;
            ~LoaderCtx() {
                if (c) skd_ctx_destroy(c);
            }
        } loader_guard{lctx};
        Db next;
        if (ns > 0 && !load_refs(ctx, lo, std::min(hi, lo + per_shard), &next.p)) chk(SKD_ERR_IO, "loading references");
        for (uint64_t i = 0; i < ns; ++i) {
            Db cur;
            std::swap(cur.p, next.p);
            const uint64_t s0 = lo + i * per_shard;
            std::thread loader;
            std::string load_err;
            if (i + 1 < ns)
                loader = std::thread([&, i] {
                    const uint64_t n0 = lo + (i + 1) * per_shard;
                    if (!load_refs(lctx, n0, std::min(hi, n0 + per_shard), &next.p))
                        load_err = std::string("loading references: ") + skd_last_error(lctx);
                });
            const skd_status st = skd_cross_knn_partial(ctx, cur.p, qdb.p, nn, dt, ref_c ? ref_c + s0 : nullptr, query_c,
                                                        a.completeness_cutoff, s0, keys.data() + i * rows * nn,
                                                        with_aux ? aux.data() + i * rows * nn : nullptr, SKD_HOST);
            if (loader.joinable()) loader.join();
            chk(st, "cross kNN");
            if (!load_err.empty()) throw std::runtime_error(load_err);
            if (next.p) chk(skd_db_adopt(ctx, next.p), "adopting a reference shard");
        }
        chk(skd_comm_knn_merge(comm, keys.data(), with_aux ? aux.data() : nullptr, (uint32_t)np, SKD_HOST, rows, nn, dt, 0, oi,
                               o0, o1, SKD_HOST),
            "merging the partial lists");
    });
    return out;
}

// ---- dense outputs on several devices.  The rows are cut into chunks of about 128 MB of results; chunk c is computed by
// device c % N into one of that device's two pinned buffers (the kernels write them directly), and the main thread
// formats and writes the chunks in row order while the devices work ahead: disjoint slices of the reference's
// DistanceMatrix, no exchange between the devices (src/distances/mod.rs:58-130, 227-297).
void dense_multi(const DistArgs &a, Out &out, const MultiSketch &refs, const std::string &ref_prefix,
                 const std::vector<uint64_t> *subset_idx, const MultiSketch *queries, const std::string &q_prefix,
                 skd_dist_type dt, const double *ref_c, const double *query_c) {
    const int world = a.gpus;
    const uint64_t n = refs.number_samples_loaded();
    const bool cross = queries != nullptr;
    const uint64_t nq = cross ? queries->number_samples_loaded() : 0;
    const int ncols = dt.kind == SKD_CORE_ACC ? 2 : 1;
    const size_t kChunkBytes = (size_t)128 << 20;
    const uint64_t width = cross ? nq : n;
    const uint64_t rows_per = std::max<uint64_t>(64, kChunkBytes / std::max<uint64_t>(1, width * ncols * 4) / 64 * 64);
    const uint64_t last_row = cross ? n : (n >= 2 ? n - 1 : 0);  // self: the last row has no pairs
    struct Chunk {
        uint64_t r0, r1;
    };
    std::vector<Chunk> chunks;
    for (uint64_t r0 = 0; r0 < last_row; r0 += rows_per) chunks.push_back({r0, std::min(cross ? n : n, r0 + rows_per)});
    if (chunks.empty()) return;
    auto elems = [&](const Chunk &c) {
        return cross ? (c.r1 - c.r0) * nq : skd_self_pairs_in_rows(n, c.r0, c.r1);
    };
    size_t buf_elems = 0;
    for (auto &c : chunks) buf_elems = std::max<size_t>(buf_elems, elems(c));
    // slot s of device d holds chunk `holds`; state 0 = free, 1 = full
    struct Slot {
        void *p = nullptr;
        int state = 0;
        size_t chunk = 0;
    };
    std::vector<std::array<Slot, 2>> slots(world);
    std::mutex m;
    std::condition_variable cv;
    bool failed = false;
    const std::string ref_skd = ref_prefix + ".skd", q_skd = q_prefix + ".skd";
    size_t window = skd_comm_window_bytes_db(n, refs.sketchsize64, (uint32_t)refs.kmer_lengths.size(), world);
    if (cross) window = std::max(window, skd_comm_window_bytes_db(nq, refs.sketchsize64, (uint32_t)refs.kmer_lengths.size(), world));
    std::exception_ptr writer_err;
    std::thread writer([&] {
        try {
            for (size_t c = 0; c < chunks.size(); ++c) {
                const int d = (int)(c % world);
                Slot *s = nullptr;
                {
                    std::unique_lock<std::mutex> lk(m);
                    cv.wait(lk, [&] {
                        for (auto &sl : slots[d])
                            if (sl.state == 1 && sl.chunk == c) s = &sl;
                        return s != nullptr || failed;
                    });
                    if (failed) return;
                }
                const Chunk &ch = chunks[c];
                const float *base = (const float *)s->p;
                if (cross) {
                    write_rows(out, ch.r0, ch.r1, (int)a.threads, [&](uint64_t i, std::string &txt) {
                        const float *dd = base + (i - ch.r0) * nq * ncols;
                        const std::string &rn = refs.sketch_name(i);
                        for (uint64_t j = 0; j < nq; ++j, dd += ncols)
                            append_line(txt, rn, queries->sketch_name(j), dd[0], ncols == 2 ? dd + 1 : nullptr);
                    });
                } else {
                    write_rows(out, ch.r0, std::min(ch.r1, n - 1), (int)a.threads, [&](uint64_t i, std::string &txt) {
                        const float *dd = base + skd_self_pairs_in_rows(n, ch.r0, i) * ncols;
                        const std::string &rn = refs.sketch_name(i);
                        for (uint64_t j = i + 1; j < n; ++j, dd += ncols)
                            append_line(txt, rn, refs.sketch_name(j), dd[0], ncols == 2 ? dd + 1 : nullptr);
                    });
                }
                {
                    std::lock_guard<std::mutex> lk(m);
                    s->state = 0;
                }
                cv.notify_all();
            }
        } catch (...) {
            writer_err = std::current_exception();
            std::lock_guard<std::mutex> lk(m);
            failed = true;
            cv.notify_all();
        }
    });
    std::exception_ptr group_err;
    try {
        run_peer_group(world, a.gpu, window, [&](int d, skd_ctx *ctx, skd_comm *comm) {
            auto chk = [&](skd_status st, const char *what) {
                if (st != SKD_OK) throw std::runtime_error(std::string(what) + ": " + skd_last_error(ctx));
            };
            skd_db *rdb = nullptr, *qdb = nullptr;
            try {
                for (auto &sl : slots[d]) chk(skd_host_alloc(ctx, buf_elems * ncols * 4 + 16, &sl.p), "result buffer");
                chk(skd_comm_db_load_skd(comm, ref_skd.c_str(), subset_idx ? subset_idx->data() : nullptr, n, refs.sample_stride,
                                         refs.kmer_stride, refs.sketchsize64, refs.sketch_size, refs.kmer_lengths.data(),
                                         (uint32_t)refs.kmer_lengths.size(), &rdb),
                    "loading references");
                if (cross)
                    chk(skd_comm_db_load_skd(comm, q_skd.c_str(), nullptr, nq, queries->sample_stride, queries->kmer_stride,
                                             queries->sketchsize64, queries->sketch_size, queries->kmer_lengths.data(),
                                             (uint32_t)queries->kmer_lengths.size(), &qdb),
                        "loading queries");
                for (size_t c = (size_t)d; c < chunks.size(); c += world) {
                    Slot *s = nullptr;
                    {
                        std::unique_lock<std::mutex> lk(m);
                        cv.wait(lk, [&] {
                            for (auto &sl : slots[d])
                                if (sl.state == 0) s = &sl;
                            return s != nullptr || failed;
                        });
                        if (failed) break;
                    }
                    const Chunk &ch = chunks[c];
                    if (cross)
                        chk(skd_cross_dists_all(ctx, rdb, qdb, dt, ref_c, query_c, a.completeness_cutoff, ch.r0, ch.r1,
                                                (float *)s->p, SKD_HOST),
                            "dense distances");
                    else
                        chk(skd_self_dists_all(ctx, rdb, dt, ref_c, a.completeness_cutoff, ch.r0, ch.r1, (float *)s->p, SKD_HOST),
                            "dense distances");
                    {
                        std::lock_guard<std::mutex> lk(m);
                        s->chunk = c;
                        s->state = 1;
                    }
                    cv.notify_all();
                }
                {   // the buffers are read by the writer until the last chunk of this device is out
                    std::unique_lock<std::mutex> lk(m);
                    cv.wait(lk, [&] { return failed || (slots[d][0].state == 0 && slots[d][1].state == 0); });
                }
            } catch (...) {
                {
                    std::lock_guard<std::mutex> lk(m);
                    failed = true;
                }
                cv.notify_all();
                if (rdb) skd_db_free(rdb);
                if (qdb) skd_db_free(qdb);
                for (auto &sl : slots[d])
                    if (sl.p) skd_host_free(ctx, sl.p);
                throw;
            }
            if (rdb) skd_db_free(rdb);
            if (qdb) skd_db_free(qdb);
            for (auto &sl : slots[d])
                if (sl.p) skd_host_free(ctx, sl.p);
        });
    } catch (...) {
        group_err = std::current_exception();
        std::lock_guard<std::mutex> lk(m);
        failed = true;
        cv.notify_all();
    }
    writer.join();
    if (group_err) std::rethrow_exception(group_err);
    if (writer_err) std::rethrow_exception(writer_err);
}

int run(const DistArgs &a) {
    g_level = a.quiet ? 0 : (a.verbose ? 2 : 1);
    log_info("Using " + std::to_string(a.threads) + " threads (host); distances run on CUDA device " + std::to_string(a.gpu));
    Gpu gpu;  // declared first: destroyed after the databases that belong to its context
    // ---- host phase: everything that can fail without the GPU (same messages as the reference)
    const std::string ref_name = strip_sketch_extension(a.ref_db);
    Loaded refs;
    try {
        refs.meta = MultiSketch::load_metadata(ref_name);
    } catch (const std::exception &) {
        throw Panic("Could not read sketch metadata from " + a.ref_db + ".skm");
    }
    std::vector<uint64_t> subset_idx;
    if (!a.subset.empty()) {
        try {
            std::vector<std::string> names = read_subset_names(a.subset);
            log_info("Read " + std::to_string(names.size()) + " names to subset");
            subset_idx = refs.meta.select_subset(names);
        } catch (const std::exception &e) {
            throw Panic(e.what());
        }
    }
    const uint64_t n = refs.meta.number_samples_loaded();
    std::vector<std::string> warnings;
    std::vector<double> ref_c, query_c;
    const bool has_ref_c = !a.ref_completeness_file.empty();
    if (has_ref_c) {
        ref_c = read_completeness_file(a.ref_completeness_file, refs.meta, &warnings);
        for (auto &w : warnings) log_warn(w);
        warnings.clear();
    }
    if (a.kmer >= 0 && !refs.meta.get_k_idx((uint64_t)a.kmer))  // set_k: src/distances/mod.rs:25-37
        throw Panic("Error setting k size: K-mer size " + std::to_string(a.kmer) + " not found in file");
    Loaded queries;
    std::string qname;
    const bool has_query_c = a.has_query && !a.query_completeness_file.empty();
    if (a.has_query) {
        qname = strip_sketch_extension(a.query_db);
        try {
            queries.meta = MultiSketch::load_metadata(qname);
        } catch (const std::exception &) {
            throw Panic("Could not read sketch metadata from " + a.query_db + ".skm");
        }
        if (has_query_c) {
            query_c = read_completeness_file(a.query_completeness_file, queries.meta, &warnings);
            for (auto &w : warnings) log_warn(w);
        }
    }

    // ---- several devices: sparse outputs only (for dense outputs the text tail, not the GPU, is the bottleneck).
    // The same driver runs a single device when a cross --knn run has more reference data than one shard: the
    // references are then streamed through HBM shard by shard, each loaded while the previous one is compared.
    const bool sharded_refs = a.has_query && a.knn >= 0 &&
                              n * refs.meta.sample_stride * 8 > ((uint64_t)a.ref_shard_mb << 20);
    if (a.gpus > 1 && skd_device_count() > 0 && skd_device_count() < a.gpus)
        log_warn("--gpus " + std::to_string(a.gpus) + " but only " + std::to_string(skd_device_count()) +
                 " device(s) visible: the ranks share them");
    if ((a.gpus > 1 || sharded_refs) && a.knn >= 0) {
        Out out(a);
        skd_dist_type dt;
        dt.kind = a.kmer < 0 ? SKD_CORE_ACC : (a.ani ? SKD_ANI : SKD_JACCARD);
        dt.k_idx = a.kmer < 0 ? 0 : (int32_t)*refs.meta.get_k_idx((uint64_t)a.kmer);
        if (dt.kind == SKD_CORE_ACC && refs.meta.kmer_lengths.size() < 2)
            throw Panic("Need at least two k-mer lengths to calculate core/accessory distances");
        uint64_t nn = (uint64_t)a.knn;
        if (!a.has_query) {
            if (n < 2) throw Panic("attempt to subtract with overflow: fewer than two samples loaded");
            if (nn >= n) {
                log_warn("knn=" + std::to_string(nn) + " is higher than number of samples=" + std::to_string(n));
                nn = n - 1;
            }
        } else {
            if (n == 0) throw Panic("Reference database has no loaded samples");
            if (queries.meta.number_samples_loaded() == 0) throw Panic("Query database has no loaded samples");
            if (!refs.meta.is_compatible_with(queries.meta))
                throw std::runtime_error("reference and query databases differ in sketch size or k-mer lengths");
            if (nn > n) {
                log_warn("knn=" + std::to_string(nn) + " is higher than number of reference samples=" + std::to_string(n));
                nn = n;
            }
        }
        if (nn == 0) throw std::runtime_error("knn must be at least 1");
        log_info("Calculating sparse distances with " + std::to_string(nn) + " nearest neighbours on " +
                 std::to_string(a.gpus) + " device(s)" + (sharded_refs ? ", references in shards" : ""));
        KnnOut r = multi_gpu_knn(a, refs.meta, ref_name, a.subset.empty() ? nullptr : &subset_idx,
                                 a.has_query ? &queries.meta : nullptr, qname, dt, nn, has_ref_c ? ref_c.data() : nullptr,
                                 has_query_c ? query_c.data() : nullptr);
        const MultiSketch &row_meta = a.has_query ? queries.meta : refs.meta;
        const bool two = dt.kind == SKD_CORE_ACC;
        for (uint64_t row = 0; row < r.rows; ++row) {
            const std::string &rn = row_meta.sketch_name(row);
            for (uint64_t t = 0; t < nn; ++t) {
                const std::string &cn = refs.meta.sketch_name(r.idx[row * nn + t]);
                if (two)
                    out.line(rn, cn, r.d0[row * nn + t], &r.d1[row * nn + t]);
                else if (r.d0[row * nn + t] < 1.0f || cn != rn)
                    out.line(rn, cn, r.d0[row * nn + t], nullptr);
            }
        }
        out.flush();
        return 0;
    }
    if (a.gpus > 1) {  // dense outputs: disjoint row chunks per device, written in row order
        Out out(a);
        skd_dist_type dt;
        dt.kind = a.kmer < 0 ? SKD_CORE_ACC : (a.ani ? SKD_ANI : SKD_JACCARD);
        dt.k_idx = a.kmer < 0 ? 0 : (int32_t)*refs.meta.get_k_idx((uint64_t)a.kmer);
        if (dt.kind == SKD_CORE_ACC) {
            if (refs.meta.kmer_lengths.size() < 2)
                throw Panic("Need at least two k-mer lengths to calculate core/accessory distances");
            log_info("Distances: core/accessory regression");
        } else {
            log_info(std::string(dt.kind == SKD_ANI ? "Distances: ANI at k=" : "Distances: Jaccard distances at k=") +
                     std::to_string(a.kmer));
        }
        if (a.has_query && !refs.meta.is_compatible_with(queries.meta))
            throw std::runtime_error("reference and query databases differ in sketch size or k-mer lengths");
        log_info(std::string(a.has_query ? "Calculating all ref vs query distances" : "Calculating all ref vs ref distances") +
                 " on " + std::to_string(a.gpus) + " devices");
        const bool any = a.has_query ? (n > 0 && queries.meta.number_samples_loaded() > 0) : n >= 2;
        if (any) {
            log_info("Writing out in long matrix form");
            dense_multi(a, out, refs.meta, ref_name, a.subset.empty() ? nullptr : &subset_idx,
                        a.has_query ? &queries.meta : nullptr, qname, dt, has_ref_c ? ref_c.data() : nullptr,
                        has_query_c ? query_c.data() : nullptr);
        }
        out.flush();
        return 0;
    }

    // ---- device phase
    Out out(a);
    if (skd_ctx_create(a.gpu, &gpu.ctx) != SKD_OK) throw std::runtime_error(skd_last_error(nullptr));
    log_info("Loading sketch data from " + ref_name + ".skd");
    load_db(gpu, refs, ref_name, a.subset.empty() ? nullptr : &subset_idx);
    skd_dist_type dt;
    {
        skd_status st = skd_set_k(gpu.ctx, refs.db, a.kmer, a.ani ? 1 : 0, &dt);
        if (st != SKD_OK) throw Panic(std::string("Error setting k size: ") + skd_last_error(gpu.ctx));
        if (dt.kind == SKD_CORE_ACC)
            log_info("Distances: core/accessory regression");
        else
            log_info(std::string(dt.kind == SKD_ANI ? "Distances: ANI at k=" : "Distances: Jaccard distances at k=") +
                     std::to_string(a.kmer));
    }
    const int ncols = dt.kind == SKD_CORE_ACC ? 2 : 1;
    if (a.has_query) {
        log_info("Loading query sketch data from " + qname + ".skd");
        load_db(gpu, queries, qname, nullptr);
    }

    const size_t kChunkBytes = (size_t)128 << 20;  // dense results are produced and written in row chunks
    if (!a.has_query) {
        if (a.knn < 0) {
            log_info("Calculating all ref vs ref distances");
            if (n >= 2) {
                uint64_t rows_per = std::max<uint64_t>(64, kChunkBytes / std::max<uint64_t>(1, n * ncols * 4) / 64 * 64);
                Pinned buf(gpu, std::min<size_t>(skd_self_pairs_in_rows(n, 0, n), (size_t)rows_per * n) * ncols * 4 + 16);
                log_info("Writing out in long matrix form");
                for (uint64_t r0 = 0; r0 < n - 1; r0 += rows_per) {
                    uint64_t r1 = std::min(n, r0 + rows_per);
                    gpu.check(skd_self_dists_all(gpu.ctx, refs.db, dt, has_ref_c ? ref_c.data() : nullptr,
                                                 a.completeness_cutoff, r0, r1, (float *)buf.p, SKD_HOST));
                    const float *base = (const float *)buf.p;
                    write_rows(out, r0, std::min(r1, n - 1), (int)a.threads, [&](uint64_t i, std::string &txt) {
                        const float *d = base + skd_self_pairs_in_rows(n, r0, i) * ncols;
                        const std::string &rn = refs.meta.sketch_name(i);
                        for (uint64_t j = i + 1; j < n; ++j, d += ncols)
                            append_line(txt, rn, refs.meta.sketch_name(j), d[0], ncols == 2 ? d + 1 : nullptr);
                    });
                }
            } else {
                gpu.check(skd_self_dists_all(gpu.ctx, refs.db, dt, nullptr, a.completeness_cutoff, 0, n, nullptr, SKD_HOST));
            }
        } else {
            if (n < 2) throw Panic("attempt to subtract with overflow: fewer than two samples loaded");
            uint64_t nn = (uint64_t)a.knn;
            if (nn >= n) {  // a genome cannot be its own neighbour: src/lib.rs:379-382
                log_warn("knn=" + std::to_string(nn) + " is higher than number of samples=" + std::to_string(n));
                nn = n - 1;
            }
            log_info("Calculating sparse ref vs ref distances with " + std::to_string(nn) + " nearest neighbours");
            std::vector<uint64_t> idx(n * nn);
            std::vector<float> d0(n * nn), d1(ncols == 2 ? n * nn : 0);
            gpu.check(skd_self_dists_knn(gpu.ctx, refs.db, nn, dt, has_ref_c ? ref_c.data() : nullptr,
                                         a.completeness_cutoff, 0, n, idx.data(), d0.data(),
                                         ncols == 2 ? d1.data() : nullptr, SKD_HOST));
            log_info("Writing out in sparse matrix form");
            for (uint64_t r = 0; r < n; ++r) {
                const std::string &rn = refs.meta.sketch_name(r);
                for (uint64_t t = 0; t < nn; ++t) {
                    const std::string &cn = refs.meta.sketch_name(idx[r * nn + t]);
                    if (ncols == 2)
                        out.line(rn, cn, d0[r * nn + t], &d1[r * nn + t]);
                    else if (d0[r * nn + t] < 1.0f || cn != rn)  // padding rule: distance_matrix.rs:376-381
                        out.line(rn, cn, d0[r * nn + t], nullptr);
                }
            }
        }
    } else {
        const uint64_t nq = queries.meta.number_samples_loaded();
        const double *rc = has_ref_c ? ref_c.data() : nullptr, *qc = has_query_c ? query_c.data() : nullptr;
        if (a.knn >= 0) {
            uint64_t nn = (uint64_t)a.knn;
            if (nn > n) {  // src/lib.rs:411-414
                log_warn("knn=" + std::to_string(nn) + " is higher than number of reference samples=" + std::to_string(n));
                nn = n;
            }
            log_info("Calculating sparse ref vs query distances with " + std::to_string(nn) + " nearest neighbours");
            uint64_t used = nn;
            std::vector<uint64_t> idx(nq * std::max<uint64_t>(nn, 1));
            std::vector<float> d0(idx.size()), d1(ncols == 2 ? idx.size() : 0);
            gpu.check(skd_cross_dists_knn(gpu.ctx, refs.db, queries.db, nn, dt, rc, qc, a.completeness_cutoff, 0, nq,
                                          idx.data(), d0.data(), ncols == 2 ? d1.data() : nullptr, SKD_HOST, &used));
            log_info("Writing out in sparse matrix form");
            for (uint64_t r = 0; r < nq; ++r) {
                const std::string &rn = queries.meta.sketch_name(r);  // rows are queries: distance_matrix.rs:364
                for (uint64_t t = 0; t < used; ++t) {
                    const std::string &cn = refs.meta.sketch_name(idx[r * used + t]);
                    if (ncols == 2)
                        out.line(rn, cn, d0[r * used + t], &d1[r * used + t]);
                    else if (d0[r * used + t] < 1.0f || cn != rn)
                        out.line(rn, cn, d0[r * used + t], nullptr);
                }
            }
        } else {
            log_info("Calculating all ref vs query distances");
            if (n > 0 && nq > 0) {
                uint64_t rows_per = std::max<uint64_t>(64, kChunkBytes / std::max<uint64_t>(1, nq * ncols * 4) / 64 * 64);
                Pinned buf(gpu, (size_t)std::min(rows_per, n) * nq * ncols * 4 + 16);
                log_info("Writing out in long matrix form");
                for (uint64_t r0 = 0; r0 < n; r0 += rows_per) {
                    uint64_t r1 = std::min(n, r0 + rows_per);
                    gpu.check(skd_cross_dists_all(gpu.ctx, refs.db, queries.db, dt, rc, qc, a.completeness_cutoff, r0, r1,
                                                  (float *)buf.p, SKD_HOST));
                    const float *base = (const float *)buf.p;
                    write_rows(out, r0, r1, (int)a.threads, [&](uint64_t i, std::string &txt) {
                        const float *d = base + (i - r0) * nq * ncols;
                        const std::string &rn = refs.meta.sketch_name(i);
                        for (uint64_t j = 0; j < nq; ++j, d += ncols)
                            append_line(txt, rn, queries.meta.sketch_name(j), d[0], ncols == 2 ? d + 1 : nullptr);
                    });
                }
            } else {
                gpu.check(skd_cross_dists_all(gpu.ctx, refs.db, queries.db, dt, rc, qc, a.completeness_cutoff, 0, n,
                                              nullptr, SKD_HOST));
            }
        }
    }
    out.flush();
    return 0;
}

// get_input_list: src/io.rs:182-224 (file list: name [file...] per line) / read_input_fastas: src/io.rs:20-38
InputList get_input_list(const std::string &file_list, const std::vector<std::string> &seq_files) {
    InputList inputs;
    if (!file_list.empty()) {
        std::ifstream in(file_list);
        if (!in) throw Panic("Unable to open file_list " + file_list);
        std::string line;
        while (std::getline(in, line)) {
            std::vector<std::string> f;
            size_t pos = 0;
            while (pos < line.size()) {
                while (pos < line.size() && isspace((unsigned char)line[pos])) ++pos;
                size_t e = pos;
                while (e < line.size() && !isspace((unsigned char)line[e])) ++e;
                if (e > pos) f.push_back(line.substr(pos, e - pos));
                pos = e;
            }
            if (f.empty()) throw Panic("Unable to parse line in file_list");
            if (f.size() == 1) inputs.push_back({f[0], {f[0]}});
            else inputs.push_back({f[0], std::vector<std::string>(f.begin() + 1, f.end())});
        }
    } else {
        // read_input_fastas (src/io.rs:20-38): the file name without directories when the extension is one of
        // fa, fasta, fastq, fq (optionally .gz); any other argument names its sample by the whole string
        auto known_ext = [](const std::string &f) {
            for (const char *e : {".fa", ".fasta", ".fa.gz", ".fasta.gz", ".fastq", ".fastq.gz", ".fq", ".fq.gz"}) {
                const size_t n = strlen(e);
                if (f.size() > n && f.compare(f.size() - n, n, e) == 0) return true;
            }
            return false;
        };
        for (auto &file : seq_files) {
            size_t slash = file.rfind('/');
            const bool base = known_ext(file) && slash != std::string::npos && slash > 0 && slash + 1 < file.size();
            inputs.push_back({base ? file.substr(slash + 1) : file, {file}});
        }
    }
    return inputs;
}

}  // namespace

// ---- `sketchlib sketch`: flags of src/cli.rs:128-182 that apply to DNA assemblies
int run_sketch(int argc, char **argv) {
    std::vector<std::string> seq_files;
    std::string file_list, output;
    std::vector<uint64_t> k_vals, k_seq;
    uint64_t sketch_size = 1000;  // DEFAULT_SKETCHSIZE, src/cli.rs:17-19
    bool single_strand = false, verbose = false, quiet = false;
    long long threads = 1;
    auto split_u64 = [](const std::string &v, const char *what) {
        std::vector<uint64_t> out;
        size_t pos = 0;
        while (pos <= v.size()) {
            size_t c = v.find(',', pos);
            if (c == std::string::npos) c = v.size();
            out.push_back((uint64_t)parse_uint(v.substr(pos, c - pos), what));
            pos = c + 1;
        }
        return out;
    };
    bool saw = false;
    for (int i = 1; i < argc; ++i) {
        std::string t = argv[i];
        auto value = [&](const char *name) -> std::string {
            size_t eq = t.find('=');
            if (t.rfind("--", 0) == 0 && eq != std::string::npos) return t.substr(eq + 1);
            if (t.rfind("--", 0) != 0 && t.size() > 2) return t[2] == '=' ? t.substr(3) : t.substr(2);
            if (i + 1 >= argc) throw UsageError(std::string("a value is required for '") + name + "' but none was supplied");
            return argv[++i];
        };
        auto is = [&](const char *name) { return t == name || t.rfind(std::string(name) + "=", 0) == 0; };
        if (t == "-v" || t == "--verbose") verbose = true;
        else if (t == "--quiet") quiet = true;
        else if (t == "--single-strand") single_strand = true;
        else if (is("--k-vals")) k_vals = split_u64(value("--k-vals <K_VALS>"), "--k-vals <K_VALS>");
        else if (is("--k-seq")) k_seq = split_u64(value("--k-seq <K_SEQ>"), "--k-seq <K_SEQ>");
        else if (is("--sketch-size")) sketch_size = (uint64_t)parse_uint(value("--sketch-size"), "--sketch-size <SKETCH_SIZE>");
        else if (is("--threads")) threads = parse_uint(value("--threads"), "--threads <THREADS>");
        else if (is("--seq-type")) {
            if (value("--seq-type") != "dna") throw UsageError("this build sketches DNA only (--seq-type dna)");
        } else if (is("--min-count") || is("--min-qual") || is("--level")) (void)value("option");
        else if (t == "--concat-fasta") throw UsageError("--concat-fasta currently only supported with --seq-type aa");
        else if (t.rfind("-k", 0) == 0 && t.rfind("--", 0) != 0) k_vals = split_u64(value("-k <K_VALS>"), "-k <K_VALS>");
        else if (t.rfind("-s", 0) == 0 && t.rfind("--", 0) != 0) sketch_size = (uint64_t)parse_uint(value("-s"), "-s <SKETCH_SIZE>");
        else if (t.rfind("-f", 0) == 0 && t.rfind("--", 0) != 0) file_list = value("-f <FILE_LIST>");
        else if (t.rfind("-o", 0) == 0 && t.rfind("--", 0) != 0) output = value("-o <OUTPUT>");
        else if (t.size() > 1 && t[0] == '-') throw UsageError("unexpected argument '" + t + "' found");
        else if (!saw) saw = true;  // the sub-command word itself
        else seq_files.push_back(t);
    }
    g_level = quiet ? 0 : (verbose ? 2 : 1);
    if (output.empty()) throw UsageError("the following required arguments were not provided:\n  -o <OUTPUT>");
    if (file_list.empty() == seq_files.empty())
        throw UsageError("exactly one of <SEQ_FILES>... and -f <FILE_LIST> is required");
    if (k_vals.empty() == k_seq.empty()) throw UsageError("exactly one of --k-vals and --k-seq is required");
    std::vector<uint64_t> kmers = k_vals;
    if (!k_seq.empty()) {  // parse_kmers: (start..=end).step_by(step), src/io.rs:140-160
        if (k_seq.size() != 3 || k_seq[2] == 0) throw UsageError("--k-seq takes start,end,step");
        for (uint64_t k = k_seq[0]; k <= k_seq[1]; k += k_seq[2]) kmers.push_back(k);
    }
    InputList inputs = get_input_list(file_list, seq_files);
    log_info("Parsed " + std::to_string(inputs.size()) + " samples in input list");
    try {
        sketch_files_dna(output, inputs, kmers, sketch_size, !single_strand, (int)std::max<long long>(1, threads));
    } catch (const std::exception &e) {
        throw Panic(e.what());
    }
    return 0;
}


// ---- `sketchlib inverted build | query | precluster`: flags of src/cli.rs:324-462, control flow of src/lib.rs:484-792.
// build / query are host-only; precluster makes the candidate lists on the host (Inverted::any_shared_bins over the
// .skq bins) and runs the distances + per-row top-k on the GPU (skd_self_dists_knn_candidates).
const char *INVERTED_USAGE =
    "Usage: sketchlib inverted build [OPTIONS] -o <OUTPUT> [SEQ_FILES]...\n"
    "         -f <FILE_LIST>  --write-skq  --species-names <F>  --metadata <F>  -s <SKETCH_SIZE> [1000]  -k <KMER> [21]\n"
    "         --single-strand  --threads <N>\n"
    "       sketchlib inverted query [OPTIONS] <SKI> [SEQ_FILES]...\n"
    "         -f <FILE_LIST>  -o <OUTPUT>  --query-type <match-count|all-bins|any-bins>  --threads <N>\n"
    "       sketchlib inverted precluster [OPTIONS] <SKI> (--skd <DB> | --count)\n"
    "         -o <OUTPUT>  --knn <KNN> [50]  --ani  --threads <N>  --ref-completeness-file <F>\n"
    "         --completeness-cutoff <X> [0.64]  --retain-unmatched <singleton|bruteforce>  --gpu <ID>\n";

struct InvArgs {
    std::string cmd;  // build | query | precluster
    std::vector<std::string> pos;
    std::string file_list, output, species_names, metadata, skd, query_type = "match-count", retain, ref_completeness_file;
    bool has_output = false, write_skq = false, single_strand = false, count = false, ani = false, verbose = false,
         quiet = false;
    uint64_t sketch_size = 1000, kmer = 21, knn = 50;  // DEFAULT_SKETCHSIZE, DEFAULT_KMER, DEFAULT_KNN
    long long threads = 1;
    double completeness_cutoff = 0.64;
    int gpu = 0;
};

InvArgs parse_inverted_args(int argc, char **argv) {
    InvArgs a;
    int words = 0;
    for (int i = 1; i < argc; ++i) {
        std::string t = argv[i];
        auto value = [&](const char *name) -> std::string {
            size_t eq = t.find('=');
            if (t.rfind("--", 0) == 0 && eq != std::string::npos) return t.substr(eq + 1);
            if (t.rfind("--", 0) != 0 && t.size() > 2) return t[2] == '=' ? t.substr(3) : t.substr(2);
            if (i + 1 >= argc) throw UsageError(std::string("a value is required for '") + name + "' but none was supplied");
            return argv[++i];
        };
        auto is = [&](const char *name) { return t == name || t.rfind(std::string(name) + "=", 0) == 0; };
        auto shortopt = [&](char c) { return t.size() >= 2 && t[0] == '-' && t[1] == c; };
        if (t == "-h" || t == "--help") {
            fputs(INVERTED_USAGE, stdout);
            exit(0);
        } else if (t == "-v" || t == "--verbose") a.verbose = true;
        else if (t == "--quiet") a.quiet = true;
        else if (t == "--write-skq") a.write_skq = true;
        else if (t == "--single-strand") a.single_strand = true;
        else if (t == "--count") a.count = true;
        else if (t == "--ani") a.ani = true;
        else if (is("--species-names")) a.species_names = value("--species-names <SPECIES_NAMES>");
        else if (is("--metadata")) a.metadata = value("--metadata <METADATA>");
        else if (is("--sketch-size")) a.sketch_size = (uint64_t)parse_uint(value("--sketch-size"), "--sketch-size <SKETCH_SIZE>");
        else if (is("--kmer-length")) a.kmer = (uint64_t)parse_uint(value("--kmer-length"), "--kmer-length <KMER_LENGTH>");
        else if (is("--skd")) a.skd = value("--skd <SKD>");
        else if (is("--knn")) a.knn = (uint64_t)parse_uint(value("--knn <KNN>"), "--knn <KNN>");
        else if (is("--query-type")) a.query_type = value("--query-type <QUERY_TYPE>");
        else if (is("--retain-unmatched")) a.retain = value("--retain-unmatched <RETAIN_UNMATCHED>");
        else if (is("--ref-completeness-file")) a.ref_completeness_file = value("--ref-completeness-file");
        else if (is("--completeness-cutoff")) {
            std::string v = value("--completeness-cutoff");
            char *end = nullptr;
            a.completeness_cutoff = strtod(v.c_str(), &end);
            if (v.empty() || end != v.c_str() + v.size())
                throw UsageError("invalid value '" + v + "' for '--completeness-cutoff <COMPLETENESS_CUTOFF>'");
        } else if (is("--threads")) {
            a.threads = parse_uint(value("--threads"), "--threads <THREADS>");
            if (a.threads < 1) throw UsageError("invalid value for '--threads <THREADS>': Threads must be one or higher");
        } else if (is("--gpu")) a.gpu = (int)parse_uint(value("--gpu <ID>"), "--gpu <ID>");
        else if (is("--min-count") || is("--min-qual")) (void)value("option");  // reads are not supported by this build
        else if (t.rfind("--", 0) == 0) throw UsageError("unexpected argument '" + t + "' found");
        else if (shortopt('o')) {
            a.output = value("-o <OUTPUT>");
            a.has_output = true;
        } else if (shortopt('f')) a.file_list = value("-f <FILE_LIST>");
        else if (shortopt('s')) a.sketch_size = (uint64_t)parse_uint(value("-s"), "-s <SKETCH_SIZE>");
        else if (shortopt('k')) a.kmer = (uint64_t)parse_uint(value("-k"), "-k <KMER_LENGTH>");
        else if (t.size() > 1 && t[0] == '-') throw UsageError("unexpected argument '" + t + "' found");
        else if (words == 0) ++words;  // "inverted"
        else if (words == 1) {
            a.cmd = t;
            ++words;
        } else a.pos.push_back(t);
    }
    if (a.cmd != "build" && a.cmd != "query" && a.cmd != "precluster")
        throw UsageError("inverted needs a subcommand: build, query or precluster");
    return a;
}

FILE *open_output(const InvArgs &a) {
    if (!a.has_output) return stdout;
    FILE *f = fopen(a.output.c_str(), "wb");
    if (!f) throw Panic("could not create " + a.output);
    return f;
}

int inverted_build(const InvArgs &a) {
    if (a.output.empty()) throw UsageError("the following required arguments were not provided:\n  -o <OUTPUT>");
    if (a.file_list.empty() == a.pos.empty())
        throw UsageError("exactly one of <SEQ_FILES>... and -f <FILE_LIST> is required");
    log_info("Getting input files");
    InputList inputs = get_input_list(a.file_list, a.pos);
    log_info("Parsed " + std::to_string(inputs.size()) + " samples in input list");
    std::vector<std::string> distinct;  // first-appearance order
    {
        std::unordered_map<std::string, size_t> seen;
        for (auto &in : inputs)
            if (seen.emplace(in.first, distinct.size()).second) distinct.push_back(in.first);
    }
    std::vector<size_t> order(inputs.size());
    std::unordered_map<std::string, std::string> label_of;
    bool have_labels = false;
    if (!a.species_names.empty()) {
        log_info("Reordering samples using labels in " + a.species_names);
        std::vector<std::string> msgs;
        try {
            order = reorder_input_files(inputs, a.species_names, &label_of, &have_labels, &msgs);
        } catch (const std::exception &e) {
            throw Panic(e.what());
        }
        for (auto &m : msgs) log_info(m);
        if (!have_labels) log_warn("Could not find any sample names in " + a.species_names);
    } else {
        std::unordered_map<std::string, size_t> pos;
        for (size_t i = 0; i < distinct.size(); ++i) pos[distinct[i]] = i;
        for (size_t i = 0; i < inputs.size(); ++i) order[i] = pos[inputs[i].first];
    }
    std::optional<std::vector<std::string>> labels, metadata;
    if (have_labels) {
        labels.emplace(distinct.size(), "");
        for (size_t i = 0; i < inputs.size(); ++i) {
            auto it = label_of.find(inputs[i].first);
            if (order[i] < labels->size()) (*labels)[order[i]] = it == label_of.end() ? "" : it->second;
        }
    }
    if (!a.metadata.empty()) {  // parse_metadata_info: src/io.rs:118-137
        std::ifstream f(a.metadata);
        if (!f) throw Panic("Unable to open species name file " + a.metadata);
        std::unordered_map<std::string, std::string> dict;
        std::string line;
        while (std::getline(f, line)) {
            size_t tab = line.find('\t');
            if (tab == std::string::npos) throw Panic("index out of bounds: metadata lines need label<tab>metadata");
            size_t tab2 = line.find('\t', tab + 1);
            if (!dict.emplace(line.substr(0, tab), line.substr(tab + 1, tab2 == std::string::npos ? tab2 : tab2 - tab - 1)).second)
                throw Panic("Some entry in metadata is duplicated");
        }
        log_info("Got metadata for " + std::to_string(dict.size()) + " labels");
        metadata.emplace(distinct.size(), "");
        for (size_t i = 0; i < inputs.size(); ++i) {
            auto it = dict.find(inputs[i].first);
            if (it == dict.end()) throw Panic("no metadata for sample " + inputs[i].first);
            if (order[i] < metadata->size()) (*metadata)[order[i]] = it->second;
        }
    }
    const std::string skq = a.write_skq ? a.output + ".skq" : "";
    log_info("Creating sketches");
    if (!skq.empty()) log_info("Writing bins for use with precluster as " + skq);
    try {
        Inverted inv = Inverted::build(inputs, order, a.kmer, a.sketch_size, !a.single_strand, (int)a.threads, skq, metadata,
                                       labels, nullptr);
        log_info("Saving inverted index to " + a.output + ".ski");
        inv.save(a.output);
        log_info("Index info:\n" + inv.debug_info());
    } catch (const std::exception &e) {
        throw Panic(e.what());
    }
    return 0;
}

Inverted load_ski(const std::string &ski) {
    const std::string prefix = strip_sketch_extension(ski);
    log_info("Loading inverted index from " + prefix + ".ski");
    return Inverted::load(prefix);  // std::runtime_error -> "Error: ..." exit 1, like the reference's `?`
}

int inverted_query(const InvArgs &a) {
    if (a.pos.empty()) throw UsageError("the following required arguments were not provided:\n  <SKI>");
    int mode = a.query_type == "match-count" ? 0 : a.query_type == "all-bins" ? 1 : a.query_type == "any-bins" ? 2 : -1;
    if (mode < 0)
        throw UsageError("invalid value '" + a.query_type + "' for '--query-type <QUERY_TYPE>'\n  [possible values: match-count, all-bins, any-bins]");
    std::vector<std::string> seq_files(a.pos.begin() + 1, a.pos.end());
    if (a.file_list.empty() == seq_files.empty())
        throw UsageError("exactly one of <SEQ_FILES>... and -f <FILE_LIST> is required");
    FILE *out = open_output(a);
    Inverted inv = load_ski(a.pos[0]);
    log_info("Read inverted index:\n" + inv.debug_info());
    log_info("Getting input queries");
    InputList inputs = get_input_list(a.file_list, seq_files);
    log_info("Parsed " + std::to_string(inputs.size()) + " samples in input query list");
    log_info("Sketching input queries");
    std::vector<size_t> order(inputs.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = i;
    std::vector<std::vector<uint16_t>> queries;
    std::vector<std::string> names;
    try {
        Inverted::sketch_files(inputs, order, inv.kmer_size, inv.sketch_size(), inv.rc, (int)a.threads, &queries, &names);
    } catch (const std::exception &e) {
        throw Panic(e.what());
    }
    log_info("Running queries in mode: " + a.query_type);
    std::string text = "Query";
    if (mode == 0) {
        for (auto &n : inv.sample_names) text += "\t" + n;
        text += "\n";
    } else {
        text += "\tMatches\n";
    }
    for (size_t q = 0; q < queries.size(); ++q) {  // the reference writes rows in completion order; here: input order
        text += names[q];
        if (mode == 0) {
            for (uint32_t c : inv.query_against_inverted_index(queries[q].data())) text += "\t" + std::to_string(c);
        } else {
            std::vector<uint32_t> hits = mode == 1 ? inv.all_shared_bins(queries[q].data()) : inv.any_shared_bins(queries[q].data());
            for (size_t h = 0; h < hits.size(); ++h) text += (h ? "," : "\t") + inv.sample_names[hits[h]];
        }
        text += "\n";
    }
    fwrite(text.data(), 1, text.size(), out);
    if (out != stdout) fclose(out);
    return 0;
}

int inverted_precluster(const InvArgs &a) {
    if (a.pos.empty()) throw UsageError("the following required arguments were not provided:\n  <SKI>");
    if (a.pos.size() > 1) throw UsageError("unexpected argument '" + a.pos[1] + "' found");
    if (a.count == !a.skd.empty())
        throw UsageError(a.count ? "the argument '--skd <SKD>' cannot be used with '--count'"
                                 : "the following required arguments were not provided:\n  <--skd <SKD>|--count>");
    int retain = 0;
    if (!a.retain.empty()) {
        retain = a.retain == "singleton" ? 1 : a.retain == "bruteforce" ? 2 : -1;
        if (retain < 0)
            throw UsageError("invalid value '" + a.retain + "' for '--retain-unmatched <RETAIN_UNMATCHED>'\n  [possible values: singleton, bruteforce]");
    }
    Gpu gpu;  // destroyed after the database
    const std::string input_prefix = strip_sketch_extension(a.pos[0]);
    Inverted inv = load_ski(a.pos[0]);
    if (a.count) {
        const uint64_t ns = inv.n_samples();
        printf("Identified %llu prefilter pairs from a max of %llu\n", (unsigned long long)inv.any_shared_bin_pairs(),
               (unsigned long long)(ns * (ns - 1) / 2));
        return 0;
    }
    FILE *out = open_output(a);
    const std::string skq_file = input_prefix + ".skq";
    log_info("Loading queries from " + skq_file);
    const size_t stride = inv.sketch_size();
    std::vector<uint16_t> skq;
    try {
        skq = read_skq(skq_file, inv.n_samples() * stride);
    } catch (const std::exception &e) {
        throw Panic(e.what());
    }
    const std::string ref_name = strip_sketch_extension(a.skd);
    Loaded refs;
    try {
        refs.meta = MultiSketch::load_metadata(ref_name);
    } catch (const std::exception &) {
        throw Panic("Could not read sketch metadata from " + ref_name + ".skm");
    }
    const uint64_t n = refs.meta.number_samples_loaded();
    if (n < 2) throw Panic("attempt to subtract with overflow: fewer than two samples loaded");
    uint64_t knn = a.knn;
    if (knn >= n) {
        log_warn("knn=" + std::to_string(knn) + " is higher than number of samples=" + std::to_string(n));
        knn = n - 1;
    }
    if (knn == 0) throw std::runtime_error("knn must be at least 1");
    if (!refs.meta.get_k_idx(inv.kmer_size))
        throw Panic("K-mer size " + std::to_string(inv.kmer_size) + " used for .ski not found in .skd: K-mer size " +
                    std::to_string(inv.kmer_size) + " not found in file");
    std::vector<double> ref_c;
    if (!a.ref_completeness_file.empty()) {
        std::vector<std::string> warnings;
        ref_c = read_completeness_file(a.ref_completeness_file, refs.meta, &warnings);
        for (auto &w : warnings) log_warn(w);
    }
    // i (skd order) -> ski index, and back (src/distances/mod.rs:412-438)
    std::unordered_map<std::string, size_t> ski_of;
    for (size_t s = 0; s < inv.n_samples(); ++s) ski_of[inv.sample_names[s]] = s;
    std::vector<size_t> ski_index(n);
    std::vector<int64_t> skd_from_ski(inv.n_samples(), -1);
    std::string not_found;
    for (uint64_t i = 0; i < n; ++i) {
        auto it = ski_of.find(refs.meta.sketch_name(i));
        if (it == ski_of.end()) {
            not_found += (not_found.empty() ? "\"" : ", \"") + refs.meta.sketch_name(i) + "\"";
            continue;
        }
        ski_index[i] = it->second;
        skd_from_ski[it->second] = (int64_t)i;
    }
    if (!not_found.empty())
        throw Panic("The following samples in the .skd could not be found in the .ski:\n[" + not_found + "]");
    log_info("Calculating sparse ref vs ref distances with " + std::to_string(knn) + " nearest neighbours");
    log_info("Preclustering with k=" + std::to_string(inv.kmer_size) + " and s=" + std::to_string(stride));
    if (retain) log_info(std::string("Retain unmatched mode: ") + (retain == 1 ? "Retain as singletons" : "Brute-force against all"));

    // Candidate lists (CSR over .skd indices), --threads rows at a time
    std::vector<std::vector<uint64_t>> cand(n);
    {
        std::atomic<uint64_t> next{0};
        auto worker = [&] {
            for (;;) {
                const uint64_t i = next.fetch_add(1);
                if (i >= n) break;
                bool other = false;
                for (uint32_t j : inv.any_shared_bins(skq.data() + ski_index[i] * stride)) {
                    const int64_t skd_j = skd_from_ski[j];
                    if (skd_j < 0) continue;  // in the .ski only
                    cand[i].push_back((uint64_t)skd_j);
                    other |= (uint64_t)skd_j != i;
                }
                if (retain == 2 && !other) {  // brute force: every sample is a candidate (mod.rs:501-524)
                    cand[i].resize(n);
                    for (uint64_t j = 0; j < n; ++j) cand[i][j] = j;
                } else {
                    std::sort(cand[i].begin(), cand[i].end());
                }
            }
        };
        std::vector<std::thread> pool;
        for (long long t = 0; t < a.threads; ++t) pool.emplace_back(worker);
        for (auto &t : pool) t.join();
    }
    std::vector<uint64_t> off(n + 1, 0), flat;
    for (uint64_t i = 0; i < n; ++i) off[i + 1] = off[i] + cand[i].size();
    flat.reserve(off[n]);
    for (auto &c : cand) flat.insert(flat.end(), c.begin(), c.end());
    log_info(std::to_string(off[n]) + " candidate comparisons (of " + std::to_string(n * (n - 1)) + ")");

    if (skd_ctx_create(a.gpu, &gpu.ctx) != SKD_OK) throw std::runtime_error(skd_last_error(nullptr));
    log_info("Loading sketch data from " + ref_name + ".skd");
    load_db(gpu, refs, ref_name, nullptr);
    skd_dist_type dt;
    if (skd_set_k(gpu.ctx, refs.db, (int64_t)inv.kmer_size, a.ani ? 1 : 0, &dt) != SKD_OK)
        throw Panic(std::string("Error setting k size: ") + skd_last_error(gpu.ctx));
    std::vector<uint64_t> idx(n * knn);
    std::vector<float> d0(n * knn);
    gpu.check(skd_self_dists_knn_candidates(gpu.ctx, refs.db, knn, dt, ref_c.empty() ? nullptr : ref_c.data(),
                                            a.completeness_cutoff, off.data(), flat.data(), retain, idx.data(), d0.data()));
    log_info("Writing out in sparse matrix form");
    std::string text;
    for (uint64_t r = 0; r < n; ++r) {
        const std::string &rn = refs.meta.sketch_name(r);
        for (uint64_t t = 0; t < knn; ++t) {
            const std::string &cn = refs.meta.sketch_name(idx[r * knn + t]);
            if (d0[r * knn + t] < 1.0f || cn != rn) append_line(text, rn, cn, d0[r * knn + t], nullptr);
        }
        if (text.size() > (1 << 22)) {
            fwrite(text.data(), 1, text.size(), out);
            text.clear();
        }
    }
    fwrite(text.data(), 1, text.size(), out);
    if (out != stdout) fclose(out);
    return 0;
}

int run_inverted(int argc, char **argv) {
    InvArgs a = parse_inverted_args(argc, argv);
    g_level = a.quiet ? 0 : (a.verbose ? 2 : 1);
    if (a.cmd == "build") return inverted_build(a);
    if (a.cmd == "query") return inverted_query(a);
    return inverted_precluster(a);
}


// ---- `sketchlib info <db|.skm|.ski> [--sample-info]`: src/lib.rs:910-947 (Debug / Display of MultiSketch and Inverted)
int run_info(int argc, char **argv) {
    std::string file;
    bool sample_info = false, verbose = false, quiet = false;
    int words = 0;
    for (int i = 1; i < argc; ++i) {
        std::string t = argv[i];
        if (t == "--sample-info") sample_info = true;
        else if (t == "-v" || t == "--verbose") verbose = true;
        else if (t == "--quiet") quiet = true;
        else if (t.size() > 1 && t[0] == '-') throw UsageError("unexpected argument '" + t + "' found");
        else if (words++ == 0) continue;  // "info"
        else if (file.empty()) file = t;
        else throw UsageError("unexpected argument '" + t + "' found");
    }
    g_level = quiet ? 0 : (verbose ? 2 : 1);
    if (file.empty()) throw UsageError("the following required arguments were not provided:\n  <SKM_FILE>");
    auto ends_with = [&](const char *e) { return file.size() >= 4 && file.compare(file.size() - 4, 4, e) == 0; };
    if (ends_with(".ski")) {
        const std::string prefix = file.substr(0, file.size() - 4);
        Inverted inv;
        try {
            inv = Inverted::load(prefix);
        } catch (const std::exception &e) {
            printf("Read error: %s\n", e.what());
            throw Panic("Could not read inverted index from " + prefix + ".ski");
        }
        if (sample_info) {
            log_info("Printing sample info");
            std::string text = "Name\n";
            for (auto &n : inv.sample_names) text += n + "\n";
            printf("%s\n", text.c_str());
        } else {
            log_info("Printing inverted index info");
            printf("%s\n", inv.debug_info().c_str());
        }
        return 0;
    }
    const std::string prefix = (ends_with(".skm") || ends_with(".skd")) ? file.substr(0, file.size() - 4) : file;
    MultiSketch m;
    try {
        m = MultiSketch::load_metadata(prefix);
    } catch (const std::exception &) {
        throw Panic("Could not read sketch metadata from " + prefix + ".skm");
    }
    if (sample_info) {
        log_info("Printing sample info");
        std::string text = "Name\tSequence length\tBase frequencies\tMissing/ambig bases\tFrom reads\tSingle strand\tDensified\n";
        for (auto &s : m.sketch_metadata)  // impl Display for Sketch, src/sketch/mod.rs:261-277 (A, C, G, T order)
            text += s.name + "\t" + std::to_string(s.seq_length) + "\t[" + std::to_string(s.acgt[0]) + ", " +
                    std::to_string(s.acgt[1]) + ", " + std::to_string(s.acgt[3]) + ", " + std::to_string(s.acgt[2]) + "]\t" +
                    std::to_string(s.non_acgt) + "\t" + (s.reads ? "true" : "false") + "\t" + (!s.rc ? "true" : "false") +
                    "\t" + (s.densified ? "true" : "false") + "\n";
        printf("%s\n", text.c_str());
    } else {
        log_info("Printing database info");
        std::string kmers;
        for (size_t i = 0; i < m.kmer_lengths.size(); ++i) kmers += (i ? ", " : "") + std::to_string(m.kmer_lengths[i]);
        printf("sketch_version=%s\nsequence_type=%s\nsketch_size=%llu\nn_samples=%zu\nkmers=[%s]\ninverted=false\n",
               m.sketch_version.c_str(), m.hash_type.c_str(), (unsigned long long)m.sketch_size, m.sketch_metadata.size(),
               kmers.c_str());
    }
    return 0;
}

int main(int argc, char **argv) {
    auto start = std::chrono::steady_clock::now();
    bool is_sketch = false, is_inverted = false, is_info = false;
    for (int i = 1; i < argc; ++i) {
        if (argv[i][0] == '-') continue;
        is_sketch = std::string(argv[i]) == "sketch";
        is_inverted = std::string(argv[i]) == "inverted";
        is_info = std::string(argv[i]) == "info";
        break;
    }
    if (is_sketch || is_inverted || is_info) {
        try {
            if (is_sketch) run_sketch(argc, argv);
            else if (is_inverted) run_inverted(argc, argv);
            else run_info(argc, argv);
        } catch (const UsageError &e) {
            fprintf(stderr, "error: %s\n", e.what());
            return 2;
        } catch (const Panic &e) {
            fprintf(stderr, "thread 'main' panicked:\n%s\n", e.what());
            return 101;
        } catch (const std::exception &e) {
            fprintf(stderr, "Error: %s\n", e.what());
            return 1;
        }
        if (g_level >= 1 && !is_info) {  // info turns the final message off (src/lib.rs:945)
            auto secs = std::chrono::duration_cast<std::chrono::seconds>(std::chrono::steady_clock::now() - start).count();
            fprintf(stderr, "\xF0\x9F\xA7\xAC\xF0\x9F\x96\x8B\xEF\xB8\x8F sketchlib done in %llds\n", (long long)secs);
        }
        return 0;
    }
    DistArgs a;
    try {
        a = parse_args(argc, argv);
    } catch (const UsageError &e) {
        fprintf(stderr, "error: %s\n\n%s", e.what(), USAGE);
        return 2;
    }
    try {
        run(a);
    } catch (const Panic &e) {
        fprintf(stderr, "thread 'main' panicked:\n%s\n", e.what());
        return 101;
    } catch (const std::exception &e) {
        fprintf(stderr, "Error: %s\n", e.what());
        return 1;
    }
    log_info("Complete");
    if (!a.quiet) {
        auto secs = std::chrono::duration_cast<std::chrono::seconds>(std::chrono::steady_clock::now() - start).count();
        fprintf(stderr, "\xF0\x9F\xA7\xAC\xF0\x9F\x96\x8B\xEF\xB8\x8F sketchlib done in %llds\n", (long long)secs);
    }
    return 0;
}
