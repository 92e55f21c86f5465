// `sketchlib sketch` for DNA assemblies (FASTA, optionally gzipped): host-only C++17.
//
// This is the producer of the hot path's input files (.skm / .skd), shipped because the reference binary cannot
// be built without a Rust toolchain; it is not accelerated (sketching runs once).
// Restates, for DNA FASTA only:
//   src/lib.rs:242-302 (Sketch arm), src/io.rs:20-38, 140-224 (input lists, k-mer parsing),
//   src/hashing/nthash_iterator.rs + src/hashing/mod.rs:82-103 (ntHash with the srol rotation: k-mers never span
//   a non-ACGTU base or a record end; canonical hash = min(forward, reverse complement)),
//   src/sketch/mod.rs:49-54, 132-153, 198-258 (bin minima, densification, 14-plane transpose),
//   src/sketch/sketch_datafile.rs:87-95 (.skd), src/sketch/multisketch.rs:46-87 (.skm: snappy frame + CBOR).
// Reads (FASTQ), amino acids and structures are not supported by this build.
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "multisketch.hpp"

namespace skdhost {

std::string snappy_frame_encode(const std::string &payload);

namespace {
constexpr int kBBits = 14;
constexpr uint64_t kSignMod = ((uint64_t)1 << 61) - 1;  // src/sketch/mod.rs:36
// ntHash seeds (doi:10.1093/bioinformatics/btw397), indexed by (ascii >> 1) & 3: A, C, T/U, G
constexpr uint64_t kSeed[4] = {0x3c8bfbb395c60474ULL, 0x3193c18562a02b4cULL, 0x295549f54be24456ULL,
                               0x20323ed082572324ULL};

inline bool valid_base(unsigned char b) {
    b |= 0x20;
    return b == 'a' || b == 'c' || b == 'g' || b == 't' || b == 'u';
}
inline uint64_t srol(uint64_t v) {  // rotate left by one, then swap bits 0 and 33
    v = (v << 1) | (v >> 63);
    uint64_t x = (v ^ (v >> 33)) & 1;
    return v ^ (x | (x << 33));
}
inline uint64_t sror(uint64_t v) {  // inverse of srol
    uint64_t x = (v ^ (v >> 33)) & 1;
    v ^= (x | (x << 33));
    return (v >> 1) | (v << 63);
}
uint64_t srol_n(uint64_t v, uint64_t n) {
    for (uint64_t i = 0; i < n; ++i) v = srol(v);
    return v;
}

struct Sample {
    std::string name;
    std::vector<std::string> files;
    std::vector<uint64_t> usigs;  // K * W * 14
    uint64_t acgt[4] = {0, 0, 0, 0};
    uint64_t non_acgt = 0;
    bool densified = false;
    std::string error;
};

// All records of all files of a sample; '\n' separates records (a break, like any non-ACGTU byte).
void read_sequences(Sample &s, std::string &seq) {
    for (auto &file : s.files) {
        gzFile f = gzopen(file.c_str(), "rb");
        if (!f) throw std::runtime_error("Invalid path/file: " + file);
        gzbuffer(f, 1 << 20);
        std::vector<char> buf(1 << 20);
        bool in_header = false, at_line_start = true, first = true;
        int n;
        while ((n = gzread(f, buf.data(), (unsigned)buf.size())) > 0) {
            for (int i = 0; i < n; ++i) {
                char c = buf[i];
                if (first) {
                    first = false;
                    if (c == '@') {
                        gzclose(f);
                        throw std::runtime_error(file + " looks like FASTQ: reads are not supported by this build");
                    }
                    if (c != '>') {
                        gzclose(f);
                        throw std::runtime_error("Invalid FASTA/Q record in " + file);
                    }
                }
                if (at_line_start && c == '>') {
                    in_header = true;
                    if (!seq.empty() && seq.back() != '\n') seq.push_back('\n');
                }
                at_line_start = c == '\n';
                if (in_header) {
                    if (c == '\n') in_header = false;
                    continue;
                }
                if (c == '\n' || c == '\r') continue;
                seq.push_back(c);
            }
        }
        gzclose(f);
        if (!seq.empty() && seq.back() != '\n') seq.push_back('\n');
    }
}

bool densify(std::vector<uint64_t> &signs) {  // src/sketch/mod.rs:226-258
    if (*std::max_element(signs.begin(), signs.end()) != UINT64_MAX) return false;
    const uint64_t nb = signs.size();
    for (uint64_t i = 0; i < nb; ++i) {
        uint64_t j = i, attempts = 0;
        while (signs[j] == UINT64_MAX) {
            uint64_t x = i * 1009 + attempts * (1000 * 1000 + 3);
            j = ((x * 48271 + 11) % (((uint64_t)1 << 31) - 1)) % nb;
            ++attempts;
        }
        signs[i] = signs[j];
    }
    return true;
}

// Bin minima of every canonical k-mer hash of `seq` (records separated by '\n'); `signs` must be pre-filled by the
// caller (UINT64_MAX = empty bin).  Returns false when no k-mer fits.  src/sketch/mod.rs:132-153, 156-176.
bool bin_minima(const std::string &seq, uint64_t k, bool rc, std::vector<uint64_t> &signs) {
    const uint64_t nbins = signs.size(), bin_size = (kSignMod + nbins - 1) / nbins;
    uint64_t out_f[4], in_r[4];  // srol^k(seed[b]) leaving the forward hash; srol^(k-1)(seed[rc b]) entering the reverse
    for (int b = 0; b < 4; ++b) {
        out_f[b] = srol_n(kSeed[b], k);
        in_r[b] = srol_n(kSeed[b ^ 2], k - 1);
    }
    uint64_t fh = 0, rh = 0, run = 0;
    bool any = false;
    const size_t len = seq.size();
    for (size_t pos = 0; pos < len; ++pos) {
        const unsigned char c = (unsigned char)seq[pos];
        if (!valid_base(c)) {
            run = 0;
            fh = rh = 0;
            continue;
        }
        const int b = (c >> 1) & 3;
        if (run < k) {  // still filling the first window after a break
            fh = srol(fh) ^ kSeed[b];
            rh ^= srol_n(kSeed[b ^ 2], run);
            ++run;
            if (run < k) continue;
        } else {        // roll: drop seq[pos-k], add seq[pos]
            const int old = ((unsigned char)seq[pos - k] >> 1) & 3;
            fh = srol(fh) ^ kSeed[b] ^ out_f[old];
            rh = sror(rh ^ kSeed[old ^ 2]) ^ in_r[b];
        }
        any = true;
        const uint64_t h = rc ? std::min(fh, rh) : fh;
        const uint64_t sign = h % kSignMod;
        uint64_t &slot = signs[sign / bin_size];
        if (sign < slot) slot = sign;
    }
    return any;
}

void sketch_sample(Sample &s, const std::vector<uint64_t> &kmers, uint64_t W, bool rc) {
    std::string seq;
    read_sequences(s, seq);
    uint64_t total_valid = 0;
    for (unsigned char c : seq) {
        if (c == '\n') continue;
        if (valid_base(c)) {
            s.acgt[(c >> 1) & 3]++;
            ++total_valid;
        } else {
            s.non_acgt++;
        }
    }
    if (total_valid == 0) throw std::runtime_error("Sample " + s.name + " has no valid sequence");
    const uint64_t nbins = 64 * W;
    s.usigs.assign(kmers.size() * W * kBBits, 0);
    std::vector<uint64_t> signs(nbins);
    for (size_t ki = 0; ki < kmers.size(); ++ki) {
        std::fill(signs.begin(), signs.end(), UINT64_MAX);
        if (!bin_minima(seq, kmers[ki], rc, signs))
            throw std::runtime_error("K-mer larger than smallest valid sequence");  // nthash_iterator.rs:60
        s.densified |= densify(signs);
        uint64_t *us = s.usigs.data() + ki * W * kBBits;
        for (uint64_t bin = 0; bin < nbins; ++bin)  // fill_usigs, src/sketch/mod.rs:215-223
            for (int p = 0; p < kBBits; ++p) us[(bin / 64) * kBBits + p] |= ((signs[bin] >> p) & 1) << (bin % 64);
    }
}

// ---- CBOR + snappy-frame writers (uncompressed chunks; any snappy frame reader accepts them)
void cbor_head(std::string &o, int major, uint64_t v) {
    if (v < 24) {
        o.push_back((char)((major << 5) | v));
    } else {
        int info = v < 256 ? 24 : v < 65536 ? 25 : v < (1ull << 32) ? 26 : 27;
        int nb = 1 << (info - 24);
        o.push_back((char)((major << 5) | info));
        for (int i = nb - 1; i >= 0; --i) o.push_back((char)((v >> (8 * i)) & 0xff));
    }
}
void cbor_text(std::string &o, const std::string &s) {
    cbor_head(o, 3, s.size());
    o += s;
}
void cbor_bool(std::string &o, bool b) { o.push_back((char)(b ? 0xf5 : 0xf4)); }

uint32_t crc32c(const unsigned char *p, size_t n) {
    static uint32_t table[256];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
            table[i] = c;
        }
        init = true;
    }
    uint32_t crc = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xff] ^ (crc >> 8);
    return crc ^ 0xFFFFFFFFu;
}

void write_skm(const std::string &prefix, const std::vector<Sample> &samples, const std::vector<uint64_t> &kmers,
               uint64_t W, bool rc, const std::string &version) {
    std::string c;
    cbor_head(c, 5, 10);
    cbor_text(c, "sketch_size");
    cbor_head(c, 0, 64 * W);
    cbor_text(c, "sketchsize64");
    cbor_head(c, 0, W);
    cbor_text(c, "kmer_lengths");
    cbor_head(c, 4, kmers.size());
    for (auto k : kmers) cbor_head(c, 0, k);
    cbor_text(c, "sketch_metadata");
    cbor_head(c, 4, samples.size());
    for (size_t i = 0; i < samples.size(); ++i) {
        const Sample &s = samples[i];
        cbor_head(c, 5, 8);
        cbor_text(c, "name");
        cbor_text(c, s.name);
        cbor_text(c, "index");
        cbor_head(c, 0, i);
        cbor_text(c, "rc");
        cbor_bool(c, rc);
        cbor_text(c, "reads");
        cbor_bool(c, false);
        cbor_text(c, "seq_length");
        cbor_head(c, 0, s.acgt[0] + s.acgt[1] + s.acgt[2] + s.acgt[3]);
        cbor_text(c, "densified");
        cbor_bool(c, s.densified);
        cbor_text(c, "acgt");
        cbor_head(c, 4, 4);
        for (int b = 0; b < 4; ++b) cbor_head(c, 0, s.acgt[b]);
        cbor_text(c, "non_acgt");
        cbor_head(c, 0, s.non_acgt);
    }
    cbor_text(c, "name_map");
    cbor_head(c, 5, samples.size());
    for (size_t i = 0; i < samples.size(); ++i) {
        cbor_text(c, samples[i].name);
        cbor_head(c, 0, i);
    }
    cbor_text(c, "bin_stride");
    cbor_head(c, 0, 1);
    cbor_text(c, "kmer_stride");
    cbor_head(c, 0, W * kBBits);
    cbor_text(c, "sample_stride");
    cbor_head(c, 0, W * kBBits * kmers.size());
    cbor_text(c, "sketch_version");
    cbor_text(c, version);
    cbor_text(c, "hash_type");
    cbor_text(c, "DNA");

    const std::string out = snappy_frame_encode(c);
    std::ofstream f(prefix + ".skm", std::ios::binary);
    if (!f) throw std::runtime_error("could not create " + prefix + ".skm");
    f.write(out.data(), (std::streamsize)out.size());
}

}  // namespace

// Snappy frame stream of uncompressed chunks (any snappy frame reader accepts them).
std::string snappy_frame_encode(const std::string &c) {
    std::string out("\xff\x06\x00\x00sNaPpY", 10);
    for (size_t off = 0; off < c.size(); off += 65536) {
        const size_t n = std::min<size_t>(65536, c.size() - off);
        uint32_t crc = crc32c((const unsigned char *)c.data() + off, n);
        uint32_t masked = ((crc >> 15) | (crc << 17)) + 0xa282ead8u;
        const size_t len = n + 4;
        out.push_back(0x01);  // uncompressed data chunk
        out.push_back((char)(len & 0xff));
        out.push_back((char)((len >> 8) & 0xff));
        out.push_back((char)((len >> 16) & 0xff));
        for (int i = 0; i < 4; ++i) out.push_back((char)((masked >> (8 * i)) & 0xff));
        out.append(c, off, n);
    }
    return out;
}

// Sketch::get_signs_no_densify for one DNA sample at one k over `nbins` bins (src/sketch/mod.rs:156-176; used by the
// inverted index, whose bin count need not be a multiple of 64).  UINT64_MAX marks an empty bin.
std::vector<uint64_t> dna_bin_minima(const std::string &name, const std::vector<std::string> &files, uint64_t k,
                                     uint64_t nbins, bool rc) {
    Sample s;
    s.name = name;
    s.files = files;
    std::string seq;
    read_sequences(s, seq);
    std::vector<uint64_t> signs(nbins, UINT64_MAX);
    bool valid = false;
    for (unsigned char c : seq)
        if (c != '\n' && valid_base(c)) {
            valid = true;
            break;
        }
    if (!valid) throw std::runtime_error("Genome " + name + " has no valid sequence");
    if (!bin_minima(seq, k, rc, signs)) throw std::runtime_error("K-mer larger than smallest valid sequence");
    return signs;
}
bool densify_bins(std::vector<uint64_t> &signs) { return densify(signs); }

// Entry point used by the CLI.  inputs: (name, files).  Throws std::runtime_error.
void sketch_files_dna(const std::string &output_prefix, const std::vector<std::pair<std::string, std::vector<std::string>>> &inputs,
                      std::vector<uint64_t> kmers, uint64_t sketch_size, bool rc, int threads) {
    std::sort(kmers.begin(), kmers.end());  // parse_kmers: src/io.rs:140-160
    if (kmers.empty()) throw std::runtime_error("Must specify --k-vals or --k-seq");
    for (auto k : kmers)
        if (k < 3) throw std::runtime_error("K-mers must be >=3");
    const uint64_t W = (sketch_size + 63) / 64;  // num_bins: src/sketch/mod.rs:49-54
    std::vector<Sample> samples(inputs.size());
    for (size_t i = 0; i < inputs.size(); ++i) {
        samples[i].name = inputs[i].first;
        samples[i].files = inputs[i].second;
    }
    std::atomic<size_t> next{0};
    auto worker = [&] {
        for (;;) {
            size_t i = next.fetch_add(1);
            if (i >= samples.size()) break;
            try {
                sketch_sample(samples[i], kmers, W, rc);
            } catch (const std::exception &e) {
                samples[i].error = e.what();
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < std::max(1, threads); ++t) pool.emplace_back(worker);
    for (auto &t : pool) t.join();
    for (auto &s : samples)
        if (!s.error.empty()) throw std::runtime_error(s.error);
    std::ofstream skd(output_prefix + ".skd", std::ios::binary);  // little-endian u64, samples in input order
    if (!skd) throw std::runtime_error("could not create " + output_prefix + ".skd");
    for (auto &s : samples) skd.write((const char *)s.usigs.data(), (std::streamsize)(s.usigs.size() * 8));
    skd.close();
    write_skm(output_prefix, samples, kmers, W, rc, "0.3.0");
}

}  // namespace skdhost
