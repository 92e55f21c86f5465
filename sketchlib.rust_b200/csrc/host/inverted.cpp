// Inverted index over untransposed sketch bins: builder, queries, `.ski` (snappy frame + MessagePack + portable
// RoaringBitmap) reader/writer and `.skq` reader.  See inverted.hpp for the reference lines each part follows.
#include "inverted.hpp"

#include <algorithm>
#include <atomic>
#include <charconv>
#include <cstring>
#include <fstream>
#include <set>
#include <stdexcept>
#include <thread>

#include "multisketch.hpp"

namespace skdhost {

std::string snappy_frame_encode(const std::string &payload);
std::vector<uint64_t> dna_bin_minima(const std::string &name, const std::vector<std::string> &files, uint64_t k,
                                     uint64_t nbins, bool rc);
bool densify_bins(std::vector<uint64_t> &signs);

namespace {

// ---- MessagePack (the subset rmp-serde emits for this struct: nil, bool, uint, str, bin, array, map)
struct Packer {
    std::string o;
    void be(uint64_t v, int nbytes) {
        for (int i = nbytes - 1; i >= 0; --i) o.push_back((char)((v >> (8 * i)) & 0xff));
    }
    void nil() { o.push_back((char)0xc0); }
    void boolean(bool b) { o.push_back((char)(b ? 0xc3 : 0xc2)); }
    void uint(uint64_t v) {  // smallest encoding, like rmp's write_uint
        if (v < 128) o.push_back((char)v);
        else if (v < 256) { o.push_back((char)0xcc); be(v, 1); }
        else if (v < 65536) { o.push_back((char)0xcd); be(v, 2); }
        else if (v < (1ull << 32)) { o.push_back((char)0xce); be(v, 4); }
        else { o.push_back((char)0xcf); be(v, 8); }
    }
    void str(const std::string &s) {
        const size_t n = s.size();
        if (n < 32) o.push_back((char)(0xa0 | n));
        else if (n < 256) { o.push_back((char)0xd9); be(n, 1); }
        else if (n < 65536) { o.push_back((char)0xda); be(n, 2); }
        else { o.push_back((char)0xdb); be(n, 4); }
        o += s;
    }
    void bin(const std::string &s) {
        const size_t n = s.size();
        if (n < 256) { o.push_back((char)0xc4); be(n, 1); }
        else if (n < 65536) { o.push_back((char)0xc5); be(n, 2); }
        else { o.push_back((char)0xc6); be(n, 4); }
        o += s;
    }
    void array(size_t n) {
        if (n < 16) o.push_back((char)(0x90 | n));
        else if (n < 65536) { o.push_back((char)0xdc); be(n, 2); }
        else { o.push_back((char)0xdd); be(n, 4); }
    }
    void map(size_t n) {
        if (n < 16) o.push_back((char)(0x80 | n));
        else if (n < 65536) { o.push_back((char)0xde); be(n, 2); }
        else { o.push_back((char)0xdf); be(n, 4); }
    }
};

struct Unpacker {
    const unsigned char *p, *end;
    explicit Unpacker(const std::string &s) : p((const unsigned char *)s.data()), end(p + s.size()) {}
    [[noreturn]] static void bad(const char *what) { throw std::runtime_error(std::string("invalid .ski: ") + what); }
    void need(size_t n) const {
        if ((size_t)(end - p) < n) bad("truncated MessagePack data");
    }
    uint64_t be(int nbytes) {
        need(nbytes);
        uint64_t v = 0;
        for (int i = 0; i < nbytes; ++i) v = (v << 8) | *p++;
        return v;
    }
    unsigned char peek() const {
        need(1);
        return *p;
    }
    bool try_nil() {
        if (peek() != 0xc0) return false;
        ++p;
        return true;
    }
    bool boolean() {
        unsigned char t = (unsigned char)be(1);
        if (t == 0xc2) return false;
        if (t == 0xc3) return true;
        bad("expected a bool");
    }
    uint64_t uint() {
        unsigned char t = (unsigned char)be(1);
        if (t < 0x80) return t;
        switch (t) {
            case 0xcc: return be(1);
            case 0xcd: return be(2);
            case 0xce: return be(4);
            case 0xcf: return be(8);
            case 0xd0: case 0xd1: case 0xd2: case 0xd3: {  // non-negative values written as signed
                int nb = 1 << (t - 0xd0);
                uint64_t v = be(nb);
                if (v >> (8 * nb - 1)) bad("negative integer");
                return v;
            }
            default: bad("expected an unsigned integer");
        }
    }
    std::string raw(size_t n) {
        need(n);
        std::string s((const char *)p, n);
        p += n;
        return s;
    }
    std::string str() {
        unsigned char t = (unsigned char)be(1);
        if ((t & 0xe0) == 0xa0) return raw(t & 0x1f);
        if (t == 0xd9) return raw(be(1));
        if (t == 0xda) return raw(be(2));
        if (t == 0xdb) return raw(be(4));
        bad("expected a string");
    }
    std::string bin() {  // serde's bytes: a bin, or (from other writers) an array of small integers
        unsigned char t = peek();
        if (t == 0xc4 || t == 0xc5 || t == 0xc6) {
            ++p;
            return raw(be(1 << (t - 0xc4)));
        }
        size_t n = array();
        std::string s;
        s.reserve(n);
        for (size_t i = 0; i < n; ++i) s.push_back((char)uint());
        return s;
    }
    // every element / entry takes at least one byte: a declared length beyond the bytes left is corrupt (and must
    // not reach a resize)
    size_t bounded(size_t n, size_t min_bytes_each) const {
        if (n > (size_t)(end - p) / min_bytes_each) bad("declared length exceeds the data");
        return n;
    }
    size_t array() {
        unsigned char t = (unsigned char)be(1);
        if ((t & 0xf0) == 0x90) return bounded(t & 0x0f, 1);
        if (t == 0xdc) return bounded(be(2), 1);
        if (t == 0xdd) return bounded(be(4), 1);
        bad("expected an array");
    }
    size_t map() {
        unsigned char t = (unsigned char)be(1);
        if ((t & 0xf0) == 0x80) return bounded(t & 0x0f, 2);
        if (t == 0xde) return bounded(be(2), 2);
        if (t == 0xdf) return bounded(be(4), 2);
        bad("expected a map");
    }
};

// ---- portable RoaringBitmap serialisation (RoaringFormatSpec).  Writer: the run-free layout roaring 0.10 emits
// (cookie 12346; array containers up to 4096 values, bitmap containers above).  Reader: also run containers.
constexpr uint32_t kCookieNoRun = 12346, kCookieRun = 12347;

void le(std::string &o, uint64_t v, int nbytes) {
    for (int i = 0; i < nbytes; ++i) o.push_back((char)((v >> (8 * i)) & 0xff));
}

std::string roaring_serialise(const std::vector<uint32_t> &sorted) {
    struct Cont {
        uint16_t key;
        size_t begin, end;
    };
    std::vector<Cont> conts;
    for (size_t i = 0; i < sorted.size();) {
        size_t j = i;
        const uint16_t key = (uint16_t)(sorted[i] >> 16);
        while (j < sorted.size() && (uint16_t)(sorted[j] >> 16) == key) ++j;
        conts.push_back({key, i, j});
        i = j;
    }
    std::string o;
    le(o, kCookieNoRun, 4);
    le(o, conts.size(), 4);
    for (auto &c : conts) {
        le(o, c.key, 2);
        le(o, c.end - c.begin - 1, 2);
    }
    uint32_t offset = 8 + 8 * (uint32_t)conts.size();
    for (auto &c : conts) {
        le(o, offset, 4);
        const size_t card = c.end - c.begin;
        offset += card > 4096 ? 8192 : (uint32_t)(2 * card);
    }
    for (auto &c : conts) {
        const size_t card = c.end - c.begin;
        if (card > 4096) {
            std::vector<uint64_t> bits(1024, 0);
            for (size_t i = c.begin; i < c.end; ++i) bits[(sorted[i] & 0xffff) >> 6] |= 1ull << (sorted[i] & 63);
            for (uint64_t w : bits) le(o, w, 8);
        } else {
            for (size_t i = c.begin; i < c.end; ++i) le(o, sorted[i] & 0xffff, 2);
        }
    }
    return o;
}

std::vector<uint32_t> roaring_deserialise(const std::string &s) {
    const unsigned char *p = (const unsigned char *)s.data(), *end = p + s.size();
    auto rd = [&](int nbytes) -> uint64_t {
        if (end - p < nbytes) throw std::runtime_error("invalid .ski: truncated bitmap");
        uint64_t v = 0;
        for (int i = 0; i < nbytes; ++i) v |= (uint64_t)p[i] << (8 * i);
        p += nbytes;
        return v;
    };
    const uint32_t cookie = (uint32_t)rd(4);
    size_t n;
    std::vector<bool> is_run;
    if ((cookie & 0xffff) == kCookieRun) {
        n = (cookie >> 16) + 1;
        is_run.assign(n, false);
        for (size_t i = 0; i < (n + 7) / 8; ++i) {
            unsigned char b = (unsigned char)rd(1);
            for (int bit = 0; bit < 8 && i * 8 + bit < n; ++bit) is_run[i * 8 + bit] = (b >> bit) & 1;
        }
    } else if (cookie == kCookieNoRun) {
        n = rd(4);
        if (n > (size_t)(end - p) / 4) throw std::runtime_error("invalid .ski: bitmap container count exceeds the data");
        is_run.assign(n, false);
    } else {
        throw std::runtime_error("invalid .ski: unknown bitmap cookie");
    }
    std::vector<std::pair<uint16_t, uint32_t>> desc(n);  // key, cardinality
    for (auto &d : desc) {
        d.first = (uint16_t)rd(2);
        d.second = (uint32_t)rd(2) + 1;
    }
    if (cookie == kCookieNoRun || n >= 4) {  // offset header (not needed for a sequential read)
        if ((size_t)(end - p) < 4 * n) throw std::runtime_error("invalid .ski: truncated bitmap");
        p += 4 * n;
    }
    std::vector<uint32_t> out;
    for (size_t c = 0; c < n; ++c) {
        const uint32_t hi = (uint32_t)desc[c].first << 16;
        if (is_run[c]) {
            const size_t runs = rd(2);
            for (size_t r = 0; r < runs; ++r) {
                const uint32_t start = (uint32_t)rd(2), len = (uint32_t)rd(2);
                for (uint32_t v = start; v <= start + len; ++v) out.push_back(hi | v);
            }
        } else if (desc[c].second > 4096) {
            for (uint32_t w = 0; w < 1024; ++w) {
                uint64_t bits = rd(8);
                while (bits) {
                    out.push_back(hi | (w << 6) | (uint32_t)__builtin_ctzll(bits));
                    bits &= bits - 1;
                }
            }
        } else {
            for (uint32_t i = 0; i < desc[c].second; ++i) out.push_back(hi | (uint32_t)rd(2));
        }
    }
    return out;
}

std::string fmt_f64(double v) {  // Rust `{}` for f64: shortest round-trip digits, positional
    char buf[400];
    auto r = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::fixed);
    return std::string(buf, r.ptr);
}

}  // namespace

// ---------------------------------------------------------------- build

void Inverted::sketch_files(const InputList &inputs, const std::vector<size_t> &file_order, uint64_t k,
                            uint64_t sketch_size, bool rc, int threads, std::vector<std::vector<uint16_t>> *sketches,
                            std::vector<std::string> *names) {
    std::set<std::string> seen, multi;
    for (auto &in : inputs)
        if (!seen.insert(in.first).second) multi.insert(in.first);
    std::vector<std::vector<uint64_t>> signs(inputs.size());
    std::vector<std::string> errors(inputs.size());
    std::atomic<size_t> next{0};
    auto worker = [&] {
        for (;;) {
            size_t i = next.fetch_add(1);
            if (i >= inputs.size()) break;
            try {
                signs[i] = dna_bin_minima(inputs[i].first, inputs[i].second, k, sketch_size, rc);
            } catch (const std::exception &e) {
                errors[i] = e.what();
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < std::max(1, threads); ++t) pool.emplace_back(worker);
    for (auto &t : pool) t.join();
    for (auto &e : errors)
        if (!e.empty()) throw std::runtime_error(e);

    sketches->assign(seen.size(), {});
    names->assign(seen.size(), "");
    for (size_t i = 0; i < inputs.size(); ++i) {
        const size_t g = file_order[i];
        if (g >= sketches->size()) throw std::runtime_error("sample order out of range (repeated sample names with labels?)");
        (*names)[g] = inputs[i].first;
        std::vector<uint16_t> &dst = (*sketches)[g];
        if (dst.empty()) {
            // A sample given once is densified here; the parts of a multi-entry sample are merged by their truncated
            // minima (src/inverted.rs:379-400; the reference's later densify pass cannot fire on u16 values)
            if (!multi.count(inputs[i].first)) densify_bins(signs[i]);
            dst.resize(sketch_size);
            for (uint64_t b = 0; b < sketch_size; ++b) dst[b] = (uint16_t)signs[i][b];
        } else {
            for (uint64_t b = 0; b < sketch_size; ++b) dst[b] = std::min(dst[b], (uint16_t)signs[i][b]);
        }
    }
}

Inverted Inverted::build(const InputList &inputs, const std::vector<size_t> &file_order, uint64_t k,
                         uint64_t sketch_size, bool rc, int threads, const std::string &skq_file,
                         const std::optional<std::vector<std::string>> &metadata,
                         const std::optional<std::vector<std::string>> &labels, std::vector<uint16_t> *bins_out) {
    std::vector<std::vector<uint16_t>> sketches;
    Inverted inv;
    sketch_files(inputs, file_order, k, sketch_size, rc, threads, &sketches, &inv.sample_names);
    std::vector<uint16_t> flat;
    flat.reserve(sketches.size() * sketch_size);
    for (auto &s : sketches) flat.insert(flat.end(), s.begin(), s.end());
    if (!skq_file.empty()) {  // SketchArrayWriter::write_sketch: little-endian u16, index order
        std::ofstream f(skq_file, std::ios::binary);
        if (!f) throw std::runtime_error("Couldn't write to " + skq_file);
        std::string bytes;
        bytes.reserve(flat.size() * 2);
        for (uint16_t v : flat) le(bytes, v, 2);
        f.write(bytes.data(), (std::streamsize)bytes.size());
    }
    inv.index.assign(sketch_size, {});
    for (size_t g = 0; g < sketches.size(); ++g)  // ascending g => ascending lists
        for (uint64_t b = 0; b < sketch_size; ++b) inv.index[b][sketches[g][b]].push_back((uint32_t)g);
    inv.metadata = metadata;
    inv.labels = labels;
    inv.kmer_size = k;
    inv.sketch_version = "0.3.0";
    inv.rc = rc;
    if (bins_out) bins_out->swap(flat);
    return inv;
}

// ---------------------------------------------------------------- queries

std::vector<uint32_t> Inverted::query_against_inverted_index(const uint16_t *sigs) const {
    std::vector<uint32_t> counts(n_samples(), 0);
    for (size_t b = 0; b < index.size(); ++b) {
        auto it = index[b].find(sigs[b]);
        if (it != index[b].end())
            for (uint32_t s : it->second) counts[s]++;
    }
    return counts;
}

std::vector<uint32_t> Inverted::all_shared_bins(const uint16_t *sigs) const {
    std::vector<uint32_t> counts = query_against_inverted_index(sigs), out;
    for (uint32_t s = 0; s < counts.size(); ++s)
        if (counts[s] == index.size()) out.push_back(s);
    return out;
}

std::vector<uint32_t> Inverted::any_shared_bins(const uint16_t *sigs) const {
    std::vector<uint32_t> out;
    for (size_t b = 0; b < index.size(); ++b) {
        auto it = index[b].find(sigs[b]);
        if (it == index[b].end()) continue;
        std::vector<uint32_t> merged;
        merged.reserve(out.size() + it->second.size());
        std::set_union(out.begin(), out.end(), it->second.begin(), it->second.end(), std::back_inserter(merged));
        out.swap(merged);
    }
    return out;
}

uint64_t Inverted::any_shared_bin_pairs() const {
    // Distinct unordered pairs sharing any (bin, value): per sample, the union of its lists restricted to larger ids
    const size_t n = n_samples();
    std::vector<std::vector<const std::vector<uint32_t> *>> lists(n);
    for (auto &bin : index)
        for (auto &kv : bin)
            if (kv.second.size() > 1)
                for (uint32_t s : kv.second) lists[s].push_back(&kv.second);
    uint64_t pairs = 0;
    std::vector<uint32_t> stamp(n, UINT32_MAX);
    for (uint32_t s = 0; s < n; ++s)
        for (auto *l : lists[s])
            for (auto it = std::upper_bound(l->begin(), l->end(), s); it != l->end(); ++it)
                if (stamp[*it] != s) {
                    stamp[*it] = s;
                    ++pairs;
                }
    return pairs;
}

std::string Inverted::debug_info() const {
    size_t mx = 0, mn = SIZE_MAX, sum = 0;
    for (auto &bin : index) {
        mx = std::max(mx, bin.size());
        mn = std::min(mn, bin.size());
        sum += bin.size();
    }
    std::string o = "sketch_version=" + sketch_version + "\nsequence_type=" + hash_type +
                    "\nsketch_size=" + std::to_string(index.size()) + "\nn_samples=" + std::to_string(n_samples()) +
                    "\nkmer=" + std::to_string(kmer_size) + "\nrc=" + (rc ? "true" : "false") + "\ninverted=true\n";
    o += "max_hashes_per_bin=" + std::to_string(mx) + "\nmin_hashes_per_bin=" + std::to_string(index.empty() ? 0 : mn) +
         "\navg_hashes_per_bin=" + fmt_f64(index.empty() ? 0.0 : (double)sum / (double)index.size());
    return o;
}

// ---------------------------------------------------------------- .ski

std::string Inverted::serialise() const {
    Packer m;
    m.array(9);
    m.array(index.size());
    for (auto &bin : index) {
        m.map(bin.size());
        for (auto &kv : bin) {
            m.uint(kv.first);
            m.bin(roaring_serialise(kv.second));
        }
    }
    m.uint(n_samples());
    m.array(sample_names.size());
    for (auto &s : sample_names) m.str(s);
    for (auto *opt : {&metadata, &labels}) {
        if (!*opt) {
            m.nil();
        } else {
            m.array((*opt)->size());
            for (auto &s : **opt) m.str(s);
        }
    }
    m.uint(kmer_size);
    m.str(sketch_version);
    m.boolean(rc);
    m.str(hash_type);  // unit variant HashType::DNA
    return m.o;
}

Inverted Inverted::deserialise(const std::string &buf) {
    Unpacker u(buf);
    if (u.array() != 9) Unpacker::bad("expected the 9 fields of Inverted");
    Inverted inv;
    inv.index.resize(u.array());
    for (auto &bin : inv.index) {
        const size_t entries = u.map();
        bin.reserve(entries);
        for (size_t e = 0; e < entries; ++e) {
            const uint16_t key = (uint16_t)u.uint();
            bin[key] = roaring_deserialise(u.bin());
        }
    }
    const uint64_t n = u.uint();
    inv.sample_names.resize(u.array());
    for (auto &s : inv.sample_names) s = u.str();
    if (n != inv.sample_names.size()) Unpacker::bad("n_samples does not match the sample names");
    for (auto &bin : inv.index)
        for (auto &kv : bin) {
            if (!std::is_sorted(kv.second.begin(), kv.second.end())) Unpacker::bad("unsorted bitmap");
            if (!kv.second.empty() && kv.second.back() >= n) Unpacker::bad("sample index beyond n_samples in a bitmap");
        }
    for (auto *opt : {&inv.metadata, &inv.labels}) {
        if (u.try_nil()) continue;
        opt->emplace(u.array());
        for (auto &s : **opt) s = u.str();
    }
    inv.kmer_size = u.uint();
    inv.sketch_version = u.str();
    inv.rc = u.boolean();
    const unsigned char t = u.peek();
    if ((t & 0xf0) == 0x80) Unpacker::bad("only DNA indexes are supported (hash_type is not a unit variant)");
    inv.hash_type = u.str();
    if (inv.hash_type != "DNA") Unpacker::bad("only DNA indexes are supported");
    return inv;
}

void Inverted::save(const std::string &file_prefix) const {
    const std::string out = snappy_frame_encode(serialise());
    std::ofstream f(file_prefix + ".ski", std::ios::binary);
    if (!f) throw std::runtime_error("could not create " + file_prefix + ".ski");
    f.write(out.data(), (std::streamsize)out.size());
}

Inverted Inverted::load(const std::string &file_prefix) {
    std::ifstream f(file_prefix + ".ski", std::ios::binary);
    if (!f) throw std::runtime_error("No such file or directory: " + file_prefix + ".ski");
    std::string raw((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    return deserialise(snappy_frame_decode(raw));
}

// ---------------------------------------------------------------- inputs

std::vector<size_t> reorder_input_files(const InputList &inputs, const std::string &species_name_file,
                                        std::unordered_map<std::string, std::string> *labels, bool *found_any,
                                        std::vector<std::string> *log) {
    std::set<std::string> input_names;
    for (auto &in : inputs) input_names.insert(in.first);
    std::ifstream f(species_name_file);
    if (!f) throw std::runtime_error("Unable to\n        open species name file " + species_name_file);
    std::unordered_map<std::string, size_t> species_idx;
    std::vector<std::pair<std::string, size_t>> label_order;
    std::string line;
    while (std::getline(f, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        std::vector<std::string> fields;  // split_terminator('\t')
        size_t pos = 0;
        while (pos < line.size()) {
            size_t t = line.find('\t', pos);
            if (t == std::string::npos) t = line.size();
            fields.push_back(line.substr(pos, t - pos));
            pos = t + 1;
        }
        if (fields.size() < 2) throw std::runtime_error("index out of bounds: species name lines need name<tab>label");
        if (input_names.count(fields[0])) {
            auto it = species_idx.find(fields[1]);
            if (it == species_idx.end()) it = species_idx.emplace(fields[1], species_idx.size()).first;
            label_order.push_back({fields[0], it->second});
        }
        (*labels)[fields[0]] = fields[1];
    }
    log->push_back(std::to_string(label_order.size()) + " samples with " + std::to_string(species_idx.size()) +
                   " unique labels");
    std::stable_sort(label_order.begin(), label_order.end(),
                     [](const auto &a, const auto &b) { return a.second < b.second; });
    std::unordered_map<std::string, size_t> reordered;
    for (size_t i = 0; i < label_order.size(); ++i) reordered[label_order[i].first] = i;
    std::vector<size_t> order(inputs.size());
    *found_any = !reordered.empty();
    if (reordered.empty()) {
        for (size_t i = 0; i < inputs.size(); ++i) order[i] = i;
        labels->clear();
        return order;
    }
    size_t next = reordered.size() - 1, found = 0;
    for (size_t i = 0; i < inputs.size(); ++i) {
        auto it = reordered.find(inputs[i].first);
        if (it != reordered.end()) {
            order[i] = it->second;
            ++found;
        } else {
            order[i] = ++next;
        }
    }
    log->push_back("Found " + std::to_string(found) + " of " + std::to_string(inputs.size()) +
                   " input samples with given labels");
    return order;
}

std::vector<uint16_t> read_skq(const std::string &path, size_t n_values) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("Couldn't read from " + path);
    std::string raw((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    std::vector<uint16_t> out(raw.size() / 2);
    for (size_t i = 0; i < out.size(); ++i)
        out[i] = (uint16_t)((unsigned char)raw[2 * i] | ((unsigned char)raw[2 * i + 1] << 8));
    if (out.size() < n_values)
        throw std::runtime_error(path + " holds " + std::to_string(out.size()) + " bins, expected " + std::to_string(n_values));
    return out;
}

}  // namespace skdhost
