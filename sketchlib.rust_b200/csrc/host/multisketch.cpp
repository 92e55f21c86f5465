// See multisketch.hpp.  Host-only C++17.
#include "multisketch.hpp"

#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

namespace skdhost {

static std::runtime_error err(const std::string &m) { return std::runtime_error(m); }

// ---------------------------------------------------------------- snappy (block + framing formats)
static std::string snappy_block(const unsigned char *p, size_t n) {
    size_t pos = 0;
    uint64_t want = 0;
    int shift = 0;
    for (;;) {
        if (pos >= n) throw err("snappy: truncated length");
        unsigned char b = p[pos++];
        want |= (uint64_t)(b & 0x7f) << shift;
        shift += 7;
        if (!(b & 0x80)) break;
        if (shift >= 35) throw err("snappy: length varint longer than 5 bytes");  // the format caps it at 32 bits
    }
    if (want > 0xffffffffull) throw err("snappy: uncompressed length exceeds 32 bits");
    std::string out;
    // a copy element expands at most 64 bytes from 2 input bytes: never trust a declared length beyond that bound
    out.reserve((size_t)std::min<uint64_t>(want, (uint64_t)n * 32 + 64));
    auto need = [&](size_t k) {
        if (pos + k > n) throw err("snappy: truncated block");
    };
    while (pos < n) {
        unsigned char tag = p[pos++];
        unsigned kind = tag & 3;
        if (kind == 0) {  // literal
            size_t len = tag >> 2;
            if (len >= 60) {
                size_t nb = len - 59;
                need(nb);
                len = 0;
                for (size_t i = 0; i < nb; ++i) len |= (size_t)p[pos + i] << (8 * i);
                pos += nb;
            }
            len += 1;
            need(len);
            out.append((const char *)p + pos, len);
            pos += len;
            continue;
        }
        size_t len, off;
        if (kind == 1) {
            need(1);
            len = ((tag >> 2) & 7) + 4;
            off = ((size_t)(tag >> 5) << 8) | p[pos];
            pos += 1;
        } else if (kind == 2) {
            need(2);
            len = (tag >> 2) + 1;
            off = p[pos] | ((size_t)p[pos + 1] << 8);
            pos += 2;
        } else {
            need(4);
            len = (tag >> 2) + 1;
            off = p[pos] | ((size_t)p[pos + 1] << 8) | ((size_t)p[pos + 2] << 16) | ((size_t)p[pos + 3] << 24);
            pos += 4;
        }
        if (off == 0 || off > out.size()) throw err("snappy: bad copy offset");
        for (size_t i = 0; i < len; ++i) out.push_back(out[out.size() - off]);  // may overlap
    }
    if (out.size() != want) throw err("snappy: wrong uncompressed length");
    return out;
}

std::string snappy_frame_decode(const std::string &raw) {
    const unsigned char *p = (const unsigned char *)raw.data();
    size_t n = raw.size(), pos = 0;
    std::string out;
    while (pos < n) {
        if (pos + 4 > n) throw err("snappy frame: truncated chunk header");
        unsigned type = p[pos];
        size_t len = p[pos + 1] | ((size_t)p[pos + 2] << 8) | ((size_t)p[pos + 3] << 16);
        pos += 4;
        if (pos + len > n) throw err("snappy frame: truncated chunk");
        if (type == 0xff) {
            if (len != 6 || memcmp(p + pos, "sNaPpY", 6) != 0) throw err("snappy frame: bad stream identifier");
        } else if (type == 0x00) {
            if (len < 4) throw err("snappy frame: short compressed chunk");
            out += snappy_block(p + pos + 4, len - 4);  // 4-byte masked CRC32C first
        } else if (type == 0x01) {
            if (len < 4) throw err("snappy frame: short raw chunk");
            out.append((const char *)p + pos + 4, len - 4);
        } else if (type >= 0x02 && type <= 0x7f) {
            throw err("snappy frame: reserved unskippable chunk");
        }  // 0x80..0xfe: skippable
        pos += len;
    }
    return out;
}

// ---------------------------------------------------------------- CBOR (RFC 8949), the subset serde emits
const CborValue *CborValue::get(const std::string &key) const {
    for (auto &kv : map)
        if (kv.first->kind == TEXT && kv.first->s == key) return kv.second.get();
    return nullptr;
}

static CborPtr cbor_item(const unsigned char *p, size_t n, size_t &pos, int depth) {
    if (depth > 64) throw err("cbor: nesting too deep");
    if (pos >= n) throw err("cbor: truncated");
    unsigned char ib = p[pos++];
    unsigned major = ib >> 5, info = ib & 31;
    uint64_t val = 0;
    bool indefinite = false;
    if (info < 24) {
        val = info;
    } else if (info <= 27) {
        size_t nb = (size_t)1 << (info - 24);
        if (pos + nb > n) throw err("cbor: truncated argument");
        for (size_t i = 0; i < nb; ++i) val = (val << 8) | p[pos + i];
        pos += nb;
    } else if (info == 31) {
        indefinite = true;
    } else {
        throw err("cbor: reserved additional info");
    }
    auto v = std::make_shared<CborValue>();
    switch (major) {
        case 0:
            v->kind = CborValue::UINT;
            v->u = val;
            return v;
        case 1:
            v->kind = CborValue::NINT;
            v->u = val;
            return v;
        case 2:
        case 3: {
            v->kind = major == 2 ? CborValue::BYTES : CborValue::TEXT;
            if (indefinite) {
                while (true) {
                    if (pos >= n) throw err("cbor: truncated string");
                    if (p[pos] == 0xff) {
                        ++pos;
                        break;
                    }
                    v->s += cbor_item(p, n, pos, depth + 1)->s;
                }
            } else {
                if (pos + val > n) throw err("cbor: truncated string");
                v->s.assign((const char *)p + pos, (size_t)val);
                pos += (size_t)val;
            }
            return v;
        }
        case 4:
            v->kind = CborValue::ARRAY;
            if (indefinite) {
                while (true) {
                    if (pos >= n) throw err("cbor: truncated array");
                    if (p[pos] == 0xff) {
                        ++pos;
                        break;
                    }
                    v->arr.push_back(cbor_item(p, n, pos, depth + 1));
                }
            } else {
                for (uint64_t i = 0; i < val; ++i) v->arr.push_back(cbor_item(p, n, pos, depth + 1));
            }
            return v;
        case 5:
            v->kind = CborValue::MAP;
            if (indefinite) {
                while (true) {
                    if (pos >= n) throw err("cbor: truncated map");
                    if (p[pos] == 0xff) {
                        ++pos;
                        break;
                    }
                    auto k = cbor_item(p, n, pos, depth + 1);
                    v->map.emplace_back(k, cbor_item(p, n, pos, depth + 1));
                }
            } else {
                for (uint64_t i = 0; i < val; ++i) {
                    auto k = cbor_item(p, n, pos, depth + 1);
                    v->map.emplace_back(k, cbor_item(p, n, pos, depth + 1));
                }
            }
            return v;
        case 6:
            return cbor_item(p, n, pos, depth + 1);  // tag: ignore
        default:
            if (info == 20 || info == 21) {
                v->kind = CborValue::BOOL;
                v->b = info == 21;
            } else if (info == 22 || info == 23) {
                v->kind = CborValue::NUL;
            } else if (info == 26) {
                v->kind = CborValue::FLOAT;
                uint32_t bits = (uint32_t)val;
                float f;
                memcpy(&f, &bits, 4);
                v->f = f;
            } else if (info == 27) {
                v->kind = CborValue::FLOAT;
                memcpy(&v->f, &val, 8);
            } else if (info == 25) {  // half precision
                v->kind = CborValue::FLOAT;
                unsigned h = (unsigned)val, e = (h >> 10) & 31, m = h & 1023;
                double d = e == 0 ? std::ldexp((double)m, -24)
                                  : e != 31 ? std::ldexp((double)(m + 1024), (int)e - 25) : (m ? NAN : INFINITY);
                v->f = (h & 0x8000) ? -d : d;
            } else {
                throw err("cbor: unsupported simple value");
            }
            return v;
    }
}

CborPtr cbor_decode(const std::string &buf) {
    size_t pos = 0;
    return cbor_item((const unsigned char *)buf.data(), buf.size(), pos, 0);
}

// ---------------------------------------------------------------- MultiSketch
static uint64_t as_u64(const CborValue *v, const char *what) {
    if (!v || v->kind != CborValue::UINT) throw err(std::string("skm: missing or non-integer field ") + what);
    return v->u;
}

std::string strip_sketch_extension(const std::string &f) {
    if (f.size() >= 4) {
        std::string e = f.substr(f.size() - 4);
        if (e == ".skm" || e == ".skd" || e == ".ski") return f.substr(0, f.size() - 4);
    }
    return f;
}

MultiSketch MultiSketch::load_metadata(const std::string &file_prefix) {
    std::ifstream in(file_prefix + ".skm", std::ios::binary);
    if (!in) throw err("Could not read sketch metadata from " + file_prefix + ".skm");
    std::stringstream ss;
    ss << in.rdbuf();
    CborPtr root = cbor_decode(snappy_frame_decode(ss.str()));
    if (root->kind != CborValue::MAP) throw err("skm: top-level item is not a map");
    MultiSketch m;
    m.prefix = file_prefix;
    m.sketch_size = as_u64(root->get("sketch_size"), "sketch_size");
    if (const CborValue *v = root->get("sketchsize64")) m.sketchsize64 = as_u64(v, "sketchsize64");
    if (m.sketchsize64 == 0) {  // field added in v0.2.0: multisketch.rs:97-100
        m.sketchsize64 = m.sketch_size;
        m.sketch_size *= 64;
    }
    const CborValue *ks = root->get("kmer_lengths");
    if (!ks || ks->kind != CborValue::ARRAY) throw err("skm: missing kmer_lengths");
    for (auto &k : ks->arr) m.kmer_lengths.push_back(as_u64(k.get(), "kmer_lengths[]"));
    const CborValue *md = root->get("sketch_metadata");
    if (!md || md->kind != CborValue::ARRAY) throw err("skm: missing sketch_metadata");
    for (auto &e : md->arr) {
        if (e->kind != CborValue::MAP) throw err("skm: sketch_metadata entry is not a map");
        SketchMeta s;
        const CborValue *nm = e->get("name");
        if (!nm || nm->kind != CborValue::TEXT) throw err("skm: sketch without a name");
        s.name = nm->s;
        s.index = as_u64(e->get("index"), "index");
        if (const CborValue *v = e->get("rc")) s.rc = v->b;
        if (const CborValue *v = e->get("reads")) s.reads = v->b;
        if (const CborValue *v = e->get("densified")) s.densified = v->b;
        if (const CborValue *v = e->get("seq_length")) s.seq_length = v->u;
        if (const CborValue *v = e->get("non_acgt")) s.non_acgt = v->u;
        if (const CborValue *v = e->get("acgt"); v && v->kind == CborValue::ARRAY)
            for (size_t b = 0; b < 4 && b < v->arr.size(); ++b) s.acgt[b] = v->arr[b]->u;
        m.sketch_metadata.push_back(std::move(s));
    }
    if (const CborValue *nmap = root->get("name_map"); nmap && nmap->kind == CborValue::MAP) {
        for (auto &kv : nmap->map)
            if (kv.first->kind == CborValue::TEXT && kv.second->kind == CborValue::UINT)
                m.name_map[kv.first->s] = (size_t)kv.second->u;
    }
    // The reference's `delete` (remove_metadata) shrinks sketch_metadata but leaves name_map stale, so entries that
    // point past the end occur in files the reference itself wrote; Rust panics on the indexing, here every use
    // is bounds-checked (select_subset / get_sample_index) and such entries never reach an array access.
    m.bin_stride = as_u64(root->get("bin_stride"), "bin_stride");
    m.kmer_stride = as_u64(root->get("kmer_stride"), "kmer_stride");
    m.sample_stride = as_u64(root->get("sample_stride"), "sample_stride");
    if (const CborValue *v = root->get("sketch_version"); v && v->kind == CborValue::TEXT) m.sketch_version = v->s;
    if (const CborValue *v = root->get("hash_type")) {
        // serde's externally tagged enum: "DNA" | "PDB" | {"AA": "Level1"}
        if (v->kind == CborValue::TEXT)
            m.hash_type = v->s;
        else if (v->kind == CborValue::MAP && !v->map.empty())
            m.hash_type = v->map[0].first->s + ":" + v->map[0].second->s;
    }
    return m;
}

std::vector<uint64_t> MultiSketch::select_subset(const std::vector<std::string> &names) {
    std::vector<size_t> reindex;
    std::vector<uint64_t> read_indices;
    for (auto &name : names) {
        auto it = name_map.find(name);
        if (it == name_map.end()) throw err("Could not find requested sample " + name + " in sketch metadata");
        if (it->second >= sketch_metadata.size())  // stale name_map entry: the reference panics on this index
            throw err("index out of bounds: the len is " + std::to_string(sketch_metadata.size()) + " but the index is " +
                      std::to_string(it->second) + " (name_map entry of " + name + ")");
        read_indices.push_back(sketch_metadata[it->second].index);
        reindex.push_back(it->second);
    }
    block_reindex = std::move(reindex);
    return read_indices;
}

std::optional<size_t> MultiSketch::get_k_idx(uint64_t k) const {
    for (size_t i = 0; i < kmer_lengths.size(); ++i)
        if (kmer_lengths[i] == k) return i;
    return std::nullopt;
}

std::optional<size_t> MultiSketch::get_sample_index(const std::string &name) const {
    if (block_reindex) {
        for (size_t l = 0; l < block_reindex->size(); ++l)
            if ((*block_reindex)[l] < sketch_metadata.size() && sketch_metadata[(*block_reindex)[l]].name == name) return l;
        return std::nullopt;
    }
    auto it = name_map.find(name);
    if (it == name_map.end()) return std::nullopt;
    if (it->second >= sketch_metadata.size())  // stale name_map entry (see load_metadata)
        throw err("index out of bounds: the len is " + std::to_string(sketch_metadata.size()) + " but the index is " +
                  std::to_string(it->second) + " (name_map entry of " + name + ")");
    return it->second;
}

std::vector<std::string> read_subset_names(const std::string &subset_file) {
    std::ifstream in(subset_file);
    if (!in) throw err("Unable to open " + subset_file);
    std::vector<std::string> names;
    std::string line;
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();  // BufRead::lines strips "\r\n" too
        names.push_back(line);
    }
    return names;
}

// Rust's f64::from_str (src/io.rs:268): decimal digits with optional sign, fraction and exponent, or
// inf / infinity / nan in any case; no hex floats, no locale (strtod accepts both), no surrounding garbage.
static bool parse_rust_f64(const std::string &s, double *out) {
    if (s.empty()) return false;
    size_t p = 0;
    bool neg = false;
    if (s[0] == '+' || s[0] == '-') {
        neg = s[0] == '-';
        p = 1;
    }
    std::string low;
    for (size_t i = p; i < s.size(); ++i) low.push_back((char)tolower((unsigned char)s[i]));
    if (low == "inf" || low == "infinity") {
        *out = neg ? -INFINITY : INFINITY;
        return true;
    }
    if (low == "nan") {
        *out = NAN;
        return true;
    }
    if (low.empty()) return false;
    for (char ch : low)
        if (!(isdigit((unsigned char)ch) || ch == '.' || ch == 'e' || ch == '+' || ch == '-')) return false;
    const char *b = s.data() + p, *e = s.data() + s.size();
    double v = 0.0;
    auto r = std::from_chars(b, e, v, std::chars_format::general);
    if (r.ptr != e) return false;
    if (r.ec == std::errc::result_out_of_range) {  // Rust saturates: huge -> inf, tiny -> 0
        bool big = false;
        size_t epos = low.find('e');
        if (epos != std::string::npos) big = low.find('-', epos) == std::string::npos;
        else big = true;
        v = big ? INFINITY : 0.0;
    } else if (r.ec != std::errc()) {
        return false;
    }
    *out = neg ? -v : v;
    return true;
}

static std::string fmt_f64(double v) {
    if (std::isnan(v)) return "NaN";  // Rust Display
    if (std::isinf(v)) return v < 0 ? "-inf" : "inf";
    char buf[512];
    auto r = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::fixed);
    return std::string(buf, r.ptr);
}

static std::string trim(const std::string &s) {
    size_t a = 0, b = s.size();
    while (a < b && isspace((unsigned char)s[a])) ++a;
    while (b > a && isspace((unsigned char)s[b - 1])) --b;
    return s.substr(a, b - a);
}

static std::string join(const std::vector<std::string> &v, const char *sep) {
    std::string out;
    for (size_t i = 0; i < v.size(); ++i) {
        if (i) out += sep;
        out += v[i];
    }
    return out;
}

std::vector<double> read_completeness_file(const std::string &file, const MultiSketch &sketches,
                                           std::vector<std::string> *warn) {
    std::vector<double> vec(sketches.number_samples_loaded(), 1.0);
    std::ifstream in(file);
    if (!in) throw err("Failed to open completeness file: " + file);
    std::vector<std::string> not_in_sketch, out_of_range;
    std::vector<bool> matched(vec.size(), false);
    std::string line;
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        size_t tab = line.find('\t');
        if (tab == std::string::npos) continue;  // split_once('\t')? -> skipped
        std::string id = line.substr(0, tab), val = trim(line.substr(tab + 1));
        double c = 0.0;
        if (!parse_rust_f64(val, &c)) {
            if (warn) warn->push_back("Could not parse completeness value for '" + id + "': '" + line.substr(tab + 1) + "' — skipping");
            continue;
        }
        if (!(c >= 0.0 && c <= 1.0)) {
            out_of_range.push_back(id + ": " + fmt_f64(c));
            continue;
        }
        if (auto idx = sketches.get_sample_index(id)) {
            vec[*idx] = c;
            matched[*idx] = true;
        } else {
            not_in_sketch.push_back(id);
        }
    }
    if (!out_of_range.empty())
        throw err("Completeness values must be in [0.0, 1.0], not percentages. Found " +
                  std::to_string(out_of_range.size()) + " out-of-range value(s) in " + file + ":\n  " +
                  join(out_of_range, "\n  "));
    if (warn && !not_in_sketch.empty())
        warn->push_back(std::to_string(not_in_sketch.size()) +
                        " genome(s) in completeness file not found in sketch database (ignored): " + join(not_in_sketch, ", "));
    std::vector<std::string> missing;
    for (size_t i = 0; i < matched.size(); ++i)
        if (!matched[i]) missing.push_back(sketches.sketch_name(i));
    if (warn && !missing.empty())
        warn->push_back(std::to_string(missing.size()) + " genome(s) not found in completeness file, using default 1.0: " +
                        join(missing, ", "));
    return vec;
}

std::string fmt_f32(float v) {
    // Rust's `{}`: the shortest digits that round-trip, written positionally (zeros pad a large or a small exponent).
    // to_chars(fixed) alone would print the exact integer digits of large values (3.4e38 -> 339999995214...).
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v < 0 ? "-inf" : "inf";
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::scientific);  // d[.ddd]e[+-]xx, shortest
    std::string sci(buf, r.ptr), out;
    size_t pos = 0;
    if (sci[0] == '-') {
        out = "-";
        pos = 1;
    }
    const size_t e = sci.find('e');
    std::string digits;
    for (size_t i = pos; i < e; ++i)
        if (sci[i] != '.') digits.push_back(sci[i]);
    const int exp10 = atoi(sci.c_str() + e + 1);  // value = d.ddd * 10^exp10
    if (digits == "0") return out + "0";
    const int point = exp10 + 1;  // digits before the decimal point
    if (point <= 0) {
        out += "0.";
        out.append((size_t)(-point), '0');
        out += digits;
    } else if ((size_t)point >= digits.size()) {
        out += digits;
        out.append((size_t)point - digits.size(), '0');
    } else {
        out += digits.substr(0, (size_t)point) + "." + digits.substr((size_t)point);
    }
    return out;
}

}  // namespace skdhost
