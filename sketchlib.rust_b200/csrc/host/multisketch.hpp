// Host-side mirror of the reference's MultiSketch as the `dist` command uses it (C++17, no CUDA).
//   reference: src/sketch/multisketch.rs:23-43 (fields), :90-103 (load_metadata, legacy fix-up),
//              :106-165 (accessors), :187-210 (subset -> block_reindex)
// The `.skm` container is a snappy *frame* stream around one CBOR map (crates snap 1.1 / ciborium 0.2 in
// the reference); both decoders are written here from the published formats (Google snappy
// framing_format.txt / format_description.txt, RFC 8949) and pinned by the four `.skm` fixtures the
// reference ships (tests/test_host_cli.py).
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace skdhost {

struct CborValue;
using CborPtr = std::shared_ptr<CborValue>;
struct CborValue {
    enum Kind { UINT, NINT, BYTES, TEXT, ARRAY, MAP, BOOL, NUL, FLOAT } kind = NUL;
    uint64_t u = 0;   // UINT value, or -1 - value for NINT
    double f = 0.0;
    bool b = false;
    std::string s;    // TEXT / BYTES
    std::vector<CborPtr> arr;
    std::vector<std::pair<CborPtr, CborPtr>> map;
    const CborValue *get(const std::string &key) const;  // MAP lookup by text key (nullptr if absent)
};

std::string snappy_frame_decode(const std::string &raw);
CborPtr cbor_decode(const std::string &buf);

struct SketchMeta {  // src/sketch/mod.rs:57-70 (serialised fields of Sketch)
    std::string name;
    uint64_t index = 0;
    bool rc = true, reads = false, densified = false;
    uint64_t seq_length = 0;
    uint64_t acgt[4] = {0, 0, 0, 0};  // indexed by (ascii >> 1) & 3: A, C, T, G
    uint64_t non_acgt = 0;
};

class MultiSketch {
  public:
    uint64_t sketch_size = 0;   // bins
    uint64_t sketchsize64 = 0;
    std::vector<uint64_t> kmer_lengths;
    std::vector<SketchMeta> sketch_metadata;
    std::unordered_map<std::string, size_t> name_map;
    uint64_t bin_stride = 1, kmer_stride = 0, sample_stride = 0;
    std::string sketch_version, hash_type;
    std::optional<std::vector<size_t>> block_reindex;  // set by select_subset (read_sketch_data_block)
    std::string prefix;

    // load_metadata: multisketch.rs:90-103.  Throws std::runtime_error.
    static MultiSketch load_metadata(const std::string &file_prefix);

    // The name -> metadata index part of read_sketch_data_block (multisketch.rs:187-200); returns the FILE
    // positions (Sketch::index) to hand to skd_db_load_skd.  Throws "Could not find requested sample ...".
    std::vector<uint64_t> select_subset(const std::vector<std::string> &names);

    size_t number_samples_loaded() const {  // :106-111
        return block_reindex ? block_reindex->size() : sketch_metadata.size();
    }
    const std::string &sketch_name(size_t idx) const {  // :139-144
        return block_reindex ? sketch_metadata[(*block_reindex)[idx]].name : sketch_metadata[idx].name;
    }
    std::optional<size_t> get_k_idx(uint64_t k) const;                   // :115-121
    std::optional<size_t> get_sample_index(const std::string &name) const;  // :148-164
    bool is_compatible_with(const MultiSketch &o) const {                // :222-226
        return kmer_lengths == o.kmer_lengths && sketch_size == o.sketch_size && hash_type == o.hash_type;
    }
};

// strip_sketch_extension: src/utils.rs:9-15
std::string strip_sketch_extension(const std::string &file_name);
// read_subset_names: src/io.rs:227-236
std::vector<std::string> read_subset_names(const std::string &subset_file);
// read_completeness_file: src/io.rs:240-324.  Warnings go to `warn` (one string per log::warn!).
std::vector<double> read_completeness_file(const std::string &completeness_file, const MultiSketch &sketches,
                                           std::vector<std::string> *warn);
// Rust `{}` for f32: shortest digits that round-trip, positional notation ("1", "0.008820182").
std::string fmt_f32(float v);

}  // namespace skdhost
