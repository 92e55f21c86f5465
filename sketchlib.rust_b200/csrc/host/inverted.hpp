// Host-side inverted index (`.ski`) and untransposed bins (`.skq`) as the reference's `inverted build` /
// `inverted precluster` / `inverted query` commands use them (C++17, no CUDA).
//   reference: src/inverted.rs:46-57 (fields), :64-113 (new), :188-210 (save/load), :229-268 (queries),
//              :271-300 (any_shared_bin_list), :303-414 (sketch_files_inverted), :465-495 (build_inverted_index)
// The index is the candidate generator of the precluster distance path; the distances
// themselves run on the GPU (skd_self_dists_knn_candidates).
// Container: snappy *frame* stream around one MessagePack value (crates snap 1.1 / rmp-serde 1.3): the struct
// as a 9-element array in field order; every bitmap is the portable RoaringBitmap serialisation (crate roaring
// 0.10) in a msgpack `bin`.  Written here from the published formats (msgpack spec, RoaringFormatSpec); the
// reference ships no `.ski` fixture, so the byte format is pinned only by round trips and by an independent
// MessagePack decoder in the tests, while the index *content* is pinned by the reference's `.skq` fixture and
// its query / precluster goldens.
#pragma once
#include <cstdint>
#include <optional>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

namespace skdhost {

using InputList = std::vector<std::pair<std::string, std::vector<std::string>>>;  // (sample name, files)

class Inverted {
  public:
    // index[bin][value] = ascending sample indices holding `value` in that bin
    std::vector<std::unordered_map<uint16_t, std::vector<uint32_t>>> index;
    std::vector<std::string> sample_names;
    std::optional<std::vector<std::string>> metadata, labels;
    uint64_t kmer_size = 0;
    std::string sketch_version;
    bool rc = true;
    std::string hash_type = "DNA";

    size_t n_samples() const { return sample_names.size(); }
    size_t sketch_size() const { return index.size(); }

    // Inverted::new.  `file_order[i]` is the index position of input i.  The untransposed bins (n x sketch_size
    // u16, index order) are returned in *bins_out and, if skq_file is not empty, written there.
    static Inverted build(const InputList &inputs, const std::vector<size_t> &file_order, uint64_t k,
                          uint64_t sketch_size, bool rc, int threads, const std::string &skq_file,
                          const std::optional<std::vector<std::string>> &metadata,
                          const std::optional<std::vector<std::string>> &labels, std::vector<uint16_t> *bins_out);
    // sketch_files_inverted: densified u16 bins per distinct sample + names in index order
    static void sketch_files(const InputList &inputs, const std::vector<size_t> &file_order, uint64_t k,
                             uint64_t sketch_size, bool rc, int threads, std::vector<std::vector<uint16_t>> *sketches,
                             std::vector<std::string> *names);

    void save(const std::string &file_prefix) const;
    static Inverted load(const std::string &file_prefix);
    std::string serialise() const;                       // the MessagePack value (before the snappy frame)
    static Inverted deserialise(const std::string &buf);

    std::vector<uint32_t> query_against_inverted_index(const uint16_t *sigs) const;  // :229-240
    std::vector<uint32_t> all_shared_bins(const uint16_t *sigs) const;               // :243-256
    std::vector<uint32_t> any_shared_bins(const uint16_t *sigs) const;               // :259-268
    uint64_t any_shared_bin_pairs() const;  // |any_shared_bin_list()|, :271-300
    std::string debug_info() const;         // impl Debug, :516-543
};

// reorder_input_files (src/io.rs:40-114): index position of every input from a species/cluster label file; the
// name -> label map is returned through *labels (left empty when no input name occurs in the file).
std::vector<size_t> reorder_input_files(const InputList &inputs, const std::string &species_name_file,
                                        std::unordered_map<std::string, std::string> *labels, bool *found_any,
                                        std::vector<std::string> *log);
// `.skq`: n_samples x sketch_size little-endian u16 (SketchArrayReader::read_all_from_skq)
std::vector<uint16_t> read_skq(const std::string &path, size_t n_values);

}  // namespace skdhost
