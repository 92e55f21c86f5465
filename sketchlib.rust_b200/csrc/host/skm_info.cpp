// skm_info <db prefix|.skm> [--subset file] [--completeness file]: prints what the host layer read from a
// sketch database as JSON-ish lines.  Host-only; used by the CPU tests to pin the C++ .skm reader against
// the reference's fixtures (and the oracle's independent Python reader).
#include <cstdio>
#include <cstring>
#include <string>

#include "multisketch.hpp"

using namespace skdhost;

int main(int argc, char **argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: skm_info <db> [--subset file] [--completeness file]\n");
        return 2;
    }
    try {
        MultiSketch m = MultiSketch::load_metadata(strip_sketch_extension(argv[1]));
        std::vector<uint64_t> file_idx;
        for (int i = 2; i + 1 < argc; i += 2) {
            if (!strcmp(argv[i], "--subset")) file_idx = m.select_subset(read_subset_names(argv[i + 1]));
        }
        printf("sketch_size\t%llu\nsketchsize64\t%llu\nbin_stride\t%llu\nkmer_stride\t%llu\nsample_stride\t%llu\n",
               (unsigned long long)m.sketch_size, (unsigned long long)m.sketchsize64, (unsigned long long)m.bin_stride,
               (unsigned long long)m.kmer_stride, (unsigned long long)m.sample_stride);
        printf("sketch_version\t%s\nhash_type\t%s\nkmer_lengths", m.sketch_version.c_str(), m.hash_type.c_str());
        for (auto k : m.kmer_lengths) printf("\t%llu", (unsigned long long)k);
        printf("\nn_loaded\t%zu\n", m.number_samples_loaded());
        for (size_t i = 0; i < m.number_samples_loaded(); ++i) {
            size_t mi = m.block_reindex ? (*m.block_reindex)[i] : i;
            printf("sample\t%zu\t%s\t%llu\n", i, m.sketch_name(i).c_str(), (unsigned long long)m.sketch_metadata[mi].index);
        }
        for (int i = 2; i + 1 < argc; i += 2) {
            if (!strcmp(argv[i], "--completeness")) {
                std::vector<std::string> warn;
                auto c = read_completeness_file(argv[i + 1], m, &warn);
                for (size_t j = 0; j < c.size(); ++j) printf("completeness\t%zu\t%.17g\n", j, c[j]);
                for (auto &w : warn) printf("warn\t%s\n", w.c_str());
            }
        }
        for (int i = 2; i < argc; ++i)
            if (!strcmp(argv[i], "--fmt") && i + 1 < argc) {
                float f = strtof(argv[i + 1], nullptr);
                printf("fmt\t%s\n", fmt_f32(f).c_str());
            }
    } catch (const std::exception &e) {
        fprintf(stderr, "Error: %s\n", e.what());
        return 1;
    }
    return 0;
}
