// deacon_host.hpp -- the host side above the C ABI (include/deacon_cuda.h), in C++ because the reference's
// toolchain (cargo / rustc) is not available in this image.  It mirrors the reference's library surface
// (src/lib.rs:26-286): FilterConfig / IndexConfig with the same fields and defaults, run_filter,
// build_index, index_info, union_index, diff_index, FilterSummary.  The arithmetic of the hot path
// (extraction, lookup, classification, set algebra, .idx codec) all happens on the GPU behind the C ABI;
// this layer reads FASTA/FASTQ, batches records into pinned buffers, and writes records and summaries in
// the reference's formats (src/local_filter.rs:60-92, 575-810).
#pragma once
#include <cstdint>
#include <optional>
#include <string>
#include <vector>

namespace deacon {

constexpr const char *VERSION = "0.10.0-b200";   // CARGO_PKG_VERSION of the reference + this build

// JSON summary (src/filter_common.rs:9-37); field order = serialisation order
struct FilterSummary {
    std::string version, index, input;
    std::optional<std::string> input2;
    std::string output;
    std::optional<std::string> output2;
    uint8_t k = 0, w = 0;
    uint64_t abs_threshold = 0;
    double rel_threshold = 0;
    uint64_t prefix_length = 0;
    bool deplete = false, rename = false;
    uint64_t seqs_in = 0, seqs_out = 0;
    double seqs_out_proportion = 0;
    uint64_t seqs_removed = 0;
    double seqs_removed_proportion = 0;
    uint64_t bp_in = 0, bp_out = 0;
    double bp_out_proportion = 0;
    uint64_t bp_removed = 0;
    double bp_removed_proportion = 0;
    double time = 0;
    uint64_t seqs_per_second = 0, bp_per_second = 0;
    std::string to_json() const;   // serde_json::to_writer_pretty layout
};

// src/lib.rs:39-87 (defaults: FilterConfig::new, src/lib.rs:89-110)
struct FilterConfig {
    std::string minimizers_path;
    std::string input_path = "-";
    std::optional<std::string> input2_path;
    std::string output_path = "-";
    std::optional<std::string> output2_path;
    uint64_t abs_threshold = 2;
    double rel_threshold = 0.01;
    uint64_t prefix_length = 0;
    std::optional<std::string> summary_path;
    bool deplete = false;
    bool rename = false;
    unsigned threads = 0;            // host threads for parsing / gathering / formatting (0 = all)
    int compression_level = 2;
    bool debug = false;
    bool quiet = false;
    std::vector<int> devices = {0};  // extension: GPUs to shard the batches over (index replicated)
    uint64_t batch_mbp = 64;         // extension: bases per GPU batch, in millions

    FilterSummary execute() const;   // filter::run (src/local_filter.rs:575)
};

// src/lib.rs:187-211 (defaults: IndexConfig::new, src/lib.rs:213-226)
struct IndexConfig {
    std::string input_path;
    uint8_t kmer_length = 31;
    uint8_t window_size = 15;
    std::optional<std::string> output_path;   // nullopt = stdout
    uint64_t capacity_millions = 400;         // accepted for compatibility; the GPU build sizes itself
    unsigned threads = 8;
    bool quiet = false;
    float entropy_threshold = 0.0f;
    int device = 0;

    void execute() const;   // index::build (src/index.rs:167-308)
};

FilterSummary run_filter(const FilterConfig &config);
void build_index(const IndexConfig &config);
void index_info(const std::string &index_path, int device = 0);                                  // src/index.rs:539-560
void union_index(const std::vector<std::string> &inputs, const std::optional<std::string> &output,
                 std::optional<uint64_t> capacity_millions, int device = 0);                       // src/index.rs:563-664
void diff_index(const std::string &first, const std::string &second, std::optional<uint8_t> kmer_length,
                std::optional<uint8_t> window_size, const std::optional<std::string> &output, int device = 0);  // src/index.rs:421-537

std::string format_duration(double seconds);   // Rust's {:.2?} of a Duration
std::string format_f64(double v);              // serde_json / Display of an f64

}  // namespace deacon
