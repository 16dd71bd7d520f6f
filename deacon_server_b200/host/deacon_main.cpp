// deacon-b200: command-line front end with the reference's sub-commands and flags (src/main.rs:9-233):
//   deacon-b200 index build|info|union|diff ...      deacon-b200 filter INDEX [INPUT] [INPUT2] ...
// All work is done by deacon_host.cpp over the C ABI; there is no CPU path.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "dcn_fastx.hpp"
#include "deacon_host.hpp"

namespace {

const char *USAGE =
    "Usage: deacon-b200 <COMMAND>\n\n"
    "Commands:\n"
    "  index   Build and compose minimizer indexes (build, info, union, diff)\n"
    "  filter  Keep or discard DNA fastx records with sufficient minimizer hits to an index\n\n"
    "Options:\n"
    "  -h, --help     Print help\n"
    "  -V, --version  Print version\n";

const char *FILTER_USAGE =
    "Usage: deacon-b200 filter [OPTIONS] <INDEX> [INPUT] [INPUT2]\n\n"
    "Arguments:\n"
    "  <INDEX>   Path to minimizer index file\n"
    "  [INPUT]   Optional path to fastx file (or - for stdin) [default: -]\n"
    "  [INPUT2]  Optional path to second paired fastx file (or - for interleaved stdin)\n\n"
    "Options:\n"
    "  -o, --output <OUTPUT>                Path to output fastx file (or - for stdout; detects .gz, .zst and .xz) [default: -]\n"
    "  -O, --output2 <OUTPUT2>              Optional path to second paired output fastx file\n"
    "  -a, --abs-threshold <N>              Minimum absolute number of minimizer hits for a match [default: 2]\n"
    "  -r, --rel-threshold <F>              Minimum relative proportion (0.0-1.0) of minimizer hits for a match [default: 0.01]\n"
    "  -p, --prefix-length <N>              Search only the first N nucleotides per sequence (0 = entire sequence) [default: 0]\n"
    "  -d, --deplete                        Discard matching sequences (invert filtering behaviour)\n"
    "  -R, --rename                         Replace sequence headers with incrementing numbers\n"
    "  -s, --summary <SUMMARY>              Path to JSON summary output file\n"
    "  -t, --threads <N>                    Number of host threads (0 = auto) [default: 8]\n"
    "      --compression-level <N>          Output compression level (1-9 for gz & xz; 1-22 for zstd) [default: 2]\n"
    "      --debug                          Output sequences with minimizer hits to stderr\n"
    "  -q, --quiet                          Suppress progress reporting\n"
    "      --devices <LIST>                 GPUs to shard the batches over, e.g. 0,1,2,3 [default: 0]\n"
    "      --batch-mbp <N>                  Bases per GPU batch, in millions [default: 64]\n";

const char *INDEX_USAGE =
    "Usage: deacon-b200 index <COMMAND>\n\n"
    "Commands:\n"
    "  build  Index minimizers contained within a fastx file\n"
    "         build [-k K] [-w W] [-o OUTPUT] [-c CAPACITY] [-t THREADS] [-q] [-e ENTROPY] <INPUT>\n"
    "  info   Show index information: info <INDEX>\n"
    "  union  Combine multiple minimizer indexes (A u B...): union [-o OUTPUT] [-c CAPACITY] <INPUTS>...\n"
    "  diff   Subtract minimizers in one index from another (A - B): diff [-k K] [-w W] [-o OUTPUT] <FIRST> <SECOND>\n";

struct UsageError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// clap-style argument splitting: -x VALUE, -xVALUE, --long VALUE, --long=VALUE, bundled short flags, "--"
struct Spec {
    std::map<std::string, std::string> alias;   // "-o" -> "output"
    std::map<std::string, bool> takes_value;    // "output" -> true
};
struct Parsed {
    std::map<std::string, std::string> opt;
    std::vector<std::string> pos;
    bool has(const std::string &k) const { return opt.count(k) != 0; }
};
Parsed parse_args(const std::vector<std::string> &args, const Spec &spec) {
    Parsed out;
    bool only_pos = false;
    for (size_t i = 0; i < args.size(); i++) {
        const std::string &a = args[i];
        if (only_pos || a == "-" || a.empty() || a[0] != '-') { out.pos.push_back(a); continue; }
        if (a == "--") { only_pos = true; continue; }
        auto take = [&](const std::string &name, const std::string *inline_value) {
            auto tv = spec.takes_value.find(name);
            if (tv == spec.takes_value.end()) throw UsageError("unexpected argument '" + a + "' found");
            if (!tv->second) {
                if (inline_value) throw UsageError("unexpected value for '--" + name + "'");
                out.opt[name] = "true";
                return;
            }
            if (inline_value) out.opt[name] = *inline_value;
            else {
                if (i + 1 >= args.size()) throw UsageError("a value is required for '--" + name + "' but none was supplied");
                out.opt[name] = args[++i];
            }
        };
        if (a.compare(0, 2, "--") == 0) {
            const size_t eq = a.find('=');
            const std::string name = a.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
            if (eq == std::string::npos) take(name, nullptr);
            else { const std::string v = a.substr(eq + 1); take(name, &v); }
        } else {
            for (size_t j = 1; j < a.size(); j++) {
                const std::string key = std::string("-") + a[j];
                auto al = spec.alias.find(key);
                if (al == spec.alias.end()) throw UsageError("unexpected argument '" + key + "' found");
                if (spec.takes_value.at(al->second)) {
                    if (j + 1 < a.size()) { const std::string v = a.substr(j + 1 + (a[j + 1] == '=' ? 1 : 0)); take(al->second, &v); }
                    else take(al->second, nullptr);
                    break;
                }
                take(al->second, nullptr);
            }
        }
    }
    return out;
}

uint64_t to_u64(const std::string &s, const char *what, uint64_t lo, uint64_t hi) {
    char *end = nullptr;
    errno = 0;
    if (s.empty() || s[0] == '-') throw UsageError(std::string("invalid value '") + s + "' for " + what);
    unsigned long long v = strtoull(s.c_str(), &end, 10);
    if (errno || *end) throw UsageError(std::string("invalid value '") + s + "' for " + what);
    if (v < lo || v > hi) throw UsageError(std::string("invalid value '") + s + "' for " + what + ": out of range " + std::to_string(lo) + ".." + std::to_string(hi));
    return v;
}
double to_f64(const std::string &s, const char *what) {
    char *end = nullptr;
    errno = 0;
    double v = strtod(s.c_str(), &end);
    if (s.empty() || errno || *end) throw UsageError(std::string("invalid value '") + s + "' for " + what);
    return v;
}

int cmd_filter(const std::vector<std::string> &args) {
    Spec spec;
    spec.alias = {{"-o", "output"}, {"-O", "output2"}, {"-a", "abs-threshold"}, {"-r", "rel-threshold"}, {"-p", "prefix-length"},
                  {"-d", "deplete"}, {"-R", "rename"}, {"-s", "summary"}, {"-t", "threads"}, {"-q", "quiet"}, {"-h", "help"}};
    spec.takes_value = {{"output", true}, {"output2", true}, {"abs-threshold", true}, {"rel-threshold", true}, {"prefix-length", true},
                        {"deplete", false}, {"rename", false}, {"summary", true}, {"threads", true}, {"compression-level", true},
                        {"debug", false}, {"quiet", false}, {"devices", true}, {"batch-mbp", true}, {"help", false}};
    const Parsed p = parse_args(args, spec);
    if (p.has("help")) { fputs(FILTER_USAGE, stdout); return 0; }
    if (p.pos.empty()) throw UsageError("the following required arguments were not provided:\n  <INDEX>");
    if (p.pos.size() > 3) throw UsageError("unexpected argument '" + p.pos[3] + "' found");
    deacon::FilterConfig c;
    c.minimizers_path = p.pos[0];
    if (p.pos.size() > 1) c.input_path = p.pos[1];
    if (p.pos.size() > 2) c.input2_path = p.pos[2];
    if (p.has("output")) c.output_path = p.opt.at("output");
    if (p.has("output2")) c.output2_path = p.opt.at("output2");
    if (p.has("abs-threshold")) c.abs_threshold = to_u64(p.opt.at("abs-threshold"), "'--abs-threshold <ABS_THRESHOLD>'", 1, 65535);
    if (p.has("rel-threshold")) c.rel_threshold = to_f64(p.opt.at("rel-threshold"), "'--rel-threshold <REL_THRESHOLD>'");
    if (p.has("prefix-length")) c.prefix_length = to_u64(p.opt.at("prefix-length"), "'--prefix-length <PREFIX_LENGTH>'", 0, 0xFFFFFFFFull);
    c.deplete = p.has("deplete");
    c.rename = p.has("rename");
    if (p.has("summary")) c.summary_path = p.opt.at("summary");
    c.threads = p.has("threads") ? (unsigned)to_u64(p.opt.at("threads"), "'--threads <THREADS>'", 0, 4096) : 8;   // src/main.rs:68
    if (p.has("compression-level")) c.compression_level = (int)to_u64(p.opt.at("compression-level"), "'--compression-level'", 0, 255);
    c.debug = p.has("debug");
    c.quiet = p.has("quiet");
    if (p.has("batch-mbp")) c.batch_mbp = to_u64(p.opt.at("batch-mbp"), "'--batch-mbp'", 1, 16384);
    if (p.has("devices")) {
        c.devices.clear();
        const std::string &s = p.opt.at("devices");
        for (size_t at = 0; at <= s.size();) {
            const size_t comma = std::min(s.find(',', at), s.size());
            c.devices.push_back((int)to_u64(s.substr(at, comma - at), "'--devices'", 0, 1023));
            at = comma + 1;
        }
    }
    c.execute();
    return 0;
}

int cmd_index(const std::vector<std::string> &args) {
    if (args.empty() || args[0] == "-h" || args[0] == "--help") {
        fputs(INDEX_USAGE, args.empty() ? stderr : stdout);
        return args.empty() ? 2 : 0;
    }
    const std::string sub = args[0];
    const std::vector<std::string> rest(args.begin() + 1, args.end());
    Spec spec;
    spec.alias = {{"-k", "kmer-length"}, {"-w", "window-size"}, {"-o", "output"}, {"-c", "capacity"}, {"-t", "threads"}, {"-q", "quiet"},
                  {"-e", "entropy-threshold"}, {"-h", "help"}};
    spec.takes_value = {{"kmer-length", true}, {"window-size", true}, {"output", true}, {"capacity", true}, {"threads", true},
                        {"quiet", false}, {"entropy-threshold", true}, {"device", true}, {"help", false}};
    const Parsed p = parse_args(rest, spec);
    if (p.has("help")) { fputs(INDEX_USAGE, stdout); return 0; }
    const int device = p.has("device") ? (int)to_u64(p.opt.at("device"), "'--device'", 0, 1023) : 0;
    std::optional<std::string> output;
    if (p.has("output") && p.opt.at("output") != "-") output = p.opt.at("output");
    if (sub == "build") {
        if (p.pos.size() != 1) throw UsageError("the following required arguments were not provided:\n  <INPUT>");
        deacon::IndexConfig c;
        c.input_path = p.pos[0];
        if (p.has("kmer-length")) c.kmer_length = (uint8_t)to_u64(p.opt.at("kmer-length"), "'-k <KMER_LENGTH>'", 1, 57);   // src/main.rs:166
        if (p.has("window-size")) c.window_size = (uint8_t)to_u64(p.opt.at("window-size"), "'-w <WINDOW_SIZE>'", 0, 255);
        c.output_path = output;
        if (p.has("capacity")) c.capacity_millions = to_u64(p.opt.at("capacity"), "'--capacity'", 0, ~0ull);
        if (p.has("threads")) c.threads = (unsigned)to_u64(p.opt.at("threads"), "'--threads'", 0, 4096);
        c.quiet = p.has("quiet");
        if (p.has("entropy-threshold")) c.entropy_threshold = (float)to_f64(p.opt.at("entropy-threshold"), "'--entropy-threshold'");
        c.device = device;
        c.execute();
    } else if (sub == "info") {
        if (p.pos.size() != 1) throw UsageError("the following required arguments were not provided:\n  <INDEX>");
        deacon::index_info(p.pos[0], device);
    } else if (sub == "union") {
        if (p.pos.empty()) throw UsageError("the following required arguments were not provided:\n  <INPUTS>...");
        std::optional<uint64_t> cap;
        if (p.has("capacity")) cap = to_u64(p.opt.at("capacity"), "'--capacity'", 0, ~0ull);
        deacon::union_index(p.pos, output, cap, device);
    } else if (sub == "diff") {
        if (p.pos.size() != 2) throw UsageError("the following required arguments were not provided:\n  <FIRST> <SECOND>");
        std::optional<uint8_t> k, w;
        if (p.has("kmer-length")) k = (uint8_t)to_u64(p.opt.at("kmer-length"), "'--kmer-length'", 1, 32);   // src/main.rs:223
        if (p.has("window-size")) w = (uint8_t)to_u64(p.opt.at("window-size"), "'--window-size'", 0, 255);
        deacon::diff_index(p.pos[0], p.pos[1], k, w, output, device);
    } else {
        throw UsageError("unrecognized subcommand '" + sub + "'");
    }
    return 0;
}

// Test hooks (no GPU needed): `_parse FILE [THREADS] [BLOCK_BYTES]` prints every record as id<TAB>seq<TAB>qual<TAB>verbatim,
// `_recode IN OUT [LEVEL]` copies IN to OUT through the codec layers (formats from IN's magic bytes and OUT's extension).
int cmd_selftest_parse(const std::vector<std::string> &args) {
    if (args.empty()) throw UsageError("_parse needs a file");
    const int threads = args.size() > 1 ? (int)to_u64(args[1], "threads", 1, 256) : 1;
    const size_t block = args.size() > 2 ? (size_t)to_u64(args[2], "block bytes", 16, 1ull << 32) : (32u << 20);
    deacon::Pool pool(threads);
    deacon::FastxReader reader(args[0], &pool, block);
    std::string line, seq;
    while (auto ch = reader.next()) {
        for (const deacon::Rec &r : ch->recs) {
            seq.resize(r.seq_len);
            deacon::copy_seq(r, seq.data());
            line.assign(r.id, r.id_len).append("\t").append(seq).append("\t");
            if (r.qual) line.append(r.qual, r.seq_len);
            line.append(r.verbatim ? "\t1\n" : "\t0\n");
            fwrite(line.data(), 1, line.size(), stdout);
        }
    }
    return 0;
}
int cmd_selftest_recode(const std::vector<std::string> &args) {
    if (args.size() < 2) throw UsageError("_recode needs IN and OUT");
    auto src = deacon::open_source(args[0]);
    auto dst = deacon::get_writer(args[1], args.size() > 2 ? (int)to_u64(args[2], "level", 0, 22) : 2);
    std::vector<char> buf(1 << 20);
    while (size_t n = src->read(buf.data(), buf.size())) dst->write(buf.data(), n);
    dst->finish();
    return 0;
}

}  // namespace

int main(int argc, char **argv) {
    std::vector<std::string> args(argv + 1, argv + argc);
    try {
        if (args.empty()) { fputs(USAGE, stderr); return 2; }
        if (args[0] == "--version" || args[0] == "-V") { printf("deacon-b200 %s\n", deacon::VERSION); return 0; }
        if (args[0] == "--help" || args[0] == "-h") { fputs(USAGE, stdout); return 0; }
        const std::vector<std::string> rest(args.begin() + 1, args.end());
        if (args[0] == "filter") return cmd_filter(rest);
        if (args[0] == "index") return cmd_index(rest);
        if (args[0] == "_parse") return cmd_selftest_parse(rest);
        if (args[0] == "_recode") return cmd_selftest_recode(rest);
        if (args[0] == "server" || args[0] == "client")
            throw UsageError("the HTTP server / client pair is not part of this build: the batch engine it wraps is the C ABI "
                             "(dcn_lookup_batch, dcn_extract; see INTEGRATION.md)");
        throw UsageError("unrecognized subcommand '" + args[0] + "'");
    } catch (const UsageError &e) {
        fprintf(stderr, "error: %s\n\n%s", e.what(), USAGE);
        return 2;
    } catch (const std::exception &e) {
        fprintf(stderr, "Error: %s\n", e.what());
        return 1;
    }
}
