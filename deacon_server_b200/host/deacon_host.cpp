// deacon_host.cpp -- see deacon_host.hpp.  Every decision comes from the GPU through the C ABI; nothing in
// this file hashes, looks up or classifies.
#include "deacon_host.hpp"

#include <sys/stat.h>

#include <algorithm>
#include <charconv>
#include <chrono>
#include <cstdio>
#include <deque>
#include <map>

#include "../../include/deacon_cuda.h"
#include "dcn_fastx.hpp"

namespace deacon {

namespace {

using Clock = std::chrono::steady_clock;
double seconds_since(Clock::time_point t0) { return std::chrono::duration<double>(Clock::now() - t0).count(); }

// ------------------------------------------------------------------ the GPU context
struct Gpu {
    dcn_ctx *ctx;
    explicit Gpu(int device) : ctx(dcn_ctx_create(device)) {
        if (!ctx) throw Error(std::string("Failed to create a GPU context (there is no CPU fallback): ") + dcn_last_error(nullptr));
    }
    ~Gpu() { dcn_ctx_destroy(ctx); }
    Gpu(const Gpu &) = delete;
    Gpu &operator=(const Gpu &) = delete;
    void check(int rc) const {
        if (rc != DCN_OK) throw Error(dcn_last_error(ctx));
    }
};

std::vector<uint8_t> read_whole_file(const std::string &path, const char *what) {
    struct stat st;
    if (stat(path.c_str(), &st) != 0) throw Error(std::string("Failed to open ") + what + " \"" + path + "\"");
    FdSource src(path);
    std::vector<uint8_t> data((size_t)st.st_size);
    size_t got = 0;
    while (got < data.size()) {
        size_t r = src.read(reinterpret_cast<char *>(data.data()) + got, data.size() - got);
        if (!r) break;
        got += r;
    }
    data.resize(got);
    return data;
}

struct IdxInfo {
    uint8_t version = 0, k = 0, w = 0;
    uint64_t n_in_file = 0, n_set = 0;
};
// load_minimizer_hashes (src/index.rs:80-107): the file goes to the GPU as bytes and is decoded there
IdxInfo decode_idx_file(const Gpu &gpu, const std::string &path, int mode, bool make_resident) {
    const std::vector<uint8_t> file = read_whole_file(path, "index file");
    IdxInfo info;
    gpu.check(dcn_idx_decode(gpu.ctx, file.data(), file.size(), mode, make_resident ? 1 : 0, &info.version, &info.k, &info.w,
                             &info.n_in_file, &info.n_set));
    return info;
}

// write_minimizers (src/index.rs:130-164): the working key set of the ctx, encoded on the GPU
void write_working_set(const Gpu &gpu, const std::optional<std::string> &output) {
    uint64_t len = 0;
    int rc = dcn_idx_encode(gpu.ctx, nullptr, 0, &len);
    if (rc != DCN_OK && rc != DCN_ERR_OVERFLOW) gpu.check(rc);
    std::vector<uint8_t> out((size_t)len);
    gpu.check(dcn_idx_encode(gpu.ctx, out.data(), out.size(), &len));
    FdSink sink(output && *output != "-" ? *output : std::string("-"));
    sink.write(reinterpret_cast<const char *>(out.data()), (size_t)len);
    sink.finish();
}

unsigned host_threads(unsigned requested) {
    if (requested) return requested;
    unsigned n = std::thread::hardware_concurrency();
    return n ? n : 1;
}

// ------------------------------------------------------------------ bounded channel between pipeline stages
template <class T>
class Channel {
  public:
    explicit Channel(size_t cap) : cap_(cap) {}
    bool push(T v) {
        std::unique_lock<std::mutex> g(m_);
        not_full_.wait(g, [&] { return q_.size() < cap_ || closed_; });
        if (closed_) return false;
        q_.push_back(std::move(v));
        not_empty_.notify_one();
        return true;
    }
    bool pop(T &v) {
        std::unique_lock<std::mutex> g(m_);
        not_empty_.wait(g, [&] { return !q_.empty() || done_ || closed_; });
        if (closed_ || q_.empty()) return false;
        v = std::move(q_.front());
        q_.pop_front();
        not_full_.notify_one();
        return true;
    }
    void done() {   // the producer has nothing more: consumers drain what is queued
        std::lock_guard<std::mutex> g(m_);
        done_ = true;
        not_empty_.notify_all();
    }
    void close() {  // abort: wake everyone, drop what is queued
        std::lock_guard<std::mutex> g(m_);
        closed_ = true;
        not_empty_.notify_all();
        not_full_.notify_all();
    }

  private:
    size_t cap_;
    std::mutex m_;
    std::condition_variable not_full_, not_empty_;
    std::deque<T> q_;
    bool done_ = false, closed_ = false;
};

// ------------------------------------------------------------------ batches
struct PinSlot {   // pinned staging the GPU reads a batch from: gathered bases + record offsets
    char *bases = nullptr;
    uint64_t *off = nullptr;
    size_t bases_cap = 0, off_cap = 0;
    ~PinSlot() {
        if (bases) dcn_host_free(bases);
        if (off) dcn_host_free(off);
    }
    void ensure(size_t n_bases, size_t n_rec) {
        if (n_bases > bases_cap) {
            if (bases) dcn_host_free(bases);
            bases_cap = std::max<size_t>(n_bases + n_bases / 8, 1u << 20);
            bases = static_cast<char *>(dcn_host_alloc(bases_cap));
            if (!bases) { bases_cap = 0; throw Error("Failed to allocate pinned host memory for a batch"); }
        }
        if (n_rec + 1 > off_cap) {
            if (off) dcn_host_free(off);
            off_cap = std::max<size_t>(n_rec + 1 + n_rec / 8, 1u << 16);
            off = static_cast<uint64_t *>(dcn_host_alloc(off_cap * sizeof(uint64_t)));
            if (!off) { off_cap = 0; throw Error("Failed to allocate pinned host memory for a batch"); }
        }
    }
};

struct Batch {
    uint64_t seq_no = 0;
    std::vector<std::shared_ptr<Chunk>> chunks;   // keep the record views alive
    std::vector<const Rec *> recs;                // unit order; paired: records 2i, 2i + 1 are mates
    std::vector<uint8_t> fastq;                   // per record: quality present
    uint64_t n_bases = 0;
    PinSlot *pin = nullptr;
    std::vector<uint8_t> keep;
    std::vector<uint32_t> hits, total;
    // --debug, single-end: minimizer CSR + which of them are counted hits (src/local_filter.rs:351-363)
    std::vector<uint64_t> dbg_off;
    std::vector<uint32_t> dbg_pos;
    std::vector<uint8_t> dbg_flag;
};

struct Stats {   // ProcessingStats (src/local_filter.rs:179-187)
    uint64_t total_seqs = 0, filtered_seqs = 0, total_bp = 0, output_bp = 0, filtered_bp = 0, output_seq_counter = 0;
};

// format_record_to_buffer (src/local_filter.rs:60-92), split into "how many bytes" and "write them" so that the kept
// records of a batch can be laid out by a prefix sum and written by several threads
unsigned decimal_digits(uint64_t v) {
    unsigned d = 1;
    while (v >= 10) { v /= 10; d++; }
    return d;
}
size_t record_size(const Rec &r, bool fastq, uint64_t counter, bool rename) {
    if (!rename && r.verbatim) return r.raw_len;
    const size_t id = rename ? decimal_digits(counter) : r.id_len;
    return 1 + id + 1 + (size_t)r.seq_len + (fastq ? 3 + (size_t)r.seq_len : 0) + 1;
}
char *put_record(char *dst, const Rec &r, bool fastq, const char *seq, uint64_t counter, bool rename) {
    if (!rename && r.verbatim) { memcpy(dst, r.raw, r.raw_len); return dst + r.raw_len; }
    *dst++ = fastq ? '@' : '>';
    if (rename) {
        const unsigned d = decimal_digits(counter);
        for (unsigned i = d; i-- > 0; counter /= 10) dst[i] = (char)('0' + counter % 10);
        dst += d;
    } else {
        memcpy(dst, r.id, r.id_len);
        dst += r.id_len;
    }
    *dst++ = '\n';
    memcpy(dst, seq, r.seq_len);
    dst += r.seq_len;
    if (fastq) {
        memcpy(dst, "\n+\n", 3);
        dst += 3;
        memcpy(dst, r.qual, r.seq_len);
        dst += r.seq_len;
    }
    *dst++ = '\n';
    return dst;
}
void append_record(std::string &out, const Rec &r, bool fastq, const char *seq, uint64_t counter, bool rename) {
    const size_t at = out.size();
    out.resize(at + record_size(r, fastq, counter, rename));
    put_record(&out[at], r, fastq, seq, counter, rename);
}

}  // namespace

// ------------------------------------------------------------------ formatting helpers
std::string format_duration(double s) {
    char buf[64];
    if (s >= 1.0) snprintf(buf, sizeof(buf), "%.2fs", s);
    else if (s >= 1e-3) snprintf(buf, sizeof(buf), "%.2fms", s * 1e3);
    else if (s >= 1e-6) snprintf(buf, sizeof(buf), "%.2f\xC2\xB5s", s * 1e6);
    else snprintf(buf, sizeof(buf), "%.2fns", s * 1e9);
    return buf;
}

std::string format_f64(double v) {
    char buf[64];
    auto res = std::to_chars(buf, buf + sizeof(buf), v);   // shortest round-trip form, like ryu
    std::string s(buf, res.ptr);
    const size_t e = s.find('e');
    if (e != std::string::npos) {   // 1e-07 -> 1e-7
        std::string mant = s.substr(0, e), ex = s.substr(e + 1);
        bool neg = !ex.empty() && ex[0] == '-';
        if (!ex.empty() && (ex[0] == '-' || ex[0] == '+')) ex.erase(0, 1);
        while (ex.size() > 1 && ex[0] == '0') ex.erase(0, 1);
        return mant + "e" + (neg ? "-" : "") + ex;
    }
    if (s.find('.') == std::string::npos && s.find("inf") == std::string::npos && s.find("nan") == std::string::npos) s += ".0";
    return s;
}

static std::string json_string(const std::string &s) {
    std::string o = "\"";
    for (unsigned char c : s) {
        switch (c) {
            case '"': o += "\\\""; break;
            case '\\': o += "\\\\"; break;
            case '\n': o += "\\n"; break;
            case '\r': o += "\\r"; break;
            case '\t': o += "\\t"; break;
            case '\b': o += "\\b"; break;
            case '\f': o += "\\f"; break;
            default:
                if (c < 0x20) { char b[8]; snprintf(b, sizeof(b), "\\u%04x", c); o += b; }
                else o.push_back((char)c);
        }
    }
    return o + "\"";
}

std::string FilterSummary::to_json() const {
    std::string o = "{\n";
    bool first = true;
    auto field = [&](const char *k, const std::string &v) {
        if (!first) o += ",\n";
        first = false;
        o += "  \"" + std::string(k) + "\": " + v;
    };
    auto opt = [&](const std::optional<std::string> &v) { return v ? json_string(*v) : std::string("null"); };
    field("version", json_string(version));
    field("index", json_string(index));
    field("input", json_string(input));
    field("input2", opt(input2));
    field("output", json_string(output));
    field("output2", opt(output2));
    field("k", std::to_string(k));
    field("w", std::to_string(w));
    field("abs_threshold", std::to_string(abs_threshold));
    field("rel_threshold", format_f64(rel_threshold));
    field("prefix_length", std::to_string(prefix_length));
    field("deplete", deplete ? "true" : "false");
    field("rename", rename ? "true" : "false");
    field("seqs_in", std::to_string(seqs_in));
    field("seqs_out", std::to_string(seqs_out));
    field("seqs_out_proportion", format_f64(seqs_out_proportion));
    field("seqs_removed", std::to_string(seqs_removed));
    field("seqs_removed_proportion", format_f64(seqs_removed_proportion));
    field("bp_in", std::to_string(bp_in));
    field("bp_out", std::to_string(bp_out));
    field("bp_out_proportion", format_f64(bp_out_proportion));
    field("bp_removed", std::to_string(bp_removed));
    field("bp_removed_proportion", format_f64(bp_removed_proportion));
    field("time", format_f64(time));
    field("seqs_per_second", std::to_string(seqs_per_second));
    field("bp_per_second", std::to_string(bp_per_second));
    return o + "\n}";
}

// ------------------------------------------------------------------ filter::run (src/local_filter.rs:575-810)
FilterSummary run_filter(const FilterConfig &cfg) {
    const auto start_time = Clock::now();
    const bool quiet = cfg.quiet || cfg.debug;   // src/local_filter.rs:581
    const bool paired_stdin = cfg.input_path == "-" && cfg.input2_path && *cfg.input2_path == "-";
    const bool paired = cfg.input2_path.has_value();
    if (cfg.devices.empty()) throw Error("No GPU selected");
    if (cfg.abs_threshold == 0 || cfg.abs_threshold > 0xFFFFull) throw Error("abs_threshold must be in 1..=65535");
    if (cfg.prefix_length > 0xFFFFFFFFull) throw Error("prefix_length does not fit 32 bits");

    if (!quiet) {
        std::string options = "abs_threshold=" + std::to_string(cfg.abs_threshold) + ", rel_threshold=" + format_f64(cfg.rel_threshold);
        if (cfg.prefix_length > 0) options += ", prefix_length=" + std::to_string(cfg.prefix_length);
        if (cfg.rename) options += ", rename";
        if (cfg.threads > 0) options += ", threads=" + std::to_string(cfg.threads);
        fprintf(stderr, "Deacon v%s; mode: %s; input: %s; options: %s\n", VERSION, cfg.deplete ? "deplete" : "search",
                paired_stdin ? "interleaved" : paired ? "paired" : "single", options.c_str());
    }

    // Batches of 64 Mbp by default: large enough for full PCIe / kernel rate, small enough that pinning the three
    // staging buffers (a few hundred MB at ~2 GB/s) does not show in the start-up time.
    const uint64_t target_bases = std::max<uint64_t>(1, cfg.batch_mbp) * 1000000ull;
    const size_t n_slots = cfg.devices.size() + 2;
    std::vector<PinSlot> slots(n_slots);

    // one context per GPU, the index replicated in each
    std::vector<std::unique_ptr<Gpu>> gpus;
    IdxInfo idx;
    {
        const std::vector<uint8_t> file = read_whole_file(cfg.minimizers_path, "index file");
        for (int d : cfg.devices) {
            gpus.emplace_back(new Gpu(d));
            const Gpu &g = *gpus.back();
            g.check(dcn_idx_decode(g.ctx, file.data(), file.size(), DCN_SET_REPLACE, 1, &idx.version, &idx.k, &idx.w, &idx.n_in_file,
                                   &idx.n_set));
        }
    }
    if (!quiet) fprintf(stderr, "Loaded index (k=%u, w=%u) in %s\n", idx.k, idx.w, format_duration(seconds_since(start_time)).c_str());

    std::unique_ptr<Sink> writer = get_writer(cfg.output_path, cfg.compression_level);
    std::unique_ptr<Sink> writer2;
    if (cfg.output2_path && cfg.input2_path) writer2 = get_writer(*cfg.output2_path, cfg.compression_level);

    // one fork-join pool per stage that uses one, so that parsing and gathering overlap instead of taking turns
    const int T = (int)host_threads(cfg.threads);
    Pool pool(std::max(1, T / 2)), read_pool1(std::max(1, T / 2)), read_pool2(paired && !paired_stdin ? std::max(1, T / 2) : 1);
    Pool write_pool(cfg.debug ? 1 : std::max(1, T / 2));
    Channel<PinSlot *> free_slots(n_slots);
    for (auto &s : slots) free_slots.push(&s);
    Channel<std::shared_ptr<Chunk>> q_in1(3), q_in2(3);
    Channel<std::unique_ptr<Batch>> q_gpu(gpus.size() + 1), q_out(gpus.size() + 2);

    std::mutex err_m;
    std::exception_ptr first_error;
    auto abort_all = [&] {
        q_in1.close(); q_in2.close(); q_gpu.close(); q_out.close(); free_slots.close();
    };
    auto guarded = [&](auto fn) {
        return [&, fn] {
            try {
                fn();
            } catch (...) {
                {
                    std::lock_guard<std::mutex> g(err_m);
                    if (!first_error) first_error = std::current_exception();
                }
                abort_all();
            }
        };
    };

    // busy seconds per stage (DCN_HOST_TIMING=1 prints them): where a file-to-file run spends its time
    std::atomic<uint64_t> busy_ns[4] = {{0}, {0}, {0}, {0}};
    struct Busy {
        std::atomic<uint64_t> &acc;
        Clock::time_point t0 = Clock::now();
        ~Busy() { acc += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(Clock::now() - t0).count(); }
    };

    // ---- stage 1: read + parse
    auto read_stage = [&](const std::string &path, Channel<std::shared_ptr<Chunk>> &q, Pool &parse_pool) {
        FastxReader reader(path, &parse_pool);
        for (;;) {
            std::shared_ptr<Chunk> ch;
            {
                Busy t{busy_ns[0]};
                ch = reader.next();
            }
            if (!ch) break;
            if (!q.push(ch)) return;
        }
        q.done();
    };

    // ---- stage 2: assemble batches (pair the two streams, gather the bases into pinned memory)
    auto assemble_stage = [&] {
        uint64_t seq_no = 0;
        std::unique_ptr<Batch> b(new Batch);
        auto flush = [&]() -> bool {
            if (b->recs.empty()) return true;
            PinSlot *slot = nullptr;
            if (!free_slots.pop(slot)) return false;
            Busy t{busy_ns[1]};
            const size_t n = b->recs.size();
            uint64_t nb = 0;
            for (const Rec *r : b->recs) nb += r->seq_len;
            if (n > 0xFFFFFFF0ull) throw Error("batch holds too many records");
            slot->ensure((size_t)nb, n);
            uint64_t at = 0;
            for (size_t i = 0; i < n; i++) { slot->off[i] = at; at += b->recs[i]->seq_len; }
            slot->off[n] = at;
            const size_t step = 2048;
            pool.run((n + step - 1) / step, [&](size_t blk) {
                const size_t lo = blk * step, hi = std::min(n, lo + step);
                for (size_t i = lo; i < hi; i++) copy_seq(*b->recs[i], slot->bases + slot->off[i]);
            });
            b->pin = slot; b->n_bases = nb; b->seq_no = seq_no++;
            if (!q_gpu.push(std::move(b))) return false;
            b.reset(new Batch);
            return true;
        };
        auto add = [&](const std::shared_ptr<Chunk> &ch, size_t i) {
            b->recs.push_back(&ch->recs[i]);
            b->fastq.push_back(ch->fastq ? 1 : 0);
            b->n_bases += ch->recs[i].seq_len;
        };
        if (!paired) {
            std::shared_ptr<Chunk> ch;
            while (q_in1.pop(ch)) {
                b->chunks.push_back(ch);
                for (size_t i = 0; i < ch->recs.size(); i++) add(ch, i);
                if (b->n_bases >= target_bases && !flush()) return;
            }
        } else if (paired_stdin) {
            // interleaved pairs on stdin: consecutive records are mates; an odd record waits for the next block
            std::shared_ptr<Chunk> ch, held_chunk;
            size_t held = 0;
            bool have_held = false;
            while (q_in1.pop(ch)) {
                b->chunks.push_back(ch);
                size_t i = 0;
                if (have_held && !ch->recs.empty()) { add(held_chunk, held); add(ch, 0); i = 1; have_held = false; }
                for (; i + 1 < ch->recs.size(); i += 2) { add(ch, i); add(ch, i + 1); }
                if (i < ch->recs.size()) { held_chunk = ch; held = i; have_held = true; }
                if (b->n_bases >= target_bases) {
                    if (!flush()) return;
                    if (have_held) b->chunks.push_back(held_chunk);
                }
            }
            if (have_held) throw Error("Interleaved input holds an odd number of records");
        } else {
            std::shared_ptr<Chunk> c1, c2;
            size_t i1 = 0, i2 = 0;
            for (;;) {
                if (!c1 || i1 == c1->recs.size()) { c1.reset(); i1 = 0; if (q_in1.pop(c1)) b->chunks.push_back(c1); }
                if (!c2 || i2 == c2->recs.size()) { c2.reset(); i2 = 0; if (q_in2.pop(c2)) b->chunks.push_back(c2); }
                if (!c1 && !c2) break;
                if (!c1 || !c2) {
                    std::lock_guard<std::mutex> g(err_m);
                    if (first_error) return;   // a reader failed: that error is the one to report
                    throw Error("Paired inputs hold different numbers of records");
                }
                const size_t n = std::min(c1->recs.size() - i1, c2->recs.size() - i2);
                for (size_t j = 0; j < n; j++) { add(c1, i1 + j); add(c2, i2 + j); }
                i1 += n; i2 += n;
                if (b->n_bases >= target_bases) {
                    if (!flush()) return;
                    if (c1 && i1 < c1->recs.size()) b->chunks.push_back(c1);
                    if (c2 && i2 < c2->recs.size()) b->chunks.push_back(c2);
                }
            }
        }
        if (!flush()) return;
        q_gpu.done();
    };

    // ---- stage 3: the GPU (one thread per context)
    std::atomic<int> gpu_running{(int)gpus.size()};
    auto gpu_stage = [&](size_t gi) {
        const Gpu &g = *gpus[gi];
        std::unique_ptr<Batch> b;
        while (q_gpu.pop(b)) {
            const uint32_t n_rec = (uint32_t)b->recs.size();
            const uint32_t n_units = paired ? n_rec / 2 : n_rec;
            b->keep.assign(n_units, 0); b->hits.assign(n_units, 0); b->total.assign(n_units, 0);
            const uint8_t *bases = reinterpret_cast<const uint8_t *>(b->pin->bases);
            Busy t{busy_ns[2]};
            if (cfg.debug && !paired) {
                // extraction with positions (B3) + lookup with hit flags (B2): the k-mers of the DEBUG line
                uint64_t cap = b->n_bases / 4 + n_rec + 16;
                std::vector<uint64_t> hashes;
                b->dbg_off.assign((size_t)n_rec + 1, 0);
                for (int attempt = 0;; attempt++) {
                    hashes.resize((size_t)cap); b->dbg_pos.resize((size_t)cap);
                    int rc = dcn_extract(g.ctx, DCN_FLAVOUR_FILTER, bases, b->pin->off, n_rec, idx.k, idx.w, (uint32_t)cfg.prefix_length,
                                         0.0f, hashes.data(), b->dbg_pos.data(), b->dbg_off.data(), cap);
                    if (rc == DCN_ERR_OVERFLOW && attempt == 0) { cap = b->dbg_off[n_rec] + 16; continue; }
                    g.check(rc);
                    break;
                }
                b->dbg_flag.assign((size_t)b->dbg_off[n_rec] + 1, 0);
                g.check(dcn_lookup_batch_flags(g.ctx, hashes.data(), b->dbg_off.data(), n_rec, (uint32_t)cfg.abs_threshold, cfg.rel_threshold,
                                               cfg.deplete ? 1 : 0, b->keep.data(), b->hits.data(), b->total.data(), b->dbg_flag.data()));
            } else {
                g.check(dcn_filter_batch(g.ctx, bases, b->pin->off, n_rec, paired ? 1 : 0, (uint32_t)cfg.prefix_length,
                                         (uint32_t)cfg.abs_threshold, cfg.rel_threshold, cfg.deplete ? 1 : 0, b->keep.data(), b->hits.data(),
                                         b->total.data()));
            }
            if (!q_out.push(std::move(b))) return;
        }
        if (--gpu_running == 0) q_out.done();
    };

    // ---- stage 4: write the kept records in input order, count
    Stats stats;
    auto write_stage = [&] {
        std::map<uint64_t, std::unique_ptr<Batch>> pending;
        uint64_t next_no = 0;
        std::string out1, out2, dbg;
        std::unique_ptr<Batch> in;
        auto emit = [&](Batch &b) {
            Busy t{busy_ns[3]};
            const size_t rpu = paired ? 2 : 1;
            const size_t n_units = b.recs.size() / rpu;
            for (size_t u = 0; u < n_units; u++) {
                const Rec &r1 = *b.recs[u * rpu];
                const Rec *r2 = paired ? b.recs[u * rpu + 1] : nullptr;
                const uint64_t bp = (uint64_t)r1.seq_len + (r2 ? r2->seq_len : 0);
                stats.total_seqs += rpu;
                stats.total_bp += bp;
                const bool keep = b.keep[u] != 0;
                if (cfg.debug) {
                    if (!paired) {   // src/local_filter.rs:354-363: every record, with the matching k-mers
                        dbg.assign("DEBUG: ").append(r1.id, r1.id_len);
                        dbg += " hits=" + std::to_string(b.hits[u]) + "/" + std::to_string(b.total[u]) + " keep=" + (keep ? "true" : "false") + " kmers=[";
                        bool first = true;
                        for (uint64_t j = b.dbg_off[u]; j < b.dbg_off[u + 1]; j++) {
                            if (!b.dbg_flag[j]) continue;
                            if (!first) dbg.push_back(',');
                            first = false;
                            dbg.append(b.pin->bases + b.pin->off[u] + b.dbg_pos[j], idx.k);
                        }
                        dbg += "]\n";
                        fputs(dbg.c_str(), stderr);
                    } else if (b.hits[u] > 0) {   // src/local_filter.rs:424-434, 497-507; the k-mer list is always empty
                        dbg.assign("DEBUG: ").append(r1.id, r1.id_len).append("/").append(r2->id, r2->id_len);
                        dbg += " hits=" + std::to_string(b.hits[u]) + "/" + std::to_string(b.total[u]) + " keep=" + (keep ? "true" : "false") + " kmers=[]\n";
                        fputs(dbg.c_str(), stderr);
                    }
                }
                if (keep) {
                    stats.output_bp += bp;
                    const char *s1 = b.pin->bases + b.pin->off[u * rpu];
                    append_record(out1, r1, b.fastq[u * rpu] != 0, s1, ++stats.output_seq_counter, cfg.rename);
                    if (r2) {
                        const char *s2 = b.pin->bases + b.pin->off[u * rpu + 1];
                        append_record(writer2 ? out2 : out1, *r2, b.fastq[u * rpu + 1] != 0, s2, ++stats.output_seq_counter, cfg.rename);
                    }
                    if (out1.size() >= (8u << 20)) { writer->write(out1.data(), out1.size()); out1.clear(); }
                    if (out2.size() >= (8u << 20)) { writer2->write(out2.data(), out2.size()); out2.clear(); }
                } else {
                    stats.filtered_seqs += rpu;
                    stats.filtered_bp += bp;
                }
            }
            if (!out1.empty()) { writer->write(out1.data(), out1.size()); out1.clear(); }
            if (!out2.empty()) { writer2->write(out2.data(), out2.size()); out2.clear(); }
        };
        // The same without --debug, on the write pool: blocks of units are sized, laid out by a prefix sum (record
        // counter and byte offset of each block) and formatted in parallel into one buffer per output, written once.
        std::unique_ptr<char[]> buf1, buf2;
        size_t cap1 = 0, cap2 = 0;
        auto emit_parallel = [&](Batch &b) {
            Busy t{busy_ns[3]};
            const size_t rpu = paired ? 2 : 1;
            const size_t n_units = b.recs.size() / rpu;
            const size_t n_blk = std::max<size_t>(1, std::min<size_t>(n_units / 1024 + 1, (size_t)write_pool.size() * 4));
            const size_t per = (n_units + n_blk - 1) / n_blk;
            struct Blk { Stats st; uint64_t first_counter = 0; size_t bytes1 = 0, bytes2 = 0, off1 = 0, off2 = 0; };
            std::vector<Blk> blk(n_blk);
            auto walk = [&](size_t i, bool sizes, bool fill) {   // one pass over block i
                Blk &k = blk[i];
                uint64_t counter = k.first_counter;
                char *d1 = fill ? buf1.get() + k.off1 : nullptr, *d2 = fill && writer2 ? buf2.get() + k.off2 : nullptr;
                if (!fill) { k.st = Stats(); k.bytes1 = k.bytes2 = 0; }
                for (size_t u = i * per, ue = std::min(n_units, u + per); u < ue; u++) {
                    const Rec &r1 = *b.recs[u * rpu];
                    const Rec *r2 = paired ? b.recs[u * rpu + 1] : nullptr;
                    const bool keep = b.keep[u] != 0;
                    if (!fill) {
                        const uint64_t bp = (uint64_t)r1.seq_len + (r2 ? r2->seq_len : 0);
                        k.st.total_seqs += rpu; k.st.total_bp += bp;
                        if (keep) { k.st.output_bp += bp; k.st.output_seq_counter += rpu; }
                        else { k.st.filtered_seqs += rpu; k.st.filtered_bp += bp; }
                    }
                    if (!keep || !(sizes || fill)) continue;
                    const bool q1 = b.fastq[u * rpu] != 0;
                    if (fill) d1 = put_record(d1, r1, q1, b.pin->bases + b.pin->off[u * rpu], counter + 1, cfg.rename);
                    else k.bytes1 += record_size(r1, q1, counter + 1, cfg.rename);
                    if (r2) {
                        const bool q2 = b.fastq[u * rpu + 1] != 0;
                        const char *s2 = b.pin->bases + b.pin->off[u * rpu + 1];
                        if (fill) (writer2 ? d2 : d1) = put_record(writer2 ? d2 : d1, *r2, q2, s2, counter + 2, cfg.rename);
                        else (writer2 ? k.bytes2 : k.bytes1) += record_size(*r2, q2, counter + 2, cfg.rename);
                    }
                    counter += rpu;
                }
            };
            // sizes depend on the counters only when renaming: then count first, size second
            write_pool.run(n_blk, [&](size_t i) { walk(i, !cfg.rename, false); });
            uint64_t counter = stats.output_seq_counter;
            for (auto &k : blk) { k.first_counter = counter; counter += k.st.output_seq_counter; }
            if (cfg.rename) write_pool.run(n_blk, [&](size_t i) { walk(i, true, false); });
            size_t tot1 = 0, tot2 = 0;
            for (auto &k : blk) { k.off1 = tot1; k.off2 = tot2; tot1 += k.bytes1; tot2 += k.bytes2; }
            if (tot1 > cap1) { cap1 = tot1 + tot1 / 4; buf1.reset(new char[cap1]); }
            if (tot2 > cap2) { cap2 = tot2 + tot2 / 4; buf2.reset(new char[cap2]); }
            if (tot1 + tot2) write_pool.run(n_blk, [&](size_t i) { walk(i, false, true); });
            if (tot1) writer->write(buf1.get(), tot1);
            if (tot2) writer2->write(buf2.get(), tot2);
            for (auto &k : blk) {
                stats.total_seqs += k.st.total_seqs; stats.filtered_seqs += k.st.filtered_seqs; stats.total_bp += k.st.total_bp;
                stats.output_bp += k.st.output_bp; stats.filtered_bp += k.st.filtered_bp; stats.output_seq_counter += k.st.output_seq_counter;
            }
        };
        while (q_out.pop(in)) {
            pending[in->seq_no] = std::move(in);
            for (auto it = pending.find(next_no); it != pending.end(); it = pending.find(next_no)) {
                if (cfg.debug) emit(*it->second);
                else emit_parallel(*it->second);
                PinSlot *slot = it->second->pin;
                pending.erase(it);
                next_no++;
                if (!free_slots.push(slot)) return;
            }
        }
    };

    const auto filter_start = Clock::now();
    std::vector<std::thread> threads;
    threads.emplace_back(guarded([&] { read_stage(cfg.input_path, q_in1, read_pool1); }));
    if (paired && !paired_stdin) threads.emplace_back(guarded([&] { read_stage(*cfg.input2_path, q_in2, read_pool2); }));
    threads.emplace_back(guarded(assemble_stage));
    for (size_t gi = 0; gi < gpus.size(); gi++) threads.emplace_back(guarded([&, gi] { gpu_stage(gi); }));
    threads.emplace_back(guarded(write_stage));
    for (auto &t : threads) t.join();
    if (first_error) std::rethrow_exception(first_error);
    writer->finish();
    if (writer2) writer2->finish();

    // the GPU keeps the same six counters (a13); they must agree with what was written
    if (!(cfg.debug && !paired)) {
        uint64_t dev[6] = {0, 0, 0, 0, 0, 0};
        for (auto &g : gpus) {
            uint64_t c[6];
            g->check(dcn_stats_get(g->ctx, c));
            for (int i = 0; i < 6; i++) dev[i] += c[i];
        }
        if (dev[0] != stats.total_seqs || dev[1] != stats.filtered_seqs || dev[2] != stats.total_bp || dev[3] != stats.output_bp ||
            dev[4] != stats.filtered_bp || dev[5] != stats.output_seq_counter)
            throw Error("internal error: the GPU's summary counters disagree with the records written");
    }

    if (getenv("DCN_HOST_TIMING"))
        fprintf(stderr, "host stages busy: read+parse %.3fs, gather %.3fs, gpu %.3fs, write %.3fs; filter phase %.3fs (%.2f Gbp/s); wall %.3fs\n",
                busy_ns[0] * 1e-9, busy_ns[1] * 1e-9, busy_ns[2] * 1e-9, busy_ns[3] * 1e-9, seconds_since(filter_start),
                (double)stats.total_bp / seconds_since(filter_start) / 1e9, seconds_since(start_time));
    const double total_time = seconds_since(start_time);
    const double seqs_per_sec = (double)stats.total_seqs / total_time, bp_per_sec = (double)stats.total_bp / total_time;
    auto prop = [](uint64_t a, uint64_t b) { return b > 0 ? (double)a / (double)b : 0.0; };
    const uint64_t output_seqs = stats.total_seqs - stats.filtered_seqs;
    if (!quiet)
        fprintf(stderr, "Retained %llu/%llu sequences (%.3f%%), %llu/%llu bp (%.3f%%) in %s. Speed: %.0f seqs/s (%.1f Mbp/s)\n",
                (unsigned long long)output_seqs, (unsigned long long)stats.total_seqs, prop(output_seqs, stats.total_seqs) * 100.0,
                (unsigned long long)stats.output_bp, (unsigned long long)stats.total_bp, prop(stats.output_bp, stats.total_bp) * 100.0,
                format_duration(total_time).c_str(), seqs_per_sec, bp_per_sec / 1e6);

    FilterSummary s;
    s.version = std::string("deacon ") + VERSION;
    s.index = cfg.minimizers_path;
    s.input = cfg.input_path; s.input2 = cfg.input2_path;
    s.output = cfg.output_path; s.output2 = cfg.output2_path;
    s.k = idx.k; s.w = idx.w;
    s.abs_threshold = cfg.abs_threshold; s.rel_threshold = cfg.rel_threshold; s.prefix_length = cfg.prefix_length;
    s.deplete = cfg.deplete; s.rename = cfg.rename;
    s.seqs_in = stats.total_seqs; s.seqs_out = output_seqs; s.seqs_out_proportion = prop(output_seqs, stats.total_seqs);
    s.seqs_removed = stats.filtered_seqs; s.seqs_removed_proportion = prop(stats.filtered_seqs, stats.total_seqs);
    s.bp_in = stats.total_bp; s.bp_out = stats.output_bp; s.bp_out_proportion = prop(stats.output_bp, stats.total_bp);
    s.bp_removed = stats.filtered_bp; s.bp_removed_proportion = prop(stats.filtered_bp, stats.total_bp);
    s.time = total_time; s.seqs_per_second = (uint64_t)seqs_per_sec; s.bp_per_second = (uint64_t)bp_per_sec;
    if (cfg.summary_path) {
        FdSink f(*cfg.summary_path);
        const std::string js = s.to_json();
        f.write(js.data(), js.size());
        f.finish();
        if (!quiet) fprintf(stderr, "Summary saved to \"%s\"\n", cfg.summary_path->c_str());
    }
    return s;
}

FilterSummary FilterConfig::execute() const { return run_filter(*this); }

// ------------------------------------------------------------------ whole-file ingest for the index commands
namespace {
struct Sequences {
    std::vector<char> bases;
    std::vector<uint64_t> off;   // n + 1
    uint64_t n = 0;
};
// every record of a FASTA/FASTQ file as one concatenated buffer (the shape dcn_index_build takes)
Sequences read_all_sequences(const std::string &path, Pool &pool, bool list_records) {
    FastxReader reader(path, &pool);
    std::vector<std::shared_ptr<Chunk>> chunks;
    Sequences s;
    s.off.push_back(0);
    while (auto ch = reader.next()) {
        for (const Rec &r : ch->recs) {
            if (list_records) fprintf(stderr, "  %.*s (%ubp)\n", (int)r.id_len, r.id, r.seq_len);
            s.off.push_back(s.off.back() + r.seq_len);
        }
        chunks.push_back(ch);
    }
    s.n = s.off.size() - 1;
    if (s.n > 0xFFFFFFF0ull) throw Error("too many records for one index build");
    s.bases.resize((size_t)s.off.back() + 16);
    size_t base = 0;
    for (auto &ch : chunks) {
        const size_t n = ch->recs.size(), step = 256;
        pool.run((n + step - 1) / step, [&](size_t blk) {
            const size_t lo = blk * step, hi = std::min(n, lo + step);
            for (size_t i = lo; i < hi; i++) copy_seq(ch->recs[i], s.bases.data() + s.off[base + i]);
        });
        base += n;
        ch.reset();
    }
    return s;
}
}  // namespace

// ------------------------------------------------------------------ index::build (src/index.rs:167-308)
void build_index(const IndexConfig &cfg) {
    const auto start_time = Clock::now();
    std::string options = "capacity=" + std::to_string(cfg.capacity_millions) + "M";
    if (cfg.threads > 0) options += ", threads=" + std::to_string(cfg.threads);
    fprintf(stderr, "Deacon v%s; mode: build; input: single; options: %s\n", VERSION, options.c_str());
    const unsigned l = (unsigned)cfg.kmer_length + (unsigned)cfg.window_size - 1;
    if (l % 2 == 0)
        throw Error("Constraint violated: k + w - 1 must be odd (k=" + std::to_string(cfg.kmer_length) + ", w=" + std::to_string(cfg.window_size) + ")");
    Gpu gpu(cfg.device);
    Pool pool((int)host_threads(cfg.threads));
    fprintf(stderr, "Building index (k=%u, w=%u)\n", cfg.kmer_length, cfg.window_size);
    const Sequences s = read_all_sequences(cfg.input_path, pool, !cfg.quiet);
    uint64_t n_keys = 0;
    gpu.check(dcn_index_build(gpu.ctx, reinterpret_cast<const uint8_t *>(s.bases.data()), s.off.data(), (uint32_t)s.n, cfg.kmer_length,
                              cfg.window_size, cfg.entropy_threshold, 0, &n_keys));
    fprintf(stderr, "Indexed %llu minimizers from %llu sequence(s) (%llubp)\n", (unsigned long long)n_keys, (unsigned long long)s.n,
            (unsigned long long)s.off.back());
    write_working_set(gpu, cfg.output_path);
    fprintf(stderr, "Completed in %s\n", format_duration(seconds_since(start_time)).c_str());
}
void IndexConfig::execute() const { build_index(*this); }

// ------------------------------------------------------------------ index::info (src/index.rs:539-560)
void index_info(const std::string &index_path, int device) {
    const auto start_time = Clock::now();
    Gpu gpu(device);
    const IdxInfo info = decode_idx_file(gpu, index_path, DCN_SET_REPLACE, false);
    fprintf(stderr, "Index information:\n  Format version: %u\n  K-mer length (k): %u\n  Window size (w): %u\n  Distinct minimizer count: %llu\n",
            info.version, info.k, info.w, (unsigned long long)info.n_set);
    fprintf(stderr, "Retrieved index info in %s\n", format_duration(seconds_since(start_time)).c_str());
}

// ------------------------------------------------------------------ index::union (src/index.rs:563-664)
void union_index(const std::vector<std::string> &inputs, const std::optional<std::string> &output, std::optional<uint64_t>, int device) {
    const auto start_time = Clock::now();
    if (inputs.empty()) throw Error("No input files provided for union operation");
    Gpu gpu(device);
    IdxInfo info;
    for (size_t i = 0; i < inputs.size(); i++) {
        info = decode_idx_file(gpu, inputs[i], i == 0 ? DCN_SET_REPLACE : DCN_SET_UNION, false);
        if (i == 0) fprintf(stderr, "Performing union of indexes (k=%u, w=%u)\n", info.k, info.w);
        fprintf(stderr, "  %s: %llu minimizers, union now %llu\n", inputs[i].c_str(), (unsigned long long)info.n_in_file, (unsigned long long)info.n_set);
    }
    fprintf(stderr, "Union of %zu indexes: %llu distinct minimizers\n", inputs.size(), (unsigned long long)info.n_set);
    write_working_set(gpu, output);
    fprintf(stderr, "Completed union operation in %s\n", format_duration(seconds_since(start_time)).c_str());
}

// ------------------------------------------------------------------ index::diff (src/index.rs:311-537)
void diff_index(const std::string &first, const std::string &second, std::optional<uint8_t> kmer_length, std::optional<uint8_t> window_size,
                const std::optional<std::string> &output, int device) {
    const auto start_time = Clock::now();
    Gpu gpu(device);
    const IdxInfo a = decode_idx_file(gpu, first, DCN_SET_REPLACE, false);
    fprintf(stderr, "First index: loaded %llu minimizers\n", (unsigned long long)a.n_set);
    uint64_t remaining = a.n_set;
    bool as_fastx = kmer_length && window_size;
    if (!as_fastx && second != "-") {
        // an index if it decodes as one (src/index.rs:457-485); anything else is read as FASTX with the first index's k, w
        const std::vector<uint8_t> file = read_whole_file(second, "second input");
        IdxInfo b;
        const int rc = dcn_idx_decode(gpu.ctx, file.data(), file.size(), DCN_SET_SUBTRACT, 0, &b.version, &b.k, &b.w, &b.n_in_file, &b.n_set);
        if (rc == DCN_OK) {
            fprintf(stderr, "Second index: loaded %llu minimizers\n", (unsigned long long)b.n_in_file);
            remaining = b.n_set;
        } else {
            const std::string msg = dcn_last_error(gpu.ctx);
            if (msg.rfind("Incompatible headers", 0) == 0)
                throw Error("Incompatible headers: second index has k=" + std::to_string(b.k) + ", w=" + std::to_string(b.w) +
                            ", but first index has k=" + std::to_string(a.k) + ", w=" + std::to_string(a.w));
            as_fastx = true;
        }
    } else {
        as_fastx = true;
    }
    if (as_fastx) {
        const uint8_t k = kmer_length.value_or(a.k), w = window_size.value_or(a.w);
        if (k != a.k || w != a.w)
            throw Error("FASTX parameters (k=" + std::to_string(k) + ", w=" + std::to_string(w) + ") must match first index (k=" +
                        std::to_string(a.k) + ", w=" + std::to_string(a.w) + ")");
        fprintf(stderr, "Second index: processing FASTX from %s (k=%u, w=%u)\xE2\x80\xA6\n", second == "-" ? "stdin" : "file", k, w);
        Pool pool((int)host_threads(0));
        const Sequences s = read_all_sequences(second, pool, false);
        gpu.check(dcn_index_diff_sequences(gpu.ctx, reinterpret_cast<const uint8_t *>(s.bases.data()), s.off.data(), (uint32_t)s.n, &remaining));
        fprintf(stderr, "Processed %llu sequences (%llubp) from FASTX file\n", (unsigned long long)s.n, (unsigned long long)s.off.back());
    }
    fprintf(stderr, "Removed %llu minimizers, %llu remaining\n", (unsigned long long)(a.n_set - remaining), (unsigned long long)remaining);
    write_working_set(gpu, output);
    fprintf(stderr, "Completed diff operation in %s\n", format_duration(seconds_since(start_time)).c_str());
}

}  // namespace deacon
