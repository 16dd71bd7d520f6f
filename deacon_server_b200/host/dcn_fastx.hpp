// dcn_fastx.hpp -- host ingest for the C++ driver: byte sources (plain / gzip / zstd / xz), a block
// FASTA/FASTQ reader that parses a block on several threads, and the matching output sinks.
//
// Replaces, for the driver, what the reference gets from niffler + paraseq / needletail
// (src/local_filter.rs:41-55, src/index.rs:205-209): records come out as (id, newline-free sequence,
// quality) views into the block they were read from, so the only copy the host makes of a base is
// the gather into the pinned batch buffer the GPU reads.
#pragma once
#include <dlfcn.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace deacon {

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// ------------------------------------------------------------------ a small fork-join pool
class Pool {
  public:
    explicit Pool(int n) : n_(n < 1 ? 1 : n) {
        for (int i = 1; i < n_; i++) workers_.emplace_back([this, i] { loop(i); });
    }
    ~Pool() {
        {
            std::lock_guard<std::mutex> g(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : workers_) t.join();
    }
    int size() const { return n_; }
    // run fn(i) for i in [0, tasks) on the pool (the caller's thread takes part); rethrows the first error
    void run(size_t tasks, const std::function<void(size_t)> &fn) {
        if (tasks == 0) return;
        std::unique_lock<std::mutex> run_lock(run_m_);   // one fork-join at a time
        {
            std::lock_guard<std::mutex> g(m_);
            fn_ = &fn; tasks_ = tasks; next_.store(0); pending_ = n_ - 1; err_ = nullptr; gen_++;
        }
        cv_.notify_all();
        work();
        std::unique_lock<std::mutex> g(m_);
        done_.wait(g, [this] { return pending_ == 0; });
        fn_ = nullptr;
        if (err_) std::rethrow_exception(err_);
    }

  private:
    void work() {
        for (;;) {
            size_t i = next_.fetch_add(1);
            if (i >= tasks_) break;
            try {
                (*fn_)(i);
            } catch (...) {
                std::lock_guard<std::mutex> g(m_);
                if (!err_) err_ = std::current_exception();
            }
        }
    }
    void loop(int) {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [&] { return stop_ || gen_ != seen; });
                if (stop_) return;
                seen = gen_;
            }
            work();
            {
                std::lock_guard<std::mutex> g(m_);
                if (--pending_ == 0) done_.notify_all();
            }
        }
    }
    int n_;
    std::vector<std::thread> workers_;
    std::mutex m_, run_m_;
    std::condition_variable cv_, done_;
    const std::function<void(size_t)> *fn_ = nullptr;
    size_t tasks_ = 0;
    std::atomic<size_t> next_{0};
    int pending_ = 0;
    uint64_t gen_ = 0;
    bool stop_ = false;
    std::exception_ptr err_;
};

// ------------------------------------------------------------------ codecs loaded at run time
// zstd and xz ship as run-time libraries only in this image (no headers): the few entry points
// the driver needs are declared here and resolved with dlopen.
struct ZstdApi {
    struct InBuf { const void *src; size_t size, pos; };
    struct OutBuf { void *dst; size_t size, pos; };
    void *(*createDCtx)();
    size_t (*freeDCtx)(void *);
    size_t (*decompressStream)(void *, OutBuf *, InBuf *);
    void *(*createCCtx)();
    size_t (*freeCCtx)(void *);
    size_t (*setParameter)(void *, int, int);
    size_t (*compressStream2)(void *, OutBuf *, InBuf *, int);
    unsigned (*isError)(size_t);
    const char *(*getErrorName)(size_t);
    static const ZstdApi &get() {
        static ZstdApi api = load();
        return api;
    }

  private:
    static ZstdApi load() {
        void *h = dlopen("libzstd.so.1", RTLD_NOW);
        if (!h) throw Error("zstd support needs libzstd.so.1, which could not be loaded");
        ZstdApi a;
        auto sym = [&](const char *n) {
            void *p = dlsym(h, n);
            if (!p) throw Error(std::string("libzstd.so.1 lacks ") + n);
            return p;
        };
        a.createDCtx = reinterpret_cast<void *(*)()>(sym("ZSTD_createDCtx"));
        a.freeDCtx = reinterpret_cast<size_t (*)(void *)>(sym("ZSTD_freeDCtx"));
        a.decompressStream = reinterpret_cast<size_t (*)(void *, OutBuf *, InBuf *)>(sym("ZSTD_decompressStream"));
        a.createCCtx = reinterpret_cast<void *(*)()>(sym("ZSTD_createCCtx"));
        a.freeCCtx = reinterpret_cast<size_t (*)(void *)>(sym("ZSTD_freeCCtx"));
        a.setParameter = reinterpret_cast<size_t (*)(void *, int, int)>(sym("ZSTD_CCtx_setParameter"));
        a.compressStream2 = reinterpret_cast<size_t (*)(void *, OutBuf *, InBuf *, int)>(sym("ZSTD_compressStream2"));
        a.isError = reinterpret_cast<unsigned (*)(size_t)>(sym("ZSTD_isError"));
        a.getErrorName = reinterpret_cast<const char *(*)(size_t)>(sym("ZSTD_getErrorName"));
        return a;
    }
};

struct LzmaApi {
    struct Stream {   // lzma_stream of liblzma 5.x (stable ABI)
        const uint8_t *next_in; size_t avail_in; uint64_t total_in;
        uint8_t *next_out; size_t avail_out; uint64_t total_out;
        const void *allocator; void *internal;
        void *rp1, *rp2, *rp3, *rp4; uint64_t ri1, ri2; size_t ri3, ri4; int re1, re2;
    };
    int (*streamDecoder)(Stream *, uint64_t, uint32_t);
    int (*easyEncoder)(Stream *, uint32_t, int);
    int (*code)(Stream *, int);
    void (*end)(Stream *);
    static const LzmaApi &get() {
        static LzmaApi api = load();
        return api;
    }

  private:
    static LzmaApi load() {
        void *h = dlopen("liblzma.so.5", RTLD_NOW);
        if (!h) throw Error("xz support needs liblzma.so.5, which could not be loaded");
        LzmaApi a;
        auto sym = [&](const char *n) {
            void *p = dlsym(h, n);
            if (!p) throw Error(std::string("liblzma.so.5 lacks ") + n);
            return p;
        };
        a.streamDecoder = reinterpret_cast<int (*)(Stream *, uint64_t, uint32_t)>(sym("lzma_stream_decoder"));
        a.easyEncoder = reinterpret_cast<int (*)(Stream *, uint32_t, int)>(sym("lzma_easy_encoder"));
        a.code = reinterpret_cast<int (*)(Stream *, int)>(sym("lzma_code"));
        a.end = reinterpret_cast<void (*)(Stream *)>(sym("lzma_end"));
        return a;
    }
};

// ------------------------------------------------------------------ byte sources
class ByteSource {
  public:
    virtual ~ByteSource() = default;
    virtual size_t read(char *dst, size_t cap) = 0;   // 0 = end of stream
};

class FdSource : public ByteSource {
  public:
    explicit FdSource(const std::string &path) {
        if (path == "-") { fd_ = 0; own_ = false; }
        else {
            fd_ = ::open(path.c_str(), O_RDONLY);
            if (fd_ < 0) throw Error("Failed to open file " + path + ": " + std::strerror(errno));
            own_ = true;
#ifdef POSIX_FADV_SEQUENTIAL
            posix_fadvise(fd_, 0, 0, POSIX_FADV_SEQUENTIAL);
#endif
        }
    }
    ~FdSource() override { if (own_) ::close(fd_); }
    size_t read(char *dst, size_t cap) override {
        for (;;) {
            ssize_t r = ::read(fd_, dst, cap);
            if (r >= 0) return (size_t)r;
            if (errno != EINTR) throw Error(std::string("read failed: ") + std::strerror(errno));
        }
    }

  private:
    int fd_; bool own_;
};

// hands back the sniffed bytes first
class PrefixedSource : public ByteSource {
  public:
    PrefixedSource(std::string head, std::unique_ptr<ByteSource> rest) : head_(std::move(head)), rest_(std::move(rest)) {}
    size_t read(char *dst, size_t cap) override {
        if (at_ < head_.size()) {
            size_t n = std::min(cap, head_.size() - at_);
            memcpy(dst, head_.data() + at_, n);
            at_ += n;
            return n;
        }
        return rest_->read(dst, cap);
    }

  private:
    std::string head_; size_t at_ = 0;
    std::unique_ptr<ByteSource> rest_;
};

class DecodeSource : public ByteSource {   // shared input buffering of the three decoders
  protected:
    explicit DecodeSource(std::unique_ptr<ByteSource> in) : in_(std::move(in)), ibuf_(1 << 20) {}
    bool refill() {
        if (in_eof_) return false;
        ilen_ = in_->read(ibuf_.data(), ibuf_.size());
        ipos_ = 0;
        if (ilen_ == 0) in_eof_ = true;
        return ilen_ > 0;
    }
    std::unique_ptr<ByteSource> in_;
    std::vector<char> ibuf_;
    size_t ipos_ = 0, ilen_ = 0;
    bool in_eof_ = false;
};

class GzSource : public DecodeSource {   // multi-member streams (bgzip) included
  public:
    explicit GzSource(std::unique_ptr<ByteSource> in) : DecodeSource(std::move(in)) {
        memset(&z_, 0, sizeof(z_));
        if (inflateInit2(&z_, 16 + MAX_WBITS) != Z_OK) throw Error("zlib: inflateInit2 failed");
    }
    ~GzSource() override { inflateEnd(&z_); }
    size_t read(char *dst, size_t cap) override {
        size_t out = 0;
        while (out == 0 && !done_) {
            if (ipos_ == ilen_ && !refill()) {
                if (mid_member_) throw Error("gzip stream ended unexpectedly");
                done_ = true;
                break;
            }
            z_.next_in = reinterpret_cast<Bytef *>(ibuf_.data() + ipos_);
            z_.avail_in = (uInt)(ilen_ - ipos_);
            z_.next_out = reinterpret_cast<Bytef *>(dst);
            z_.avail_out = (uInt)std::min<size_t>(cap, 1u << 30);
            int rc = inflate(&z_, Z_NO_FLUSH);
            ipos_ = ilen_ - z_.avail_in;
            out = std::min<size_t>(cap, 1u << 30) - z_.avail_out;
            mid_member_ = true;
            if (rc == Z_STREAM_END) { inflateReset(&z_); mid_member_ = false; }
            else if (rc != Z_OK && rc != Z_BUF_ERROR) throw Error(std::string("gzip decode failed: ") + (z_.msg ? z_.msg : "?"));
        }
        return out;
    }

  private:
    z_stream z_;
    bool done_ = false, mid_member_ = false;
};

class ZstdSource : public DecodeSource {
  public:
    explicit ZstdSource(std::unique_ptr<ByteSource> in) : DecodeSource(std::move(in)), api_(ZstdApi::get()) {
        d_ = api_.createDCtx();
        if (!d_) throw Error("zstd: cannot create a decompression context");
    }
    ~ZstdSource() override { api_.freeDCtx(d_); }
    size_t read(char *dst, size_t cap) override {
        for (;;) {
            if (ipos_ == ilen_ && !refill()) return 0;
            ZstdApi::InBuf ib{ibuf_.data(), ilen_, ipos_};
            ZstdApi::OutBuf ob{dst, cap, 0};
            size_t rc = api_.decompressStream(d_, &ob, &ib);
            if (api_.isError(rc)) throw Error(std::string("zstd decode failed: ") + api_.getErrorName(rc));
            ipos_ = ib.pos;
            if (ob.pos) return ob.pos;
        }
    }

  private:
    const ZstdApi &api_;
    void *d_;
};

class XzSource : public DecodeSource {
  public:
    explicit XzSource(std::unique_ptr<ByteSource> in) : DecodeSource(std::move(in)), api_(LzmaApi::get()) {
        memset(&s_, 0, sizeof(s_));
        if (api_.streamDecoder(&s_, UINT64_MAX, 0x08 /* LZMA_CONCATENATED */) != 0) throw Error("xz: cannot create a decoder");
    }
    ~XzSource() override { api_.end(&s_); }
    size_t read(char *dst, size_t cap) override {
        while (!done_) {
            if (ipos_ == ilen_) refill();
            s_.next_in = reinterpret_cast<const uint8_t *>(ibuf_.data() + ipos_);
            s_.avail_in = ilen_ - ipos_;
            s_.next_out = reinterpret_cast<uint8_t *>(dst);
            s_.avail_out = cap;
            int rc = api_.code(&s_, in_eof_ ? 3 /* LZMA_FINISH */ : 0 /* LZMA_RUN */);
            ipos_ = ilen_ - s_.avail_in;
            size_t out = cap - s_.avail_out;
            if (rc == 1 /* LZMA_STREAM_END */) done_ = true;
            else if (rc != 0 && !(rc == 10 /* LZMA_BUF_ERROR */ && !in_eof_)) throw Error("xz decode failed (lzma_ret " + std::to_string(rc) + ")");
            if (out) return out;
        }
        return 0;
    }

  private:
    const LzmaApi &api_;
    LzmaApi::Stream s_;
    bool done_ = false;
};

// niffler::from_path: the format comes from the magic bytes, not the file name
inline std::unique_ptr<ByteSource> open_source(const std::string &path) {
    std::unique_ptr<ByteSource> raw(new FdSource(path));
    std::string head(6, '\0');
    size_t got = 0;
    while (got < head.size()) {
        size_t r = raw->read(&head[got], head.size() - got);
        if (!r) break;
        got += r;
    }
    head.resize(got);
    const unsigned char *h = reinterpret_cast<const unsigned char *>(head.data());
    std::unique_ptr<ByteSource> src(new PrefixedSource(head, std::move(raw)));
    if (got >= 2 && h[0] == 0x1f && h[1] == 0x8b) return std::unique_ptr<ByteSource>(new GzSource(std::move(src)));
    if (got >= 4 && h[0] == 0x28 && h[1] == 0xb5 && h[2] == 0x2f && h[3] == 0xfd) return std::unique_ptr<ByteSource>(new ZstdSource(std::move(src)));
    if (got >= 6 && h[0] == 0xfd && h[1] == '7' && h[2] == 'z' && h[3] == 'X' && h[4] == 'Z' && h[5] == 0) return std::unique_ptr<ByteSource>(new XzSource(std::move(src)));
    return src;
}

// ------------------------------------------------------------------ records and chunks
struct Rec {
    const char *id;      // header line without '>' / '@' (the full line: src/local_filter.rs:72)
    const char *seq;     // first sequence byte; seq_span raw bytes that may contain line breaks (multi-line FASTA)
    const char *qual;    // nullptr for FASTA
    const char *raw;     // the whole record as it stands in the file (through its last line break)
    uint32_t id_len, seq_span, seq_len, raw_len;
    bool verbatim;       // the raw bytes are exactly what format_record_to_buffer would write (src/local_filter.rs:60-92)
};

struct Chunk {
    std::shared_ptr<const void> hold;   // what the record views point into: a read buffer or the file mapping
    std::vector<Rec> recs;
    bool fastq = false;
};

// a whole uncompressed file mapped read-only: its records are parsed in place, no read() copy
struct Mapping {
    char *p = nullptr;
    size_t n = 0;
    ~Mapping() { if (p) munmap(p, n); }
};

// copy the newline-free sequence of a record (record.seq() of paraseq / needletail) to dst
inline void copy_seq(const Rec &r, char *dst) {
    if (r.seq_len == r.seq_span) { memcpy(dst, r.seq, r.seq_len); return; }
    const char *p = r.seq, *e = r.seq + r.seq_span;
    while (p < e) {
        const char *nl = static_cast<const char *>(memchr(p, '\n', (size_t)(e - p)));
        const char *le = nl ? nl : e;
        size_t n = (size_t)(le - p);
        if (n && le[-1] == '\r') n--;
        memcpy(dst, p, n);
        dst += n;
        p = nl ? nl + 1 : e;
    }
}

namespace detail {

inline const char *find_nl(const char *p, const char *end) { return static_cast<const char *>(memchr(p, '\n', (size_t)(end - p))); }
inline const char *skip_blank(const char *p, const char *end) {
    while (p < end && (*p == '\n' || *p == '\r')) p++;
    return p;
}
inline uint32_t trimmed(const char *b, const char *e) {   // line length without a trailing '\r'
    return (uint32_t)((e > b && e[-1] == '\r') ? e - b - 1 : e - b);
}

// One FASTQ record at p (four lines).  Returns the position after it, or nullptr when the record is
// incomplete in [p, end) (only complete if `eof`: the last line may lack its line break).
inline const char *parse_fastq_record(const char *p, const char *end, bool eof, Rec &r) {
    const char *l[4], *le[4];
    const char *q = p;
    for (int i = 0; i < 4; i++) {
        if (q >= end && !(eof && i == 3 && q == end)) return nullptr;
        l[i] = q;
        const char *nl = q < end ? find_nl(q, end) : nullptr;
        if (!nl) {
            if (!eof || i != 3) return nullptr;
            le[i] = end; q = end;
        } else { le[i] = nl; q = nl + 1; }
    }
    if (*l[0] != '@') throw Error("Invalid FASTQ record: header line does not start with '@'");
    if (l[2] >= le[2] || *l[2] != '+') throw Error("Invalid FASTQ record: separator line does not start with '+'");
    r.id = l[0] + 1; r.id_len = trimmed(l[0] + 1, le[0]);
    r.seq = l[1]; r.seq_len = r.seq_span = trimmed(l[1], le[1]);
    r.qual = l[3];
    if (trimmed(l[3], le[3]) != r.seq_len) throw Error("Invalid FASTQ record: sequence and quality lengths differ");
    r.raw = p; r.raw_len = (uint32_t)(q - p);
    if ((uint64_t)(q - p) > 0xFFFFFFFFull) throw Error("FASTQ record longer than 4 GiB");
    // verbatim <=> "@id\nseq\n+\nqual\n" exactly: bare '+' line, no '\r', final line break present
    r.verbatim = (le[2] - l[2] == 1) && le[3] < end && r.id_len == (uint32_t)(le[0] - l[0] - 1) &&
                 r.seq_len == (uint32_t)(le[1] - l[1]) && r.seq_len == (uint32_t)(le[3] - l[3]);
    return q;
}

// One FASTA record at p.  Needs the start of the next record (or eof) to know where it ends.
inline const char *parse_fasta_record(const char *p, const char *end, bool eof, Rec &r) {
    if (*p != '>') throw Error("Invalid FASTA record: header line does not start with '>'");
    const char *nl = find_nl(p, end);
    if (!nl) {
        if (!eof) return nullptr;
        nl = end;   // header only
    }
    const char *s = nl < end ? nl + 1 : end;
    const char *q = s, *next = nullptr;
    while (q < end) {   // next line that starts with '>'
        if (*q == '>') { next = q; break; }
        const char *n2 = find_nl(q, end);
        if (!n2) break;
        q = n2 + 1;
    }
    if (!next) {
        if (!eof) return nullptr;
        next = end;
    }
    const char *se = next;
    while (se > s && (se[-1] == '\n' || se[-1] == '\r')) se--;
    if ((uint64_t)(next - p) > 0xFFFFFFFFull) throw Error("FASTA record longer than 4 GiB");
    uint64_t breaks = 0;
    for (const char *c = s; c < se;) {
        const char *n2 = find_nl(c, se);
        if (!n2) break;
        breaks += 1 + ((n2 > c && n2[-1] == '\r') ? 1 : 0);
        c = n2 + 1;
    }
    r.id = p + 1; r.id_len = trimmed(p + 1, nl);
    r.seq = s; r.seq_span = (uint32_t)(se - s); r.seq_len = (uint32_t)(se - s - breaks);
    r.qual = nullptr;
    r.raw = p; r.raw_len = (uint32_t)(next - p);
    r.verbatim = breaks == 0 && r.id_len == (uint32_t)(nl - p - 1) && se < end && *se == '\n' && se + 1 == next;
    return next;
}

// first record start at or after `from` (a position inside the block), for splitting a block among threads
inline const char *sync_fastq(const char *buf, const char *from, const char *end) {
    const char *q = from;
    if (q > buf) {   // move to a line start
        const char *nl = find_nl(q - 1, end);
        if (!nl) return end;
        q = nl + 1;
    }
    while (q < end) {
        const char *l1 = find_nl(q, end);
        if (!l1) return end;
        if (*q == '@') {   // a header iff the line after next starts with '+' (a quality line that starts with '@' is followed by a header and a sequence)
            const char *l2 = find_nl(l1 + 1, end);
            if (!l2) return end;
            if (l2 + 1 < end && l2[1] == '+') return q;
        }
        q = l1 + 1;
    }
    return end;
}
inline const char *sync_fasta(const char *buf, const char *from, const char *end) {
    const char *q = from;
    if (q > buf) {
        const char *nl = find_nl(q - 1, end);
        if (!nl) return end;
        q = nl + 1;
    }
    while (q < end) {
        if (*q == '>') return q;
        const char *nl = find_nl(q, end);
        if (!nl) return end;
        q = nl + 1;
    }
    return end;
}

}  // namespace detail

// ------------------------------------------------------------------ block reader
class FastxReader {
  public:
    FastxReader(const std::string &path, Pool *pool, size_t block_bytes = 32u << 20) : pool_(pool), block_(block_bytes), path_(path) {
        if (path != "-") {   // a regular uncompressed file is mapped; pipes and compressed files are streamed
            int fd = ::open(path.c_str(), O_RDONLY);
            if (fd < 0) throw Error("Failed to open file " + path + ": " + std::strerror(errno));
            struct stat st;
            if (fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0) {
                void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
                if (m != MAP_FAILED) {
                    const unsigned char *h = static_cast<const unsigned char *>(m);
                    const size_t n = (size_t)st.st_size;
                    const bool packed = (n >= 2 && h[0] == 0x1f && h[1] == 0x8b) || (n >= 4 && h[0] == 0x28 && h[1] == 0xb5 && h[2] == 0x2f && h[3] == 0xfd) ||
                                        (n >= 6 && h[0] == 0xfd && h[1] == '7' && h[2] == 'z' && h[3] == 'X' && h[4] == 'Z' && h[5] == 0);
                    if (packed) munmap(m, n);
                    else {
                        map_ = std::make_shared<Mapping>();
                        map_->p = static_cast<char *>(m); map_->n = n;
                        madvise(m, n, MADV_SEQUENTIAL);
                    }
                }
            }
            ::close(fd);
        }
        if (!map_) src_ = open_source(path);
    }

    // Next block of whole records; nullptr at the end of the input.
    std::shared_ptr<Chunk> next() { return map_ ? next_mapped() : next_streamed(); }

    bool is_fastq() const { return fastq_; }

  private:
    void sniff(const char *p) {
        if (format_known_) return;
        if (*p == '@') fastq_ = true;
        else if (*p == '>') fastq_ = false;
        else throw Error("Failed to create reader for " + path_ + ": input is neither FASTA nor FASTQ");
        format_known_ = true;
    }
    std::shared_ptr<Chunk> next_mapped() {
        const char *base = map_->p, *file_end = base + map_->n;
        size_t want = block_;
        while (!finished_) {
            const char *p = detail::skip_blank(base + pos_, file_end);
            if (p == file_end) { finished_ = true; return nullptr; }
            sniff(p);
            const char *e = (size_t)(file_end - p) > want ? p + want : file_end;
            eof_ = e == file_end;
            auto ch = std::make_shared<Chunk>();
            ch->hold = map_;
            ch->fastq = fastq_;
            const char *consumed = parse_block(p, e, *ch);
            if (eof_) {
                if (detail::skip_blank(consumed, e) != e) throw Error("Truncated record at the end of " + path_);
                finished_ = true;
            }
            pos_ = (size_t)(consumed - base);
            if (!ch->recs.empty()) return ch;
            want *= 2;   // no complete record in the block (one very long sequence): look further
        }
        return nullptr;
    }
    std::shared_ptr<Chunk> next_streamed() {
        while (!finished_) {
            size_t cap = std::max(block_, carry_.size() * 2);
            std::shared_ptr<char> buf(new char[cap + 1], std::default_delete<char[]>());
            size_t len = carry_.size();
            if (len) memcpy(buf.get(), carry_.data(), len);
            carry_.clear();
            while (len < cap && !eof_) {
                size_t r = src_->read(buf.get() + len, cap - len);
                if (r == 0) eof_ = true;
                len += r;
            }
            auto ch = std::make_shared<Chunk>();
            ch->hold = buf;
            const char *b = buf.get(), *e = b + len;
            const char *p = detail::skip_blank(b, e);
            if (p == e) {
                if (eof_) { finished_ = true; return nullptr; }
                continue;
            }
            sniff(p);
            ch->fastq = fastq_;
            const char *consumed = parse_block(p, e, *ch);
            if (eof_) {
                if (detail::skip_blank(consumed, e) != e) throw Error("Truncated record at the end of " + path_);
                finished_ = true;
            } else {
                carry_.assign(consumed, e);
            }
            if (!ch->recs.empty()) return ch;
            // no complete record in a full block (one very long sequence): read on with a larger buffer
        }
        return nullptr;
    }
    const char *parse_range(const char *p, const char *stop, const char *end, bool eof, std::vector<Rec> &out) {
        while (p < stop) {
            Rec r;
            const char *q = fastq_ ? detail::parse_fastq_record(p, end, eof, r) : detail::parse_fasta_record(p, end, eof, r);
            if (!q) break;
            out.push_back(r);
            p = detail::skip_blank(q, end);
        }
        return p;
    }
    const char *parse_block(const char *p, const char *e, Chunk &ch) {
        const int T = pool_ ? pool_->size() : 1;
        const size_t n = (size_t)(e - p);
        if (T > 1 && n >= (size_t)T * (256u << 10)) {
            std::vector<const char *> start((size_t)T + 1);
            start[0] = p; start[(size_t)T] = e;
            pool_->run((size_t)T - 1, [&](size_t i) {
                const char *from = p + n * (i + 1) / (size_t)T;
                start[i + 1] = fastq_ ? detail::sync_fastq(p, from, e) : detail::sync_fasta(p, from, e);
            });
            for (int i = 1; i <= T; i++) if (start[(size_t)i] < start[(size_t)i - 1]) start[(size_t)i] = start[(size_t)i - 1];
            std::vector<std::vector<Rec>> parts((size_t)T);
            std::vector<const char *> reached((size_t)T);
            pool_->run((size_t)T, [&](size_t i) {
                // the last range may end in an incomplete record; the others end at a synced record start, which closes
                // their final record the way the end of the file would (a FASTA record needs the next '>' or the end to
                // know where it stops: without this every non-last range lost its last record and the block was parsed twice)
                const bool last = start[i + 1] == e;
                reached[i] = start[i] < start[i + 1] ? parse_range(start[i], start[i + 1], last ? e : start[i + 1], last ? eof_ : true, parts[i]) : start[i];
            });
            bool ok = true;
            const char *consumed = p;
            for (int i = 0; i < T && ok; i++) {
                if (start[(size_t)i] == start[(size_t)i + 1]) continue;
                if (start[(size_t)i + 1] != e && reached[(size_t)i] != start[(size_t)i + 1]) ok = false;
                consumed = reached[(size_t)i];
            }
            if (ok) {
                size_t total = 0;
                for (auto &v : parts) total += v.size();
                ch.recs.reserve(total);
                for (auto &v : parts) ch.recs.insert(ch.recs.end(), v.begin(), v.end());
                return consumed;
            }
            // the split guessed a boundary wrong (odd input): parse the block sequentially
        }
        return parse_range(p, e, e, eof_, ch.recs);
    }

    std::unique_ptr<ByteSource> src_;
    std::shared_ptr<Mapping> map_;
    size_t pos_ = 0;
    Pool *pool_;
    size_t block_;
    std::string path_;
    std::vector<char> carry_;
    bool eof_ = false, finished_ = false, format_known_ = false, fastq_ = false;
};

// ------------------------------------------------------------------ output sinks (get_writer, src/local_filter.rs:109-151)
class Sink {
  public:
    virtual ~Sink() = default;
    virtual void write(const char *p, size_t n) = 0;
    virtual void finish() = 0;
};

class FdSink : public Sink {
  public:
    explicit FdSink(const std::string &path) {
        if (path == "-") { fd_ = 1; own_ = false; }
        else {
            fd_ = ::open(path.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
            if (fd_ < 0) throw Error("Failed to create output file: " + path);
            own_ = true;
        }
    }
    ~FdSink() override { if (own_ && fd_ >= 0) ::close(fd_); }
    void write(const char *p, size_t n) override {
        while (n) {
            ssize_t w = ::write(fd_, p, n);
            if (w < 0) {
                if (errno == EINTR) continue;
                throw Error(std::string("write failed: ") + std::strerror(errno));
            }
            p += w; n -= (size_t)w;
        }
    }
    void finish() override {
        if (own_ && fd_ >= 0) { ::close(fd_); fd_ = -1; }
    }

  private:
    int fd_; bool own_;
};

class GzSink : public Sink {
  public:
    GzSink(std::unique_ptr<Sink> out, int level) : out_(std::move(out)), obuf_(1 << 20) {
        memset(&z_, 0, sizeof(z_));
        if (deflateInit2(&z_, level, Z_DEFLATED, 16 + MAX_WBITS, 8, Z_DEFAULT_STRATEGY) != Z_OK) throw Error("zlib: deflateInit2 failed");
    }
    ~GzSink() override { deflateEnd(&z_); }
    void write(const char *p, size_t n) override { pump(p, n, Z_NO_FLUSH); }
    void finish() override { pump(nullptr, 0, Z_FINISH); out_->finish(); }

  private:
    void pump(const char *p, size_t n, int flush) {
        z_.next_in = reinterpret_cast<Bytef *>(const_cast<char *>(p));
        z_.avail_in = (uInt)n;
        int rc;
        do {
            z_.next_out = reinterpret_cast<Bytef *>(obuf_.data());
            z_.avail_out = (uInt)obuf_.size();
            rc = deflate(&z_, flush);
            if (rc == Z_STREAM_ERROR) throw Error("gzip encode failed");
            out_->write(obuf_.data(), obuf_.size() - z_.avail_out);
        } while (z_.avail_out == 0 || (flush == Z_FINISH && rc != Z_STREAM_END));
    }
    std::unique_ptr<Sink> out_;
    std::vector<char> obuf_;
    z_stream z_;
};

class ZstdSink : public Sink {
  public:
    ZstdSink(std::unique_ptr<Sink> out, int level) : out_(std::move(out)), obuf_(1 << 20), api_(ZstdApi::get()) {
        c_ = api_.createCCtx();
        if (!c_) throw Error("zstd: cannot create a compression context");
        api_.setParameter(c_, 100 /* ZSTD_c_compressionLevel */, level);
    }
    ~ZstdSink() override { api_.freeCCtx(c_); }
    void write(const char *p, size_t n) override { pump(p, n, 0); }
    void finish() override { pump(nullptr, 0, 2 /* ZSTD_e_end */); out_->finish(); }

  private:
    void pump(const char *p, size_t n, int op) {
        ZstdApi::InBuf ib{p, n, 0};
        size_t rem;
        do {
            ZstdApi::OutBuf ob{obuf_.data(), obuf_.size(), 0};
            rem = api_.compressStream2(c_, &ob, &ib, op);
            if (api_.isError(rem)) throw Error(std::string("zstd encode failed: ") + api_.getErrorName(rem));
            out_->write(obuf_.data(), ob.pos);
        } while (ib.pos < ib.size || (op == 2 && rem != 0));
    }
    std::unique_ptr<Sink> out_;
    std::vector<char> obuf_;
    const ZstdApi &api_;
    void *c_;
};

class XzSink : public Sink {
  public:
    XzSink(std::unique_ptr<Sink> out, int level) : out_(std::move(out)), obuf_(1 << 20), api_(LzmaApi::get()) {
        memset(&s_, 0, sizeof(s_));
        if (api_.easyEncoder(&s_, (uint32_t)level, 4 /* LZMA_CHECK_CRC64 */) != 0) throw Error("xz: cannot create an encoder");
    }
    ~XzSink() override { api_.end(&s_); }
    void write(const char *p, size_t n) override { pump(p, n, 0); }
    void finish() override { pump(nullptr, 0, 3 /* LZMA_FINISH */); out_->finish(); }

  private:
    void pump(const char *p, size_t n, int action) {
        s_.next_in = reinterpret_cast<const uint8_t *>(p);
        s_.avail_in = n;
        int rc;
        do {
            s_.next_out = reinterpret_cast<uint8_t *>(obuf_.data());
            s_.avail_out = obuf_.size();
            rc = api_.code(&s_, action);
            if (rc != 0 && rc != 1) throw Error("xz encode failed (lzma_ret " + std::to_string(rc) + ")");
            out_->write(obuf_.data(), obuf_.size() - s_.avail_out);
        } while (s_.avail_in > 0 || (action == 3 && rc != 1));
    }
    std::unique_ptr<Sink> out_;
    std::vector<char> obuf_;
    const LzmaApi &api_;
    LzmaApi::Stream s_;
};

inline bool ends_with(const std::string &s, const char *suffix) {
    size_t n = strlen(suffix);
    return s.size() >= n && s.compare(s.size() - n, n, suffix) == 0;
}

inline void validate_compression_level(int level, int lo, int hi, const char *format) {
    if (level < lo || level > hi)
        throw Error("Invalid " + std::string(format) + " compression level " + std::to_string(level) + ". Must be between " +
                    std::to_string(lo) + " and " + std::to_string(hi) + ".");
}

inline std::unique_ptr<Sink> get_writer(const std::string &path, int level) {
    if (path == "-") return std::unique_ptr<Sink>(new FdSink(path));
    if (ends_with(path, ".gz")) {
        validate_compression_level(level, 1, 9, "gzip");
        return std::unique_ptr<Sink>(new GzSink(std::unique_ptr<Sink>(new FdSink(path)), level));
    }
    if (ends_with(path, ".zst")) {
        validate_compression_level(level, 1, 22, "zstd");
        return std::unique_ptr<Sink>(new ZstdSink(std::unique_ptr<Sink>(new FdSink(path)), level));
    }
    if (ends_with(path, ".xz")) {
        validate_compression_level(level, 0, 9, "xz");
        return std::unique_ptr<Sink>(new XzSink(std::unique_ptr<Sink>(new FdSink(path)), level));
    }
    return std::unique_ptr<Sink>(new FdSink(path));
}

}  // namespace deacon
