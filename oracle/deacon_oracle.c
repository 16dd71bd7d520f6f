/*
 * deacon_oracle.c -- CPU restatement of Deacon's filter hot path (plain C, scalar).
 *
 * TEST INFRASTRUCTURE ONLY (see deacon_oracle.h).  "parity unpinned" for minimizer selection:
 * simd-minimizers 1.3.0 / packed-seq 3.2.1 are not in /root/reference; their algorithm is
 * restated from SURVEY.md Appendix A and anchored on the reference's behavioural tests.
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 */
#include "deacon_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ threads -------------- */
/* Minimal dynamic parallel-for on pthreads (the reference uses rayon / paraseq worker pools:
 * src/remote_filter.rs:239-241, src/local_filter.rs:696-709). */
typedef void (*pf_body)(void *ctx, uint64_t begin, uint64_t end, int tid);
typedef struct { pf_body body; void *ctx; uint64_t n, grain; uint64_t *next; int tid; } pf_arg;
static void *pf_thread(void *a_) {
    pf_arg *a = (pf_arg *)a_;
    for (;;) {
        uint64_t b = __atomic_fetch_add(a->next, a->grain, __ATOMIC_RELAXED);
        if (b >= a->n) break;
        uint64_t e = b + a->grain < a->n ? b + a->grain : a->n;
        a->body(a->ctx, b, e, a->tid);
    }
    return NULL;
}
static void parallel_for(uint64_t n, uint64_t grain, int threads, pf_body body, void *ctx) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    if (grain < 1) grain = 1;
    uint64_t next = 0;
    if (threads == 1 || n <= grain) { body(ctx, 0, n, 0); return; }
    pthread_t th[256];
    pf_arg args[256];
    for (int t = 0; t < threads; t++) {
        args[t] = (pf_arg){body, ctx, n, grain, &next, t};
        if (t > 0) pthread_create(&th[t], NULL, pf_thread, &args[t]);
    }
    pf_thread(&args[0]);
    for (int t = 1; t < threads; t++) pthread_join(th[t], NULL);
}

/* ------------------------------------------------------------------ xxh3 ----------------- */
static inline uint64_t rotl64(uint64_t x, unsigned r) { return (x << r) | (x >> (64 - r)); }
static inline uint32_t rotl32(uint32_t x, unsigned r) { r &= 31u; return r ? (x << r) | (x >> (32 - r)) : x; }
static inline uint32_t rotr32(uint32_t x, unsigned r) { r &= 31u; return r ? (x >> r) | (x << (32 - r)) : x; }

/* XXH3_64bits of an 8-byte little-endian input, seed 0: XXH3_len_4to8_64b + XXH3_rrmxmx.
 * Call site: src/filter_common.rs:305, src/minimizers.rs:188 (k <= 32). */
uint64_t dcno_xxh3_u64(uint64_t v) {
    uint64_t x = (v >> 32) | (v << 32);          /* input2 + (input1 << 32) */
    x ^= 0xC73AB174C5ECD5A2ULL;                  /* secret[8..16] ^ secret[16..24] */
    x ^= rotl64(x, 49) ^ rotl64(x, 24);
    x *= 0x9FB21C651E98DF25ULL;
    x ^= (x >> 35) + 8;                          /* + len */
    x *= 0x9FB21C651E98DF25ULL;
    return x ^ (x >> 28);
}

/* XXH3_64bits of a 16-byte little-endian input, seed 0: XXH3_len_9to16_64b + XXH3_avalanche.
 * Call site: src/filter_common.rs:296, src/minimizers.rs:179 (k > 32). */
uint64_t dcno_xxh3_u128(uint64_t lo, uint64_t hi) {
    uint64_t a = lo ^ 0x6782737BEA4239B9ULL;     /* secret[24..32] ^ secret[32..40] */
    uint64_t b = hi ^ 0xAF56BC3B0996523AULL;     /* secret[40..48] ^ secret[48..56] */
    __uint128_t m = (__uint128_t)a * b;
    uint64_t acc = 16 + __builtin_bswap64(a) + b + ((uint64_t)m ^ (uint64_t)(m >> 64));
    acc ^= acc >> 37;
    acc *= 0x165667919E3779F9ULL;
    return acc ^ (acc >> 32);
}

/* ------------------------------------------------------------------ ntHash --------------- */
/* simd-minimizers' 32-bit ntHash seed table, indexed by packed-seq's 2-bit code
 * (A=0, C=1, T=2, G=3) -- SURVEY.md A.2.  Complement of code c is c^2. */
static const uint32_t NT_F[4] = {0x95c60474u, 0x62a02b4cu, 0x82572324u, 0x4be24456u};

uint32_t dcno_nthash_closed(const uint8_t *codes, int k) {
    uint32_t fw = 0, rc = 0;
    for (int i = 0; i < k; i++) {
        fw ^= rotl32(NT_F[codes[i]], (unsigned)(k - 1 - i));
        rc ^= rotl32(NT_F[codes[i] ^ 2], (unsigned)i);
    }
    return fw + rc;
}

/* Rolling evaluation of the same closed form; hk[p] = h(p) >> 16 for p in [0, n-k]. */
static void nthash_keys(const uint8_t *c, size_t n, int k, uint16_t *hk) {
    uint32_t fw = 0, rc = 0;
    for (int i = 0; i < k; i++) {
        fw ^= rotl32(NT_F[c[i]], (unsigned)(k - 1 - i));
        rc ^= rotl32(NT_F[c[i] ^ 2], (unsigned)i);
    }
    hk[0] = (uint16_t)((fw + rc) >> 16);
    for (size_t p = 1; p + k <= n; p++) {
        uint8_t out = c[p - 1], in = c[p + k - 1];
        fw = rotl32(fw, 1) ^ rotl32(NT_F[out], (unsigned)k) ^ NT_F[in];
        rc = rotr32(rc ^ NT_F[out ^ 2] ^ rotl32(NT_F[in ^ 2], (unsigned)k), 1);
        hk[p] = (uint16_t)((fw + rc) >> 16);
    }
}

/* simd_minimizers::canonical_minimizer_positions (call sites src/filter_common.rs:261-267,
 * src/minimizers.rs:143-148) -- SURVEY.md A.3:
 *   key(p) = ntHash(p) >> 16;  left(j)/right(j) = leftmost / rightmost arg-min of key over
 *   k-mer starts j..j+w-1;  window j is canonical iff #(T|G) > #(A|C) among its l=k+w-1 bases;
 *   pick = canonical ? left : right;  consecutive duplicate picks are dropped. */
size_t dcno_minimizer_positions(const uint8_t *codes, size_t n, int k, int w, uint32_t *out_pos) {
    size_t l = (size_t)k + (size_t)w - 1;
    if (k < 1 || w < 1 || n < l) return 0;
    size_t nk = n - (size_t)k + 1, nwin = n - l + 1;
    uint16_t *hk = (uint16_t *)malloc(nk * sizeof(uint16_t));
    nthash_keys(codes, n, k, hk);

    long tg = 0; /* bases with code bit 1 set (T=2, G=3) inside the current window */
    for (size_t i = 0; i < l; i++) tg += (codes[i] >> 1) & 1;

    size_t cnt = 0;
    uint32_t prev = 0;
    size_t lpos = 0, rpos = 0; /* current leftmost / rightmost arg-min; valid when >= j */
    int have = 0;
    for (size_t j = 0; j < nwin; j++) {
        if (j > 0) tg += ((codes[j + l - 1] >> 1) & 1) - ((codes[j - 1] >> 1) & 1);
        if (!have || lpos < j) { /* rescan for leftmost */
            lpos = j;
            for (size_t p = j + 1; p < j + (size_t)w; p++) if (hk[p] < hk[lpos]) lpos = p;
        } else {
            size_t p = j + (size_t)w - 1;
            if (hk[p] < hk[lpos]) lpos = p;
        }
        if (!have || rpos < j) { /* rescan for rightmost */
            rpos = j;
            for (size_t p = j + 1; p < j + (size_t)w; p++) if (hk[p] <= hk[rpos]) rpos = p;
        } else {
            size_t p = j + (size_t)w - 1;
            if (hk[p] <= hk[rpos]) rpos = p;
        }
        have = 1;
        int canonical = 2 * tg > (long)l;
        uint32_t pick = (uint32_t)(canonical ? lpos : rpos);
        if (j == 0 || pick != prev) out_pos[cnt++] = pick;
        prev = pick;
    }
    free(hk);
    return cnt;
}

/* Brute-force twin of the above (closed-form hash, full window scan); tests compare the two. */
size_t dcno_minimizer_positions_brute(const uint8_t *codes, size_t n, int k, int w, uint32_t *out_pos) {
    size_t l = (size_t)k + (size_t)w - 1;
    if (k < 1 || w < 1 || n < l) return 0;
    size_t cnt = 0;
    uint32_t prev = 0;
    for (size_t j = 0; j + l <= n; j++) {
        uint32_t bestl = 0, bestr = 0, kl = 0xffffffffu, kr = 0xffffffffu;
        for (size_t p = j; p < j + (size_t)w; p++) {
            uint32_t key = dcno_nthash_closed(codes + p, k) >> 16;
            if (key < kl) { kl = key; bestl = (uint32_t)p; }
            if (key <= kr) { kr = key; bestr = (uint32_t)p; }
        }
        size_t tg = 0;
        for (size_t i = j; i < j + l; i++) tg += (codes[i] >> 1) & 1;
        uint32_t pick = (2 * tg > l) ? bestl : bestr;
        if (j == 0 || pick != prev) out_pos[cnt++] = pick;
        prev = pick;
    }
    return cnt;
}

/* ------------------------------------------------------------------ k-mer value ---------- */
/* simd_minimizers::iter_canonical_minimizer_values{,_u128} -- SURVEY.md A.4:
 * fw = sum code(p+i) << 2i ; rc = sum (code(p+k-1-i) ^ 2) << 2i ; value = min(fw, rc). */
static uint64_t canonical_hash(const uint8_t *codes, int k) {
    if (k <= 32) {
        uint64_t fw = 0, rc = 0;
        for (int i = 0; i < k; i++) {
            fw |= (uint64_t)codes[i] << (2 * i);
            rc |= (uint64_t)(codes[k - 1 - i] ^ 2) << (2 * i);
        }
        return dcno_xxh3_u64(fw < rc ? fw : rc);
    }
    __uint128_t fw = 0, rc = 0;
    for (int i = 0; i < k; i++) {
        fw |= (__uint128_t)codes[i] << (2 * i);
        rc |= (__uint128_t)(codes[k - 1 - i] ^ 2) << (2 * i);
    }
    __uint128_t v = fw < rc ? fw : rc;
    return dcno_xxh3_u128((uint64_t)v, (uint64_t)(v >> 64));
}

static inline int is_acgt(uint8_t b) { /* src/filter_common.rs:254, src/minimizers.rs:9-14 */
    switch (b) { case 'A': case 'C': case 'G': case 'T': case 'a': case 'c': case 'g': case 't': return 1; }
    return 0;
}

/* ------------------------------------------------------------------ filter flavour ------- */
/* src/filter_common.rs:211-310 */
size_t dcno_extract_filter(const uint8_t *seq, size_t len, size_t prefix_len, int k, int w,
                           uint64_t *out_hashes, uint32_t *out_pos) {
    if (len < (size_t)k) return 0;                                  /* :217-219 raw length */
    size_t n = (prefix_len > 0 && len > prefix_len) ? prefix_len : len; /* :222-226 */
    if (n > 0 && seq[n - 1] == '\n') n--;                           /* :229 strip one trailing \n */
    if (n == 0) return 0;
    uint8_t *codes = (uint8_t *)malloc(n);
    uint8_t *bad = (uint8_t *)malloc(n);
    for (size_t i = 0; i < n; i++) {
        codes[i] = (seq[i] >> 1) & 3;   /* packed-seq lossy 2-bit packing (:238), SURVEY A.1 */
        bad[i] = !is_acgt(seq[i]);      /* :245-258 */
    }
    uint32_t *pos = (uint32_t *)malloc((n + 1) * sizeof(uint32_t));
    size_t np = dcno_minimizer_positions(codes, n, k, w, pos);      /* :261-267 */
    size_t cnt = 0;
    for (size_t i = 0; i < np; i++) {
        uint32_t p = pos[i];
        int ok = 1;
        for (int t = 0; t < k; t++) if (bad[p + t]) { ok = 0; break; } /* :275-286 */
        if (!ok) continue;
        if (out_pos) out_pos[cnt] = p;
        out_hashes[cnt++] = canonical_hash(codes + p, k);           /* :289-307 */
    }
    free(pos); free(bad); free(codes);
    return cnt;
}

/* ------------------------------------------------------------------ index flavour -------- */
static inline uint8_t iupac_map(uint8_t b) { /* src/minimizers.rs:24-43 */
    switch (b) {
        case 'A': case 'a': return 'A';
        case 'C': case 'c': return 'C';
        case 'G': case 'g': return 'G';
        case 'T': case 't': return 'T';
        case 'R': case 'r': return 'G';
        case 'Y': case 'y': return 'C';
        case 'S': case 's': return 'G';
        case 'W': case 'w': return 'A';
        case 'K': case 'k': return 'G';
        case 'M': case 'm': return 'C';
        case 'B': case 'b': return 'C';
        case 'D': case 'd': return 'G';
        case 'H': case 'h': return 'C';
        case 'V': case 'v': return 'G';
        case 'N': case 'n': return 'C';
        default: return 'C';
    }
}

/* src/minimizers.rs:73-121 -- f32 arithmetic, same operation order; compile with
 * -ffp-contract=off so p*log2f(p) is not fused. */
float dcno_scaled_entropy(const uint8_t *kmer, int k) {
    if (k < 10) return 1.0f;
    uint8_t counts[4] = {0, 0, 0, 0}, total = 0;
    for (int i = 0; i < k; i++) {
        switch (kmer[i]) {
            case 'A': case 'a': counts[0]++; total++; break;
            case 'C': case 'c': counts[1]++; total++; break;
            case 'G': case 'g': counts[2]++; total++; break;
            case 'T': case 't': counts[3]++; total++; break;
            default: break;
        }
    }
    if (total == 0) return 1.0f;
    float total_f = (float)total, entropy = 0.0f;
    for (int i = 0; i < 4; i++) {
        if (counts[i] > 0) {
            float p = (float)counts[i] / total_f;
            entropy -= p * log2f(p);
        }
    }
    return entropy / 2.0f;
}

/* src/minimizers.rs:125-191 */
size_t dcno_extract_index(const uint8_t *seq, size_t len, int k, int w, float entropy_thr,
                          uint64_t *out_hashes) {
    if (len < (size_t)k) return 0;                                  /* :135-137 */
    uint8_t *codes = (uint8_t *)malloc(len);
    for (size_t i = 0; i < len; i++) codes[i] = (iupac_map(seq[i]) >> 1) & 3; /* :139 + AsciiSeq */
    uint32_t *pos = (uint32_t *)malloc((len + 1) * sizeof(uint32_t));
    size_t np = dcno_minimizer_positions(codes, len, k, w, pos);    /* :143-148 */
    size_t cnt = 0;
    for (size_t i = 0; i < np; i++) {
        uint32_t p = pos[i];
        int ok = 1;
        for (int t = 0; t < k; t++) if (!is_acgt(seq[p + t])) { ok = 0; break; } /* :157-160 */
        if (!ok) continue;
        if (entropy_thr != 0.0f && !(dcno_scaled_entropy(seq + p, k) >= entropy_thr)) continue; /* :163-168 */
        out_hashes[cnt++] = canonical_hash(codes + p, k);           /* :172-190 */
    }
    free(pos); free(codes);
    return cnt;
}

/* ------------------------------------------------------------------ classification ------- */
/* src/filter_common.rs:84-96.  f64::round = half away from zero = C round(); `as usize`
 * saturates, negative/NaN -> 0. */
uint64_t dcno_required_hits(uint64_t abs_thr, double rel_thr, uint64_t total) {
    uint64_t rel_required = 0;
    if (total != 0) {
        double r = round(rel_thr * (double)total);
        if (!(r > 0.0)) rel_required = 0;
        else if (r >= 18446744073709551616.0) rel_required = UINT64_MAX;
        else rel_required = (uint64_t)r;
        if (rel_required < 1) rel_required = 1;
    }
    return abs_thr > rel_required ? abs_thr : rel_required;
}

/* src/filter_common.rs:99-112 */
int dcno_meets_criteria(uint64_t hits, uint64_t total, uint64_t abs_thr, double rel_thr, int deplete) {
    uint64_t required = dcno_required_hits(abs_thr, rel_thr, total);
    return deplete ? hits < required : hits >= required;
}

/* ------------------------------------------------------------------ exact set ------------ */
struct dcno_set {
    uint64_t *slots;   /* 0 = empty */
    uint64_t mask;
    uint64_t len;
    int has_zero;
};

static inline uint64_t set_slot(uint64_t key, uint64_t mask) {
    return ((key * 0x9E3779B97F4A7C15ULL) >> 20) & mask;
}

dcno_set *dcno_set_new(uint64_t expected) {
    dcno_set *s = (dcno_set *)calloc(1, sizeof(dcno_set));
    uint64_t cap = 64;
    while (cap < expected * 2 + 16) cap <<= 1;
    s->slots = (uint64_t *)calloc(cap, sizeof(uint64_t));
    s->mask = cap - 1;
    return s;
}
void dcno_set_free(dcno_set *s) { if (s) { free(s->slots); free(s); } }
uint64_t dcno_set_len(const dcno_set *s) { return s->len + (uint64_t)s->has_zero; }

static void set_grow(dcno_set *s) {
    uint64_t ocap = s->mask + 1, *old = s->slots;
    uint64_t cap = ocap * 2;
    s->slots = (uint64_t *)calloc(cap, sizeof(uint64_t));
    s->mask = cap - 1;
    for (uint64_t i = 0; i < ocap; i++) {
        uint64_t key = old[i];
        if (!key) continue;
        uint64_t p = set_slot(key, s->mask);
        while (s->slots[p]) p = (p + 1) & s->mask;
        s->slots[p] = key;
    }
    free(old);
}

typedef struct { dcno_set *s; const uint64_t *keys; uint64_t added; int zero; } ins_ctx;
static void ins_body(void *c_, uint64_t b, uint64_t e, int tid) {
    (void)tid;
    ins_ctx *c = (ins_ctx *)c_;
    dcno_set *s = c->s;
    uint64_t added = 0;
    for (uint64_t i = b; i < e; i++) {
        uint64_t key = c->keys[i];
        if (key == 0) { __atomic_store_n(&c->zero, 1, __ATOMIC_RELAXED); continue; }
        uint64_t p = set_slot(key, s->mask);
        for (;;) {
            uint64_t cur = __atomic_load_n(&s->slots[p], __ATOMIC_RELAXED);
            if (cur == key) break;
            if (cur == 0) {
                uint64_t exp = 0;
                if (__atomic_compare_exchange_n(&s->slots[p], &exp, key, 0, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) { added++; break; }
                if (exp == key) break;
            }
            p = (p + 1) & s->mask;
        }
    }
    __atomic_fetch_add(&c->added, added, __ATOMIC_RELAXED);
}
void dcno_set_insert_many(dcno_set *s, const uint64_t *keys, uint64_t n, int threads) {
    while ((s->len + n) * 2 + 16 > s->mask + 1) set_grow(s);
    ins_ctx c = {s, keys, 0, 0};
    parallel_for(n, 1u << 16, threads, ins_body, &c);
    s->len += c.added;
    if (c.zero) s->has_zero = 1;
}

int dcno_set_contains(const dcno_set *s, uint64_t key) {
    if (key == 0) return s->has_zero;
    uint64_t p = set_slot(key, s->mask);
    for (;;) {
        uint64_t cur = s->slots[p];
        if (cur == key) return 1;
        if (cur == 0) return 0;
        p = (p + 1) & s->mask;
    }
}

void dcno_set_keys(const dcno_set *s, uint64_t *out) {
    uint64_t c = 0;
    if (s->has_zero) out[c++] = 0;
    for (uint64_t i = 0; i <= s->mask; i++) if (s->slots[i]) out[c++] = s->slots[i];
}

/* ------------------------------------------------------------------ lookup + classify ---- */
static int cmp_u64(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : x > y;
}

/* src/filter_common.rs:129-155 / 172-198: distinct hashes that are in the index. */
static uint64_t distinct_hits(const dcno_set *idx, const uint64_t *h, size_t n, uint64_t *scratch) {
    size_t m = 0;
    for (size_t i = 0; i < n; i++) if (dcno_set_contains(idx, h[i])) scratch[m++] = h[i];
    if (m < 2) return m;
    if (m <= 48) { /* small: quadratic */
        uint64_t d = 0;
        for (size_t i = 0; i < m; i++) {
            int seen = 0;
            for (size_t t = 0; t < i; t++) if (scratch[t] == scratch[i]) { seen = 1; break; }
            d += !seen;
        }
        return d;
    }
    qsort(scratch, m, sizeof(uint64_t), cmp_u64);
    uint64_t d = 1;
    for (size_t i = 1; i < m; i++) d += scratch[i] != scratch[i - 1];
    return d;
}

typedef struct {
    const dcno_set *idx; const uint8_t *bases; const uint64_t *hashes; const uint64_t *rec_off;
    int paired; uint64_t prefix_len; int k, w; uint64_t abs_thr; double rel_thr; int deplete;
    uint8_t *keep; uint32_t *hits; uint32_t *total;
} batch_ctx;

/* src/remote_filter.rs:230-301 */
static void lookup_body(void *c_, uint64_t b, uint64_t e, int tid) {
    (void)tid;
    batch_ctx *c = (batch_ctx *)c_;
    uint64_t *scratch = NULL;
    size_t cap = 0;
    for (uint64_t r = b; r < e; r++) {
        size_t n = (size_t)(c->rec_off[r + 1] - c->rec_off[r]);
        if (n > cap) { cap = n * 2; free(scratch); scratch = (uint64_t *)malloc(cap * sizeof(uint64_t)); }
        uint64_t h = distinct_hits(c->idx, c->hashes + c->rec_off[r], n, scratch);
        c->hits[r] = (uint32_t)h;
        c->total[r] = (uint32_t)n;
        c->keep[r] = (uint8_t)dcno_meets_criteria(h, n, c->abs_thr, c->rel_thr, c->deplete);
    }
    free(scratch);
}
void dcno_lookup_batch(const dcno_set *idx, const uint64_t *hashes, const uint64_t *rec_off,
                       uint32_t n_rec, uint64_t abs_thr, double rel_thr, int deplete,
                       uint8_t *keep, uint32_t *hits, uint32_t *total, int threads) {
    batch_ctx c = {idx, NULL, hashes, rec_off, 0, 0, 0, 0, abs_thr, rel_thr, deplete, keep, hits, total};
    parallel_for(n_rec, 512, threads, lookup_body, &c);
}

/* src/local_filter.rs:221-285 (+ src/filter_common.rs:312-348 for pairs) */
static void filter_body(void *c_, uint64_t b, uint64_t e, int tid) {
    (void)tid;
    batch_ctx *c = (batch_ctx *)c_;
    uint64_t *hbuf = NULL, *scratch = NULL;
    size_t cap = 0;
    for (uint64_t u = b; u < e; u++) {
        uint64_t r0 = c->paired ? 2 * u : u, r1 = c->paired ? 2 * u + 2 : u + 1;
        size_t need = (size_t)(c->rec_off[r1] - c->rec_off[r0]) + 2;
        if (need > cap) {
            cap = need * 2;
            free(hbuf); free(scratch);
            hbuf = (uint64_t *)malloc(cap * sizeof(uint64_t));
            scratch = (uint64_t *)malloc(cap * sizeof(uint64_t));
        }
        size_t n = 0;
        for (uint64_t r = r0; r < r1; r++) {
            size_t len = (size_t)(c->rec_off[r + 1] - c->rec_off[r]);
            /* :324 / :336 -- a mate shorter than k contributes nothing (same guard as :217) */
            n += dcno_extract_filter(c->bases + c->rec_off[r], len, (size_t)c->prefix_len, c->k, c->w, hbuf + n, NULL);
        }
        uint64_t h = distinct_hits(c->idx, hbuf, n, scratch);
        c->hits[u] = (uint32_t)h;
        c->total[u] = (uint32_t)n;
        c->keep[u] = (uint8_t)dcno_meets_criteria(h, n, c->abs_thr, c->rel_thr, c->deplete);
    }
    free(hbuf); free(scratch);
}
void dcno_filter_batch(const dcno_set *idx, const uint8_t *bases, const uint64_t *rec_off,
                       uint32_t n_rec, int paired, uint64_t prefix_len, int k, int w,
                       uint64_t abs_thr, double rel_thr, int deplete,
                       uint8_t *keep, uint32_t *hits, uint32_t *total, int threads) {
    uint32_t n_unit = paired ? n_rec / 2 : n_rec;
    batch_ctx c = {idx, bases, NULL, rec_off, paired, prefix_len, k, w, abs_thr, rel_thr, deplete, keep, hits, total};
    parallel_for(n_unit, 256, threads, filter_body, &c);
}

/* src/index.rs:225-284: extraction per record (parallel), union into the set.
 * Long contigs are cut into chunks with an (l-1)-base overlap: a chunk sees exactly the windows
 * whose start lies in its range, and only the SET of hashes matters for an index. */
#define DCNO_CHUNK (1u << 20)
typedef struct {
    const uint8_t *seq; size_t len; int k, w; float thr;
    uint64_t **chunk_h; size_t *chunk_n;   /* per-chunk results, merged after the parallel phase */
} build_ctx;
static void build_body(void *c_, uint64_t b, uint64_t e, int tid) {
    (void)tid;
    build_ctx *c = (build_ctx *)c_;
    size_t l = (size_t)c->k + (size_t)c->w - 1;
    uint64_t *hb = (uint64_t *)malloc((DCNO_CHUNK + l + 8) * sizeof(uint64_t));
    for (uint64_t ch = b; ch < e; ch++) {
        size_t s = (size_t)ch * DCNO_CHUNK, en = s + DCNO_CHUNK + l - 1;
        if (en > c->len) en = c->len;
        size_t n = dcno_extract_index(c->seq + s, en - s, c->k, c->w, c->thr, hb);
        c->chunk_h[ch] = (uint64_t *)malloc((n ? n : 1) * sizeof(uint64_t));
        memcpy(c->chunk_h[ch], hb, n * sizeof(uint64_t));
        c->chunk_n[ch] = n;
    }
    free(hb);
}
void dcno_index_build(dcno_set *dst, const uint8_t *bases, const uint64_t *rec_off,
                      uint32_t n_rec, int k, int w, float entropy_thr, int threads) {
    for (uint32_t r = 0; r < n_rec; r++) {
        size_t len = (size_t)(rec_off[r + 1] - rec_off[r]);
        if (len < (size_t)k) continue;                               /* src/minimizers.rs:135-137 */
        size_t nchunk = (len + DCNO_CHUNK - 1) / DCNO_CHUNK;
        build_ctx c = {bases + rec_off[r], len, k, w, entropy_thr,
                       (uint64_t **)calloc(nchunk, sizeof(uint64_t *)), (size_t *)calloc(nchunk, sizeof(size_t))};
        parallel_for(nchunk, 1, threads, build_body, &c);
        size_t total = 0;
        for (size_t i = 0; i < nchunk; i++) total += c.chunk_n[i];
        uint64_t *all = (uint64_t *)malloc((total ? total : 1) * sizeof(uint64_t));
        size_t at = 0;
        for (size_t i = 0; i < nchunk; i++) {
            memcpy(all + at, c.chunk_h[i], c.chunk_n[i] * sizeof(uint64_t));
            at += c.chunk_n[i];
            free(c.chunk_h[i]);
        }
        dcno_set_insert_many(dst, all, total, threads);               /* src/index.rs:267-284 extend */
        free(all); free(c.chunk_h); free(c.chunk_n);
    }
}

/* ------------------------------------------------------------------ .idx codec ----------- */
/* bincode 2 standard config: u8 raw; u64/usize varint: <251 -> 1 byte; <2^16 -> 0xFB+u16 LE;
 * <2^32 -> 0xFC+u32 LE; else 0xFD+u64 LE.  Layout: version,k,w, varint count, varint keys
 * (src/index.rs:17-22, 130-164). */
static size_t put_varint(uint8_t *o, uint64_t v) {
    if (v < 251) { o[0] = (uint8_t)v; return 1; }
    if (v < (1ULL << 16)) { o[0] = 0xFB; o[1] = (uint8_t)v; o[2] = (uint8_t)(v >> 8); return 3; }
    if (v < (1ULL << 32)) { o[0] = 0xFC; for (int i = 0; i < 4; i++) o[1 + i] = (uint8_t)(v >> (8 * i)); return 5; }
    o[0] = 0xFD; for (int i = 0; i < 8; i++) o[1 + i] = (uint8_t)(v >> (8 * i)); return 9;
}
static int get_varint(const uint8_t *b, size_t len, size_t *off, uint64_t *v) {
    if (*off >= len) return -1;
    uint8_t t = b[(*off)++];
    int nb;
    if (t < 251) { *v = t; return 0; }
    else if (t == 0xFB) nb = 2; else if (t == 0xFC) nb = 4; else if (t == 0xFD) nb = 8;
    else return -1; /* 0xFE = u128: not valid for u64 */
    if (*off + (size_t)nb > len) return -1;
    uint64_t x = 0;
    for (int i = 0; i < nb; i++) x |= (uint64_t)b[*off + i] << (8 * i);
    *off += (size_t)nb;
    *v = x;
    return 0;
}

size_t dcno_idx_encode(const uint64_t *keys, uint64_t n, uint8_t k, uint8_t w, uint8_t *out) {
    size_t o = 0;
    out[o++] = 2; out[o++] = k; out[o++] = w;  /* IndexHeader::new, src/index.rs:25-31 */
    o += put_varint(out + o, n);
    for (uint64_t i = 0; i < n; i++) o += put_varint(out + o, keys[i]);
    return o;
}
int dcno_idx_decode_header(const uint8_t *buf, size_t len, uint8_t *version, uint8_t *k, uint8_t *w,
                           uint64_t *count, size_t *body_off) {
    if (len < 4) return -1;
    *version = buf[0]; *k = buf[1]; *w = buf[2];
    if (*version != 2) return -2;               /* src/index.rs:34-40 */
    size_t off = 3;
    if (get_varint(buf, len, &off, count)) return -1;
    *body_off = off;
    return 0;
}
int dcno_idx_decode_keys(const uint8_t *buf, size_t len, size_t body_off, uint64_t count, uint64_t *out) {
    size_t off = body_off;
    for (uint64_t i = 0; i < count; i++) if (get_varint(buf, len, &off, &out[i])) return -1;
    return 0;
}
