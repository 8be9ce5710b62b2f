/*
 * stralg_oracle.c -- CPU restatement of the stralg hot path. TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle for the B200 engine.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  Nothing under
 * stralg_b200/ links, imports or calls it; the product path has no CPU fallback.
 *
 * Every function restates, in plain C, what one reference function computes and cites the
 * reference file:line it follows (paths relative to the stralg checkout).  The arrays it
 * produces (SA, ISA, LCP, BWT, C, O, (L,R), match positions) are unique integer functions
 * of the text, so any correct algorithm yields identical bytes; the restatement is pinned
 * against the reference's own golden vectors (tests/golden/) and against the real reference
 * compiled into oracle/_ref/libstralg_ref.so (see oracle/Makefile, tests/test_oracle.py).
 *
 * Parity status: PINNED (golden vectors + live comparison with oracle/_ref).
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * Alphabet remap.  Follows stralg/remap.c:8-31 (table construction: letters present get dense
 * codes in byte order, 0 stays the sentinel), remap.c:43-58 (initialisation) and
 * remap.c:73-88,102-114 (remap returns NULL as soon as it meets a letter without a code).
 * table[c] = code or -1; rev[code] = c or -1.  alphabet_size counts the sentinel.
 * ------------------------------------------------------------------------------------------ */
struct oracle_remap {
    uint32_t alphabet_size;
    int16_t table[256];
    int16_t rev[256];
};

void oracle_remap_init(struct oracle_remap *t, const uint8_t *text, uint64_t n)
{
    uint8_t seen[256];
    memset(seen, 0, sizeof seen);
    for (uint64_t i = 0; i < n; ++i)
        seen[text[i]] = 1;
    for (int c = 0; c < 256; ++c)
        t->table[c] = t->rev[c] = -1;
    t->table[0] = 0;
    t->rev[0] = 0;
    uint32_t next = 1;
    for (int c = 1; c < 256; ++c) {
        if (!seen[c])
            continue;
        t->table[c] = (int16_t)next;
        t->rev[next] = (int16_t)c;
        ++next;
    }
    t->alphabet_size = next;
}

/* Remaps n bytes; writes a trailing 0 at out[n].  Returns 0 on success, -1 if a letter has no
 * code (the reference returns a NULL pointer in that case, remap.c:80-84). */
int oracle_remap_apply(const struct oracle_remap *t, const uint8_t *in, uint64_t n, uint8_t *out)
{
    for (uint64_t i = 0; i < n; ++i) {
        int16_t code = t->table[in[i]];
        if (code < 0)
            return -1;
        out[i] = (uint8_t)code;
    }
    out[n] = 0;
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Suffix array, definitional form.  Follows stralg/suffix_array.c:26-48: sort all len = n+1
 * suffix start positions with strcmp order on NUL-terminated unsigned bytes, so the empty
 * suffix (position n) is first and a proper prefix precedes its extensions.
 * text must hold n non-zero bytes followed by a 0.  O(n log n * LCP): small inputs only.
 * ------------------------------------------------------------------------------------------ */
static const uint8_t *g_sort_text;

static int cmp_suffix(const void *pa, const void *pb)
{
    uint32_t a = *(const uint32_t *)pa, b = *(const uint32_t *)pb;
    return strcmp((const char *)g_sort_text + a, (const char *)g_sort_text + b);
}

void oracle_sa_qsort(const uint8_t *text, uint32_t n, uint32_t *sa)
{
    uint32_t len = n + 1;
    for (uint32_t i = 0; i < len; ++i)
        sa[i] = i;
    g_sort_text = text;
    qsort(sa, len, sizeof *sa, cmp_suffix);
}

/* ------------------------------------------------------------------------------------------
 * Suffix array, scalable form (CPU prefix doubling with counting sorts).  Produces the same
 * array as suffix_array.c:32-48 / sa_is.c:466-509 / sa_is_mem.c:471-494 / skew.c:388-395,
 * which the reference's own tests assert are element-wise equal (tests/stralg/match_test.c:479,
 * 517,539).  Used where the qsort form is too slow (periodic texts).  O(n log n).
 * ------------------------------------------------------------------------------------------ */
int oracle_sa_doubling(const uint8_t *text, uint32_t n, uint32_t *sa)
{
    uint32_t len = n + 1;
    uint32_t *rank = malloc((size_t)len * 4), *tmp = malloc((size_t)len * 4);
    uint32_t *key2 = malloc((size_t)len * 4);
    size_t nb = (size_t)(len > 256 ? len : 256) + 2;
    uint32_t *cnt = malloc(nb * 4);
    if (!rank || !tmp || !key2 || !cnt) {
        free(rank); free(tmp); free(key2); free(cnt);
        return -1;
    }
    /* round 0: counting sort on the first byte (the sentinel at position n is 0). */
    memset(cnt, 0, nb * 4);
    for (uint32_t i = 0; i < len; ++i)
        cnt[(i < n ? text[i] : 0) + 1]++;
    for (uint32_t c = 1; c <= 256; ++c)
        cnt[c] += cnt[c - 1];
    for (uint32_t i = 0; i < len; ++i)
        sa[cnt[i < n ? text[i] : 0]++] = i;
    /* rank = index of the first suffix with the same first byte */
    {
        uint32_t head = 0;
        for (uint32_t r = 0; r < len; ++r) {
            uint8_t c = sa[r] < n ? text[sa[r]] : 0;
            uint8_t p = r ? (sa[r - 1] < n ? text[sa[r - 1]] : 0) : 0;
            if (r == 0 || c != p)
                head = r;
            rank[sa[r]] = head;
        }
    }
    for (uint32_t h = 1;; h *= 2) {
        /* sort by second key rank[i+h]+1 (0 when i+h runs past the end) */
        memset(cnt, 0, nb * 4);
        for (uint32_t i = 0; i < len; ++i) {
            uint32_t k = ((uint64_t)i + h < len) ? rank[i + h] + 1 : 0;
            key2[i] = k;
            cnt[k + 1]++;
        }
        for (size_t c = 1; c < nb; ++c)
            cnt[c] += cnt[c - 1];
        for (uint32_t i = 0; i < len; ++i)
            tmp[cnt[key2[i]]++] = i;
        /* stable sort by first key rank[i] */
        memset(cnt, 0, nb * 4);
        for (uint32_t i = 0; i < len; ++i)
            cnt[rank[i] + 1]++;
        for (size_t c = 1; c < nb; ++c)
            cnt[c] += cnt[c - 1];
        for (uint32_t j = 0; j < len; ++j) {
            uint32_t i = tmp[j];
            sa[cnt[rank[i]]++] = i;
        }
        /* re-rank */
        uint32_t head = 0, distinct = 0;
        tmp[sa[0]] = 0;
        distinct = 1;
        for (uint32_t r = 1; r < len; ++r) {
            uint32_t a = sa[r - 1], b = sa[r];
            if (rank[a] != rank[b] || key2[a] != key2[b]) {
                head = r;
                ++distinct;
            }
            tmp[b] = head;
        }
        memcpy(rank, tmp, (size_t)len * 4);
        if (distinct == len || (uint64_t)h * 2 > len)
            break;
    }
    free(rank); free(tmp); free(key2); free(cnt);
    return 0;
}

/* Inverse suffix array.  Follows stralg/suffix_array.c:55-62. */
void oracle_inverse(const uint32_t *sa, uint32_t len, uint32_t *isa)
{
    for (uint32_t r = 0; r < len; ++r)
        isa[sa[r]] = r;
}

/* Kasai LCP.  Follows stralg/suffix_array.c:64-85: lcp[0] = 0; lcp[j] is the common prefix
 * length of the suffixes at sa[j-1] and sa[j]; the unique sentinel ends every comparison. */
void oracle_lcp_kasai(const uint8_t *text, const uint32_t *sa, const uint32_t *isa, uint32_t len,
                      uint32_t *lcp)
{
    uint32_t l = 0;
    lcp[0] = 0;
    for (uint32_t i = 0; i < len; ++i) {
        uint32_t j = isa[i];
        if (j == 0)
            continue;
        uint32_t k = sa[j - 1];
        while (text[k + l] == text[i + l])
            ++l;
        lcp[j] = l;
        if (l)
            --l;
    }
}

/* BWT symbol at row r.  Follows stralg/bwt.c:13-20. */
void oracle_bwt(const uint8_t *text, const uint32_t *sa, uint32_t len, uint8_t *bwt)
{
    for (uint32_t r = 0; r < len; ++r)
        bwt[r] = sa[r] ? text[sa[r] - 1] : 0;
}

/* C table.  Follows stralg/bwt.c:35-45: histogram over text + sentinel, exclusive prefix. */
void oracle_c_table(const uint8_t *text, uint32_t len, uint32_t sigma, uint32_t *c)
{
    uint64_t counts[256];
    memset(counts, 0, sizeof counts);
    for (uint32_t i = 0; i < len; ++i)
        counts[text[i]]++;
    c[0] = 0;
    for (uint32_t a = 1; a < sigma; ++a)
        c[a] = c[a - 1] + (uint32_t)counts[a - 1];
}

/* Dense O table.  Follows stralg/bwt.c:47-65: o[i*sigma + a] = #{k < i : bwt[k] == a} for
 * i in [0, len].  (len+1)*sigma entries, row-major by position like the reference. */
void oracle_o_table_dense(const uint8_t *bwt, uint32_t len, uint32_t sigma, uint32_t *o)
{
    for (uint32_t a = 0; a < sigma; ++a)
        o[a] = 0;
    for (uint64_t i = 1; i <= len; ++i) {
        const uint32_t *prev = o + (i - 1) * sigma;
        uint32_t *row = o + i * sigma;
        for (uint32_t a = 0; a < sigma; ++a)
            row[a] = prev[a];
        row[bwt[i - 1]]++;
    }
}

/* Streaming restatement of bwt.c:58-65 that keeps every `stride`-th row only:
 * ck[(i/stride)*sigma + a] = O(a, i) for i = 0, stride, 2*stride, ... <= len. */
void oracle_o_checkpoints(const uint8_t *bwt, uint32_t len, uint32_t sigma, uint32_t stride,
                          uint32_t *ck)
{
    uint32_t run[256];
    memset(run, 0, sizeof run);
    for (uint64_t i = 0;; ++i) {
        if (i % stride == 0)
            memcpy(ck + (i / stride) * sigma, run, (size_t)sigma * 4);
        if (i == len)
            break;
        run[bwt[i]]++;
    }
}

static inline uint32_t occ_ck(const uint8_t *bwt, const uint32_t *ck, uint32_t sigma,
                              uint32_t stride, uint8_t a, uint32_t i)
{
    uint32_t b = i / stride;
    uint32_t v = ck[(uint64_t)b * sigma + a];
    for (uint32_t k = b * stride; k < i; ++k)
        v += bwt[k] == a;
    return v;
}

/* O(a, i) probes over the checkpointed table (for parity checks above the dense limit). */
void oracle_o_probe(const uint8_t *bwt, const uint32_t *ck, uint32_t sigma, uint32_t stride,
                    const uint8_t *a, const uint32_t *i, uint64_t count, uint32_t *out)
{
    for (uint64_t q = 0; q < count; ++q)
        out[q] = occ_ck(bwt, ck, sigma, stride, a[q], i[q]);
}

/* ------------------------------------------------------------------------------------------
 * Exact backward search.  Follows stralg/bwt.c:164-199: L = 0, R = len; a pattern longer than
 * len gives L = 1, R = 0; for i = m-1 .. 0 while L < R: L = C(a) + O(a, L), R = C(a) + O(a, R).
 * Patterns are remapped codes, pattern p is pat[off[p] .. off[p+1]).  m = 0 is undefined in the
 * reference (bwt.c:185 underflows); here an empty pattern leaves (0, len) untouched.
 * Dense variant reads the reference-layout O table, the other the checkpointed one.
 * ------------------------------------------------------------------------------------------ */
void oracle_search_dense(const uint32_t *c, const uint32_t *o, uint32_t sigma, uint32_t len,
                         const uint8_t *pat, const uint64_t *off, uint64_t npat, uint32_t *outL,
                         uint32_t *outR)
{
    for (uint64_t p = 0; p < npat; ++p) {
        uint64_t m = off[p + 1] - off[p];
        const uint8_t *x = pat + off[p];
        uint32_t L = 0, R = len;
        if (m > len) {
            L = 1;
            R = 0;
        }
        for (int64_t i = (int64_t)m - 1; i >= 0 && L < R; --i) {
            uint8_t a = x[i];
            L = c[a] + o[(uint64_t)L * sigma + a];
            R = c[a] + o[(uint64_t)R * sigma + a];
        }
        outL[p] = L;
        outR[p] = R;
    }
}

struct search_job {
    const uint32_t *c, *ck;
    const uint8_t *bwt, *pat;
    const uint64_t *off;
    uint32_t sigma, stride, len;
    uint64_t begin, end;
    uint32_t *outL, *outR;
};

static void *search_worker(void *arg)
{
    struct search_job *j = arg;
    for (uint64_t p = j->begin; p < j->end; ++p) {
        uint64_t m = j->off[p + 1] - j->off[p];
        const uint8_t *x = j->pat + j->off[p];
        uint32_t L = 0, R = j->len;
        if (m > j->len) {
            L = 1;
            R = 0;
        }
        for (int64_t i = (int64_t)m - 1; i >= 0 && L < R; --i) {
            uint8_t a = x[i];
            L = j->c[a] + occ_ck(j->bwt, j->ck, j->sigma, j->stride, a, L);
            R = j->c[a] + occ_ck(j->bwt, j->ck, j->sigma, j->stride, a, R);
        }
        j->outL[p] = L;
        j->outR[p] = R;
    }
    return 0;
}

/* Same recurrence over the checkpointed O table, patterns split over `threads` host threads
 * (the table is read-only, so shards are independent; SURVEY 8b threading note). */
void oracle_search_ck(const uint32_t *c, const uint8_t *bwt, const uint32_t *ck, uint32_t sigma,
                      uint32_t stride, uint32_t len, const uint8_t *pat, const uint64_t *off,
                      uint64_t npat, uint32_t *outL, uint32_t *outR, uint32_t threads)
{
    if (threads < 1)
        threads = 1;
    if (threads > 256)
        threads = 256;
    pthread_t tid[256];
    struct search_job jobs[256];
    for (uint32_t t = 0; t < threads; ++t) {
        struct search_job j = {c, ck, bwt, pat, off, sigma, stride, len,
                               npat * t / threads, npat * (t + 1) / threads, outL, outR};
        jobs[t] = j;
        if (threads == 1)
            search_worker(&jobs[t]);
        else
            pthread_create(&tid[t], 0, search_worker, &jobs[t]);
    }
    if (threads > 1)
        for (uint32_t t = 0; t < threads; ++t)
            pthread_join(tid[t], 0);
}

/* Match positions.  Follows stralg/bwt.c:201-217: the iterator yields sa[i] for i in [L, R),
 * in suffix-array order.  Writes a CSR: pos_off[p+1]-pos_off[p] = max(R-L, 0). */
void oracle_locate(const uint32_t *sa, const uint32_t *L, const uint32_t *R, uint64_t npat,
                   uint64_t *pos_off, uint32_t *pos)
{
    uint64_t w = 0;
    pos_off[0] = 0;
    for (uint64_t p = 0; p < npat; ++p) {
        if (L[p] < R[p])
            for (uint32_t i = L[p]; i < R[p]; ++i) {
                if (pos)
                    pos[w] = sa[i];
                ++w;
            }
        pos_off[p + 1] = w;
    }
}

/* ------------------------------------------------------------------------------------------
 * Synthetic inputs shared by the tests and the bench: xorshift64
 * (s ^= s<<13; s ^= s>>7; s ^= s<<17), symbol = 1 + (s>>33) % nsym written as remapped codes.
 * The GPU generator in stralg_b200/csrc uses a counter-based variant; this one is the CPU one.
 * ------------------------------------------------------------------------------------------ */
static inline uint64_t xs64(uint64_t *s)
{
    *s ^= *s << 13;
    *s ^= *s >> 7;
    *s ^= *s << 17;
    return *s;
}

void oracle_random_codes(uint8_t *out, uint64_t n, uint32_t nsym, uint64_t seed)
{
    uint64_t s = seed ? seed : 88172645463325252ull;
    for (uint64_t i = 0; i < n; ++i)
        out[i] = (uint8_t)(1 + (xs64(&s) >> 33) % nsym);
    out[n] = 0;
}

/* Counter-based generators matching stralg_b200/csrc/synth.cu bit for bit (test inputs only). */
static inline uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

void oracle_synth_codes(uint8_t *out, uint64_t n, uint32_t nsym, uint64_t seed)
{
    for (uint64_t i = 0; i < n; ++i)
        out[i] = (uint8_t)(1 + (splitmix64(seed + i) >> 33) % nsym);
    out[n] = 0;
}

void oracle_synth_reads(const uint8_t *text, uint64_t n, uint32_t nsym, uint8_t *reads,
                        uint64_t nreads, uint32_t m, uint32_t miss_per_1024, uint64_t seed)
{
    for (uint64_t q = 0; q < nreads; ++q) {
        uint64_t h = splitmix64(seed ^ (q * 0xD1342543DE82EF95ull));
        int miss = (h & 1023u) < miss_per_1024;
        for (uint32_t j = 0; j < m; ++j) {
            uint8_t c;
            if (miss || n < m)
                c = (uint8_t)(1 + (splitmix64(h + j) >> 33) % nsym);
            else
                c = text[(h >> 10) % (n - m + 1) + j];
            reads[q * m + j] = c;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Threaded driver around the REAL reference iterator (bench.py's reference arm only).
 * `init_fn` is the address of init_bwt_exact_match_iter from oracle/_ref/libstralg_ref.so
 * (stralg/bwt.c:164-199); `table` its struct bwt_table.  The iterator struct is declared here
 * with the layout of stralg/bwt.h:168-173.  Patterns are fixed-length, NUL-terminated copies are
 * made per call because the reference takes C strings.  The table is read-only after
 * construction, so pattern shards run on independent threads.
 * ------------------------------------------------------------------------------------------ */
struct ref_exact_iter {
    const void *sa;
    uint32_t L;
    int64_t i;
    uint32_t R;
};
typedef void (*ref_init_fn)(struct ref_exact_iter *, void *, const uint8_t *);

struct ref_job {
    ref_init_fn init;
    void *table;
    const uint8_t *pat;
    uint32_t m;
    uint64_t begin, end;
    uint32_t *outL, *outR;
};

static void *ref_worker(void *arg)
{
    struct ref_job *j = arg;
    uint8_t *buf = malloc((size_t)j->m + 1);
    for (uint64_t p = j->begin; p < j->end; ++p) {
        memcpy(buf, j->pat + p * j->m, j->m);
        buf[j->m] = 0;
        struct ref_exact_iter it;
        j->init(&it, j->table, buf);
        j->outL[p] = it.L;
        j->outR[p] = it.R;
    }
    free(buf);
    return 0;
}

void oracle_ref_search_threads(void *init_fn, void *table, const uint8_t *pat, uint32_t m,
                               uint64_t npat, uint32_t threads, uint32_t *outL, uint32_t *outR)
{
    if (threads < 1)
        threads = 1;
    if (threads > 256)
        threads = 256;
    pthread_t tid[256];
    struct ref_job jobs[256];
    for (uint32_t t = 0; t < threads; ++t) {
        struct ref_job j = {(ref_init_fn)init_fn, table, pat, m, npat * t / threads,
                            npat * (t + 1) / threads, outL, outR};
        jobs[t] = j;
        pthread_create(&tid[t], 0, ref_worker, &jobs[t]);
    }
    for (uint32_t t = 0; t < threads; ++t)
        pthread_join(tid[t], 0);
}

/* One iterator per pattern, every match fetched: the timing loop of the reference's own harness
 * (performance/suffix_array_search.c:127-141) around ITS functions (pointers into oracle/_ref);
 * patterns are NUL-terminated, stride m + 1.  bench.py's cpu_baseline leg of `compat`. */
typedef int (*ref_next_fn)(struct ref_exact_iter *, uint32_t *);
uint64_t oracle_ref_iter_loop(void *init_fn, void *next_fn, void *table, const uint8_t *patterns, uint32_t m,
                              uint64_t npat)
{
    ref_init_fn init = (ref_init_fn)init_fn;
    ref_next_fn next = (ref_next_fn)next_fn;
    uint64_t hits = 0;
    for (uint64_t q = 0; q < npat; ++q) {
        struct ref_exact_iter it;
        uint32_t pos;
        init(&it, table, patterns + q * ((uint64_t)m + 1));
        while (next(&it, &pos) & 1)
            ++hits;
    }
    return hits;
}

/* ------------------------------------------------------------------------------------------
 * Approximate (edit distance <= d) backward search.  Follows stralg/bwt.c:226-299 (the
 * recursion), :302-382 (D table over the reversed text's O table + the first level, which has
 * no deletions and no D-table test) and stralg/cigar.c:8-31 (run-length CIGAR).
 *
 * The search walks the pattern from its last symbol to its first.  A node is (L, R, i, matched
 * length, edits left).  Children, in this order: for every letter a = 1..sigma-1 a match /
 * substitution step (cost 0 if a == pattern[i] else 1) to (C[a]+O(a,L), C[a]+O(a,R), i-1);
 * then one insertion step (pattern symbol skipped: same interval, i-1, cost 1); then for every
 * letter a deletion step (interval narrowed by a, same i, cost 1).  A node is abandoned when
 * edits_left < D[i] (D[i] = lower bound on the edits pattern[0..i] needs), children with an
 * empty interval are not visited, and a node with i < 0 is a hit: (L, R, matched length) plus
 * the operations of its path in pattern order, run-length encoded.  Hits are reported in the
 * order the depth-first walk meets them; duplicates (same interval, other CIGAR) are kept.
 *
 * Dense O tables in the reference layout (o[i*sigma + a], i in [0, len]); ro may be NULL (no
 * D table: every D[i] = 0, same hits, more work).
 * ------------------------------------------------------------------------------------------ */
struct oracle_approx_out {
    uint64_t nhits, cap;
    uint32_t *L, *R, *mlen;
    uint64_t *cig_off;   /* nhits + 1 */
    char *cigars;        /* NUL-terminated strings, back to back */
    uint64_t cig_bytes, cig_cap;
};

struct approx_ctx {
    const uint32_t *c, *o;
    uint32_t sigma;
    const uint8_t *pat;
    const int *dtab;
    char *ops;  /* operations of the current path, first = the step taken on the LAST symbol */
    struct oracle_approx_out *out;
};

static void approx_hit(struct approx_ctx *x, uint32_t L, uint32_t R, uint32_t mlen, int depth)
{
    struct oracle_approx_out *o = x->out;
    if (o->nhits == o->cap) {
        o->cap = o->cap ? 2 * o->cap : 16;
        o->L = realloc(o->L, o->cap * 4);
        o->R = realloc(o->R, o->cap * 4);
        o->mlen = realloc(o->mlen, o->cap * 4);
        o->cig_off = realloc(o->cig_off, (o->cap + 1) * 8);
    }
    /* worst case: every operation its own run of up to 10 digits + letter */
    uint64_t need = o->cig_bytes + (uint64_t)depth * 12 + 2;
    if (need > o->cig_cap) {
        o->cig_cap = 2 * need;
        o->cigars = realloc(o->cigars, o->cig_cap);
    }
    o->L[o->nhits] = L;
    o->R[o->nhits] = R;
    o->mlen[o->nhits] = mlen;
    o->cig_off[o->nhits] = o->cig_bytes;
    char *w = o->cigars + o->cig_bytes;
    for (int k = depth - 1; k >= 0;) { /* pattern order = path reversed */
        int run = 1;
        while (k - run >= 0 && x->ops[k - run] == x->ops[k])
            ++run;
        w += sprintf(w, "%d%c", run, x->ops[k]);
        k -= run;
    }
    *w++ = 0;
    o->cig_bytes = (uint64_t)(w - o->cigars);
    o->nhits++;
    o->cig_off[o->nhits] = o->cig_bytes;
}

static void approx_node(struct approx_ctx *x, uint32_t L, uint32_t R, int i, uint32_t mlen, int left,
                        int depth, int root)
{
    if (!root) {
        int need = (i >= 0 && x->dtab) ? x->dtab[i] : 0;
        if (left < need)
            return;
        if (i < 0) {
            approx_hit(x, L, R, mlen, depth);
            return;
        }
    }
    const uint32_t s = x->sigma;
    for (uint32_t a = 1; a < s; ++a) {
        int cost = a == x->pat[i] ? 0 : 1;
        uint32_t l2 = x->c[a] + x->o[(uint64_t)L * s + a], r2 = x->c[a] + x->o[(uint64_t)R * s + a];
        if (left - cost < 0 || l2 >= r2)
            continue;
        x->ops[depth] = 'M';
        approx_node(x, l2, r2, i - 1, mlen + 1, left - cost, depth + 1, 0);
    }
    x->ops[depth] = 'I';
    approx_node(x, L, R, i - 1, mlen, left - 1, depth + 1, 0);
    if (root)
        return;
    for (uint32_t a = 1; a < s; ++a) {
        uint32_t l2 = x->c[a] + x->o[(uint64_t)L * s + a], r2 = x->c[a] + x->o[(uint64_t)R * s + a];
        if (l2 >= r2)
            continue;
        x->ops[depth] = 'D';
        approx_node(x, l2, r2, i, mlen + 1, left - 1, depth + 1, 0);
    }
}

/* D table, stralg/bwt.c:319-337: forward over the pattern with the reversed text's O table */
void oracle_approx_dtable(const uint32_t *c, const uint32_t *ro, uint32_t sigma, uint32_t len,
                          const uint8_t *pat, uint32_t m, int *dtab)
{
    uint32_t L = 0, R = len;
    int need = 0;
    for (uint32_t i = 0; i < m; ++i) {
        uint8_t a = pat[i];
        L = c[a] + ro[(uint64_t)L * sigma + a];
        R = c[a] + ro[(uint64_t)R * sigma + a];
        if (L >= R) {
            ++need;
            L = 0;
            R = len;
        }
        dtab[i] = need;
    }
}

void oracle_approx_dense(const uint32_t *c, const uint32_t *o, const uint32_t *ro, uint32_t sigma,
                         uint32_t len, const uint8_t *pat, uint32_t m, int max_edits,
                         struct oracle_approx_out *out)
{
    memset(out, 0, sizeof *out);
    out->cig_off = malloc(8);
    out->cig_off[0] = 0;
    if (m == 0)
        return;
    int *dtab = 0;
    if (ro) {
        dtab = malloc((size_t)m * sizeof(int));
        oracle_approx_dtable(c, ro, sigma, len, pat, m, dtab);
    }
    struct approx_ctx x = {c, o, sigma, pat, dtab, malloc((size_t)m + (size_t)max_edits + 4), out};
    approx_node(&x, 0, len, (int)m - 1, 0, max_edits, 0, 1);
    free(x.ops);
    free(dtab);
}

void oracle_approx_free(struct oracle_approx_out *out)
{
    free(out->L);
    free(out->R);
    free(out->mlen);
    free(out->cig_off);
    free(out->cigars);
    memset(out, 0, sizeof *out);
}
