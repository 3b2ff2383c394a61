/* TEST INFRASTRUCTURE ONLY — see coffee_oracle.h.  Plain C11, single-threaded, written for clarity;
 * sized for corpora the tests finish in seconds.  Never used by the product path. */
#define _GNU_SOURCE
#include "coffee_oracle.h"

#include <stdlib.h>
#include <string.h>

void co_free(void *p) { free(p); }

/* ------------------------------------------------------------------------------------------------
 * Widths (src/index.cpp:183-208).  The reference grows all-ones masks until they cover the document
 * count / the longest document, then takes popcounts; both start at 1 bit. */
static int ones_needed(uint64_t v) {
    uint64_t m = 1;
    int b = 1;
    while (m < v) {
        m = (m << 1) + 1;
        b += 1;
        if (b == 64) break;
    }
    return b;
}

int co_widths(const int64_t *doc_off, int64_t nd, int *bits1, int *bits2, int *width) {
    uint64_t maxlen = 0;
    for (int64_t d = 0; d < nd; ++d) {
        uint64_t len = (uint64_t)(doc_off[d + 1] - doc_off[d]);
        if (len > maxlen) maxlen = len;
    }
    int b1 = ones_needed((uint64_t)nd);
    int b2 = ones_needed(maxlen);
    if (bits1) *bits1 = b1;
    if (bits2) *bits2 = b2;
    if (b1 + b2 > 64) return CO_ERR_TOO_MUCH_DATA;    /* index.cpp:195-197 */
    if (b1 > 32) return CO_ERR_TOO_MANY_OBJECTS;      /* index.cpp:198-200 */
    if (width) *width = (b1 + b2 <= 32) ? 4 : 8;      /* index.cpp:203-208 */
    return CO_OK;
}

/* ------------------------------------------------------------------------------------------------
 * Suffix access (src/index.h:61-73). */
typedef struct {
    const uint8_t *text;
    const int64_t *doc_off;
    int bits1;
    uint64_t mask;
    int64_t offset; /* bytes already known equal inside the current task (index.cpp:86,91) */
} co_ctx;

static inline void suffix_of(const co_ctx *c, uint64_t packed, int64_t extra, const uint8_t **p, int64_t *len) {
    uint64_t doc = packed & c->mask;
    int64_t off = (int64_t)(packed >> c->bits1) + extra;
    int64_t dlen = c->doc_off[doc + 1] - c->doc_off[doc];
    if (off > dlen) off = dlen;
    *p = c->text + c->doc_off[doc] + off;
    *len = dlen - off;
}

/* character(): 0 at end of document, otherwise the byte read as a SIGNED char, minus CHAR_MIN, plus 1
 * (index.h:66-73).  0x80..0xFF therefore map to 1..128 and 0x00..0x7F to 129..256. */
static inline int character_of(const co_ctx *c, uint64_t packed, int64_t extra) {
    const uint8_t *p;
    int64_t len;
    suffix_of(c, packed, extra, &p, &len);
    if (len == 0) return 0;
    return (int)(int8_t)p[0] + 129;
}

/* std::string_view operator< (index.cpp:92-93 and :268): unsigned memcmp, shorter first. */
static inline int view_cmp(const uint8_t *a, int64_t la, const uint8_t *b, int64_t lb) {
    int64_t k = la < lb ? la : lb;
    int r = k ? memcmp(a, b, (size_t)k) : 0;
    if (r) return r;
    return (la > lb) - (la < lb);
}

static int leaf_cmp(const void *pa, const void *pb, void *vc) {
    const co_ctx *c = (const co_ctx *)vc;
    uint64_t a = *(const uint64_t *)pa, b = *(const uint64_t *)pb;
    const uint8_t *sa, *sb;
    int64_t la, lb;
    suffix_of(c, a, c->offset, &sa, &la);
    suffix_of(c, b, c->offset, &sb, &lb);
    int r = view_cmp(sa, la, sb, lb);
    if (r) return r;
    return (a > b) - (a < b); /* canonical tie order (note N2) */
}

static int u64_cmp(const void *pa, const void *pb) {
    uint64_t a = *(const uint64_t *)pa, b = *(const uint64_t *)pb;
    return (a > b) - (a < b);
}

typedef struct {
    int64_t begin, end, offset;
} co_task;

int co_build_sa(const uint8_t *text, const int64_t *doc_off, int64_t nd, uint64_t *sa) {
    int bits1, bits2, width;
    int rc = co_widths(doc_off, nd, &bits1, &bits2, &width);
    if (rc) return rc;
    const int64_t n = nd ? doc_off[nd] - doc_off[0] : 0;
    co_ctx ctx = {text, doc_off, bits1, ((uint64_t)1 << bits1) - 1, 0};
    /* index.cpp:209-215: doc-major fill */
    int64_t k = 0;
    for (int64_t d = 0; d < nd; ++d) {
        int64_t len = doc_off[d + 1] - doc_off[d];
        for (int64_t j = 0; j < len; ++j) sa[k++] = ((uint64_t)j << bits1) | (uint64_t)d;
    }
    if (n == 0) return CO_OK;
    /* index.cpp:218 */
    int64_t chuck = n / 256;
    if (chuck < 4096) chuck = 4096;

    uint64_t *tmp = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)n);
    int64_t cap = 1024, top = 0;
    co_task *stack = (co_task *)malloc(sizeof(co_task) * (size_t)cap);
    if (!tmp || !stack) {
        free(tmp);
        free(stack);
        return CO_ERR_NOMEM;
    }
    stack[top++] = (co_task){0, n, 0};
    while (top) {
        co_task t = stack[--top];
        int64_t len = t.end - t.begin;
        if (len <= chuck) { /* index.cpp:86-94: comparison sort from `offset`, unsigned */
            ctx.offset = t.offset;
            qsort_r(sa + t.begin, (size_t)len, sizeof(uint64_t), leaf_cmp, &ctx);
            continue;
        }
        /* index.cpp:97-125: one MSD step on character(.,offset); the reference permutes in place
         * (unstable), the order inside a bin is settled by the children, so a counting sort through
         * a scratch array yields the same result up to ties. */
        int64_t count[258];
        memset(count, 0, sizeof(count));
        for (int64_t i = t.begin; i < t.end; ++i) count[character_of(&ctx, sa[i], t.offset) + 1] += 1;
        for (int b = 1; b < 258; ++b) count[b] += count[b - 1];
        int64_t cursor[257];
        memcpy(cursor, count, sizeof(cursor));
        for (int64_t i = t.begin; i < t.end; ++i) {
            int ch = character_of(&ctx, sa[i], t.offset);
            tmp[cursor[ch]++] = sa[i];
        }
        memcpy(sa + t.begin, tmp, sizeof(uint64_t) * (size_t)len);
        /* bin 0 = suffixes that ended here: final (index.cpp:119); byte-identical -> canonical order */
        if (count[1] > 1) qsort(sa + t.begin, (size_t)count[1], sizeof(uint64_t), u64_cmp);
        for (int b = 1; b < 257; ++b) {
            int64_t lo = count[b], hi = count[b + 1];
            if (hi > lo) {
                if (top == cap) {
                    cap *= 2;
                    co_task *ns = (co_task *)realloc(stack, sizeof(co_task) * (size_t)cap);
                    if (!ns) {
                        free(tmp);
                        free(stack);
                        return CO_ERR_NOMEM;
                    }
                    stack = ns;
                }
                stack[top++] = (co_task){t.begin + lo, t.begin + hi, t.offset + 1};
            }
        }
    }
    free(tmp);
    free(stack);
    return CO_OK;
}

void co_canonicalise_sa(const uint8_t *text, const int64_t *doc_off, int64_t nd, uint64_t *sa, int64_t n, int bits1) {
    (void)nd;
    co_ctx ctx = {text, doc_off, bits1, ((uint64_t)1 << bits1) - 1, 0};
    int64_t i = 0;
    while (i < n) {
        const uint8_t *pi;
        int64_t li;
        suffix_of(&ctx, sa[i], 0, &pi, &li);
        int64_t j = i + 1;
        while (j < n) {
            const uint8_t *pj;
            int64_t lj;
            suffix_of(&ctx, sa[j], 0, &pj, &lj);
            if (lj != li || (li && memcmp(pi, pj, (size_t)li) != 0)) break;
            ++j;
        }
        if (j - i > 1) qsort(sa + i, (size_t)(j - i), sizeof(uint64_t), u64_cmp);
        i = j;
    }
}

/* ------------------------------------------------------------------------------------------------
 * Query (src/index.cpp:237-326). */
void co_search(const uint8_t *text, const int64_t *doc_off, const uint64_t *sa, int64_t n, int bits1,
               const uint8_t *kw, int64_t m, int64_t *left_out, int64_t *right_out) {
    co_ctx ctx = {text, doc_off, bits1, ((uint64_t)1 << bits1) - 1, 0};
    const uint8_t *s;
    int64_t sl;
    /* lower bound, index.cpp:262-274: L=0, R=size-1, M = L + (R-L)/2, keep left half iff kw <= suffix */
    int64_t L = 0, R = n - 1;
    while (L < R) {
        int64_t M = L + (R - L) / 2;
        suffix_of(&ctx, sa[M], 0, &s, &sl);
        if (view_cmp(kw, m, s, sl) <= 0)
            R = M;
        else
            L = M + 1;
    }
    int64_t left = L;
    /* upper bound, index.cpp:275-287: L=left-1, R=size-1, M = L + (R-L+1)/2, keep right half iff
     * the suffix starts with kw */
    L = left - 1;
    R = n - 1;
    while (L < R) {
        int64_t M = L + (R - L + 1) / 2;
        suffix_of(&ctx, sa[M], 0, &s, &sl);
        if (sl >= m && memcmp(kw, s, (size_t)m) == 0)
            L = M;
        else
            R = M - 1;
    }
    *left_out = left;
    *right_out = L + 1;
}

int64_t co_query(const uint8_t *text, const int64_t *doc_off, const int64_t *ids, const uint64_t *sa, int64_t n,
                 int bits1, const uint8_t *kw, int64_t m, int64_t **pairs_out) {
    *pairs_out = NULL;
    if (m <= 0) return -CO_ERR_EMPTY_KEYWORD; /* index.cpp:239-241 */
    int64_t left, right;
    co_search(text, doc_off, sa, n, bits1, kw, m, &left, &right);
    int64_t occ = right > left ? right - left : 0;
    int64_t *pairs = (int64_t *)malloc(sizeof(int64_t) * 2 * (size_t)(occ + 1));
    uint64_t *docs = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(occ + 1));
    const uint64_t mask = ((uint64_t)1 << bits1) - 1;
    /* index.cpp:289-315: gather doc indices, sort ascending (std::sort or the two-pass LSD radix;
     * both yield the fully sorted sequence) */
    for (int64_t i = 0; i < occ; ++i) docs[i] = sa[left + i] & mask;
    qsort(docs, (size_t)occ, sizeof(uint64_t), u64_cmp);
    /* index.cpp:316-322: run lengths -> (ids[doc], count) */
    int64_t np = 0;
    for (int64_t i = 0; i < occ;) {
        int64_t j = i + 1;
        while (j < occ && docs[j] == docs[i]) ++j;
        pairs[2 * np] = ids[docs[i]];
        pairs[2 * np + 1] = j - i;
        ++np;
        i = j;
    }
    free(docs);
    *pairs_out = pairs;
    return np;
}

/* ------------------------------------------------------------------------------------------------
 * Highlight spans (src/database.cpp:26-138). */
typedef struct {
    int32_t (*go)[256];
    int32_t *fail;
    int32_t *length;
    int32_t size;
} co_ac;

static inline int byte_index(uint8_t b) { return (int)(int8_t)b + 128; } /* database.cpp:97-99 */

static int ac_init(co_ac *a, const uint8_t *kw_bytes, const int64_t *kw_off, int64_t nkw) {
    int64_t total = 4; /* database.cpp:33-36 */
    for (int64_t k = 0; k < nkw; ++k) total += kw_off[k + 1] - kw_off[k];
    a->go = calloc((size_t)total, sizeof(*a->go));
    a->fail = calloc((size_t)total, sizeof(int32_t));
    a->length = calloc((size_t)total, sizeof(int32_t));
    a->size = 1;
    if (!a->go || !a->fail || !a->length) return CO_ERR_NOMEM;
    /* insert, database.cpp:100-110: trie walk, terminal node remembers the keyword length */
    for (int64_t k = 0; k < nkw; ++k) {
        int32_t u = 0;
        for (int64_t i = kw_off[k]; i < kw_off[k + 1]; ++i) {
            int c = byte_index(kw_bytes[i]);
            if (!a->go[u][c]) a->go[u][c] = a->size++;
            u = a->go[u][c];
        }
        a->length[u] = (int32_t)(kw_off[k + 1] - kw_off[k]);
    }
    /* getfail, database.cpp:111-137: BFS; missing edges borrow the failure target's edge; a node
     * inherits the longest keyword length found along its failure chain */
    int32_t *queue = malloc(sizeof(int32_t) * (size_t)total);
    if (!queue) return CO_ERR_NOMEM;
    int64_t head = 0, tail = 0;
    for (int c = 0; c < 256; ++c)
        if (a->go[0][c]) queue[tail++] = a->go[0][c];
    while (head < tail) {
        int32_t r = queue[head++];
        for (int c = 0; c < 256; ++c) {
            int32_t u = a->go[r][c];
            if (!u) {
                a->go[r][c] = a->go[a->fail[r]][c];
                continue;
            }
            queue[tail++] = u;
            int32_t v = a->fail[r];
            while (v && !a->go[v][c]) v = a->fail[v];
            a->fail[u] = a->go[v][c];
            if (a->length[a->fail[u]] > a->length[u]) a->length[u] = a->length[a->fail[u]];
        }
    }
    free(queue);
    return CO_OK;
}

static void ac_destroy(co_ac *a) {
    free(a->go);
    free(a->fail);
    free(a->length);
}

int64_t co_spans(const uint8_t *kw_bytes, const int64_t *kw_off, int64_t nkw, const uint8_t *text, int64_t tlen,
                 int64_t **spans_out) {
    co_ac ac;
    *spans_out = NULL;
    if (ac_init(&ac, kw_bytes, kw_off, nkw)) {
        ac_destroy(&ac);
        return -CO_ERR_NOMEM;
    }
    int64_t *spans = (int64_t *)malloc(sizeof(int64_t) * 2 * (size_t)(tlen + 1));
    int64_t ns = 0;
    int32_t node = 0;
    /* render loop, database.cpp:62-77 */
    for (int64_t i = 0; i < tlen; ++i) {
        node = ac.go[node][byte_index(text[i])];
        int32_t len = ac.length[node];
        if (!len) continue;
        int64_t begin = i - len + 1;
        while (ns && begin <= spans[2 * (ns - 1)]) --ns;        /* swallowed by the new match */
        if (ns && begin <= spans[2 * (ns - 1) + 1])             /* overlaps (shares a byte): extend */
            spans[2 * (ns - 1) + 1] = i;
        else {                                                   /* disjoint, even if adjacent */
            spans[2 * ns] = begin;
            spans[2 * ns + 1] = i;
            ++ns;
        }
    }
    ac_destroy(&ac);
    *spans_out = spans;
    return ns;
}

int64_t co_splice(const uint8_t *text, int64_t tlen, const int64_t *spans, int64_t nspans, const uint8_t *left,
                  int64_t llen, const uint8_t *right, int64_t rlen, uint8_t **out) {
    uint8_t *buf = (uint8_t *)malloc((size_t)(tlen + (llen + rlen) * nspans + 1));
    int64_t w = 0, s = 0;
    /* database.cpp:78-90 */
    for (int64_t i = 0; i < tlen; ++i) {
        if (s < nspans && i == spans[2 * s]) {
            memcpy(buf + w, left, (size_t)llen);
            w += llen;
        }
        buf[w++] = text[i];
        if (s < nspans && i == spans[2 * s + 1]) {
            memcpy(buf + w, right, (size_t)rlen);
            w += rlen;
            ++s;
        }
    }
    buf[w] = 0;
    *out = buf;
    return w;
}
