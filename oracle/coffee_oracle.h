/* TEST INFRASTRUCTURE ONLY — plain-C restatement of CoffeeDB's string-index hot path.
 *
 * This is the parity ORACLE for the CUDA engine in coffeedb_b200/.  It is never linked, imported or
 * executed by the product path; only tests/, bench.py's cpu_baseline / --impl reference legs and
 * __graft_entry__.smoke() may use it.  Every function cites the reference lines it restates
 * (paths relative to /root/reference).  Parity status: PINNED — tests/test_oracle_cpu.py checks this
 * restatement against (a) the reference compiled unmodified (oracle/_ref, built by oracle/Makefile)
 * and (b) golden vectors generated from that build (tests/golden/, script committed beside them),
 * including the README/example.py known answers.
 *
 * Corpus convention used by every entry point: documents are laid out back to back in `text`;
 * document d (its "doc index", = add() order, src/index.cpp:174-177) is text[doc_off[d], doc_off[d+1]).
 */
#ifndef COFFEE_ORACLE_H
#define COFFEE_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { CO_OK = 0, CO_ERR_TOO_MUCH_DATA = 1, CO_ERR_TOO_MANY_OBJECTS = 2, CO_ERR_EMPTY_KEYWORD = 3, CO_ERR_NOMEM = 4 };

/* src/index.cpp:183-208 — field widths and element width of the packed suffix array. */
int co_widths(const int64_t *doc_off, int64_t nd, int *bits1, int *bits2, int *width);

/* src/index.cpp:209-236 + 75-128 — the packed suffix array, element = (offset_in_doc << bits1) | doc.
 * Output is always widened to uint64.  Order restated (SURVEY.md §8 note N1, validated against the
 * compiled reference): a group of suffixes sharing a d-byte prefix is split by SIGNED byte d with
 * end-of-document first iff it is larger than chuck_size = max(4096, n/256); otherwise it is sorted by
 * unsigned memcmp.  Runs of byte-identical suffixes (whose order the reference leaves to its unstable
 * sorts) are emitted in ascending packed value — the canonical form (note N2). */
int co_build_sa(const uint8_t *text, const int64_t *doc_off, int64_t nd, uint64_t *sa_out);

/* Canonicalise a suffix array produced by the reference itself: every maximal run of byte-identical
 * suffixes is re-ordered to ascending packed value.  Nothing else moves. */
void co_canonicalise_sa(const uint8_t *text, const int64_t *doc_off, int64_t nd, uint64_t *sa, int64_t n, int bits1);

/* src/index.cpp:262-287 — the two binary-search recurrences, verbatim midpoints. */
void co_search(const uint8_t *text, const int64_t *doc_off, const uint64_t *sa, int64_t n, int bits1,
               const uint8_t *kw, int64_t m, int64_t *left, int64_t *right);

/* src/index.cpp:237-326 — full query: (ids[doc], occurrences) in ascending doc index.
 * Returns the number of pairs written through *pairs_out (malloc'd, 2 int64 per pair; co_free),
 * or -CO_ERR_EMPTY_KEYWORD. */
int64_t co_query(const uint8_t *text, const int64_t *doc_off, const int64_t *ids, const uint64_t *sa, int64_t n,
                 int bits1, const uint8_t *kw, int64_t m, int64_t **pairs_out);

/* src/database.cpp:26-77 — Aho-Corasick scan of one text for a keyword set; returns the merged
 * inclusive [begin,end] spans (2 int64 per span, malloc'd; co_free). */
int64_t co_spans(const uint8_t *kw_bytes, const int64_t *kw_off, int64_t nkw, const uint8_t *text, int64_t tlen,
                 int64_t **spans_out);

/* src/database.cpp:78-90 — splice left/right markers around the spans.  Returns rendered length;
 * *out is malloc'd (co_free). */
int64_t co_splice(const uint8_t *text, int64_t tlen, const int64_t *spans, int64_t nspans, const uint8_t *left,
                  int64_t llen, const uint8_t *right, int64_t rlen, uint8_t **out);

void co_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
