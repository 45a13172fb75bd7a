/* oracle/bpe_oracle.c -- CPU ORACLE. TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference's BPE hot path (justinhj/minbpe-cc), used only by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as the
 * CHECKER. Nothing in minbpe_cc_b200/ may include, link or call this file; the product path has
 * no CPU fallback.
 *
 * Parity status: PINNED. The functions below are checked (tests/test_oracle.py) against
 *   - the reference's own known-answer tests code/test/test.cpp:50-80, :82-106, :136-186,
 *   - the golden .model/.enc files produced by the reference itself compiled here
 *     (oracle/_ref/ref_driver, see oracle/Makefile and tests/golden/make_golden.py).
 *
 * Citations are path:line under /root/reference/code/include/.
 *
 *   oracle_split           Tokenizer.h:500-544   regex pre-tokenisation loop (PCRE2, sequential)
 *   oracle_train_rescan    Tokenizer.h:551-589   literal restatement: full walk of every chunk per merge;
 *                                                FIRST = recount from scratch each merge (:581-585),
 *                                                LEXICAL = incremental deltas (:202-306), entries never erased
 *   oracle_train_indexed   same results, O(touched) per merge (inverted index + lazy heap); used as the
 *                          judge for corpora where the literal walk would take days
 *   oracle_encode          Tokenizer.h:325-377   left-to-right any-known-pair scan, repeated to fixpoint
 *   oracle_decode          Tokenizer.h:725-751   id -> bytes with special override and invalid-id skip
 *
 * Chunks may be weighted (deduplicated): exact in both modes when unique chunks are ordered by first
 * appearance. With all weights 1 and every regex match its own chunk this is the
 * reference's data layout exactly.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "shim/pcre2.h"

#define MODE_FIRST 0
#define MODE_LEXICAL 1

typedef uint32_t tok_t;

/* ------------------------------------------------------------------------------------------------
 * pair table: open addressing, insertion-ordered dense entries (entries are never erased, PairCount.h:146/:254)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    uint64_t key;   /* (a << 32) | b : numeric order == lexicographic order of (a, b) */
    int32_t count;  /* PairCount.h:57 / :188  'int count' */
    uint64_t order; /* PairCount.h:58 insert_order (FIRST) */
} entry_t;

typedef struct {
    entry_t *e;
    uint64_t n, cap_e;
    int64_t *slot; /* -1 empty, else index into e */
    uint64_t cap_s;
} table_t;

static inline uint64_t mix64(uint64_t x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

static void table_init(table_t *t) {
    t->cap_e = 1024;
    t->e = (entry_t *)malloc(t->cap_e * sizeof(entry_t));
    t->n = 0;
    t->cap_s = 4096;
    t->slot = (int64_t *)malloc(t->cap_s * sizeof(int64_t));
    memset(t->slot, 0xff, t->cap_s * sizeof(int64_t));
}
static void table_free(table_t *t) {
    free(t->e);
    free(t->slot);
}
static void table_clear(table_t *t) {
    t->n = 0;
    memset(t->slot, 0xff, t->cap_s * sizeof(int64_t));
}
static int64_t table_find(const table_t *t, uint64_t key) {
    uint64_t m = t->cap_s - 1, h = mix64(key) & m;
    while (t->slot[h] >= 0) {
        if (t->e[t->slot[h]].key == key) return t->slot[h];
        h = (h + 1) & m;
    }
    return -1;
}
static void table_grow_slots(table_t *t) {
    uint64_t ncap = t->cap_s * 2, m = ncap - 1;
    int64_t *ns = (int64_t *)malloc(ncap * sizeof(int64_t));
    memset(ns, 0xff, ncap * sizeof(int64_t));
    for (uint64_t i = 0; i < t->n; i++) {
        uint64_t h = mix64(t->e[i].key) & m;
        while (ns[h] >= 0) h = (h + 1) & m;
        ns[h] = (int64_t)i;
    }
    free(t->slot);
    t->slot = ns;
    t->cap_s = ncap;
}
/* create_or_modify_pair (PairCount.h:141-152, :249-260). Returns entry index. */
static int64_t table_add(table_t *t, tok_t a, tok_t b, int32_t delta) {
    uint64_t key = ((uint64_t)a << 32) | b;
    int64_t i = table_find(t, key);
    if (i >= 0) {
        t->e[i].count += delta;
        return i;
    }
    if ((t->n + 1) * 2 > t->cap_s) table_grow_slots(t);
    if (t->n == t->cap_e) {
        t->cap_e *= 2;
        t->e = (entry_t *)realloc(t->e, t->cap_e * sizeof(entry_t));
    }
    uint64_t m = t->cap_s - 1, h = mix64(key) & m;
    while (t->slot[h] >= 0) h = (h + 1) & m;
    t->slot[h] = (int64_t)t->n;
    t->e[t->n].key = key;
    t->e[t->n].count = delta;
    t->e[t->n].order = t->n; /* next_insert++ (PairCount.h:149) */
    return (int64_t)t->n++;
}

/* get_top_pair_count: begin() of the ordered index.
 * FIRST   (PairCount.h:66-74):   count desc, insert_order asc
 * LEXICAL (PairCount.h:195-207): count desc, first asc, second asc */
static int64_t table_top(const table_t *t, int mode) {
    int64_t best = -1;
    for (uint64_t i = 0; i < t->n; i++) {
        if (best < 0) {
            best = (int64_t)i;
            continue;
        }
        const entry_t *x = &t->e[i], *y = &t->e[best];
        if (x->count > y->count)
            best = (int64_t)i;
        else if (x->count == y->count) {
            if (mode == MODE_FIRST ? (x->order < y->order) : (x->key < y->key)) best = (int64_t)i;
        }
    }
    return best;
}

/* ------------------------------------------------------------------------------------------------
 * Known-answer hook: drive a PairCount the way code/test/test.cpp:15-106 does.
 * ops: n triples (a, b, delta). Returns top pair after all ops in *top_a,*top_b (or rc 1 if empty).
 * ---------------------------------------------------------------------------------------------- */
int oracle_paircount_top(const int32_t *ops, uint64_t n_ops, int mode, uint32_t *top_a, uint32_t *top_b,
                         int32_t *top_count, uint64_t *n_pairs) {
    table_t t;
    table_init(&t);
    for (uint64_t i = 0; i < n_ops; i++) table_add(&t, (tok_t)ops[3 * i], (tok_t)ops[3 * i + 1], ops[3 * i + 2]);
    int64_t b = table_top(&t, mode);
    *n_pairs = t.n;
    int rc = 1;
    if (b >= 0) {
        *top_a = (uint32_t)(t.e[b].key >> 32);
        *top_b = (uint32_t)t.e[b].key;
        *top_count = t.e[b].count;
        rc = 0;
    }
    table_free(&t);
    return rc;
}

/* ------------------------------------------------------------------------------------------------
 * oracle_split: Tokenizer.h:500-544 (train) == :672-703 (encode). Sequential PCRE2 match loop with
 * PCRE2_NO_UTF_CHECK, empty-match skip (:529-533). pattern "" => whole text is one chunk (:541-544).
 * jit != 0 mirrors the constructor (:435), jit == 0 mirrors load() (:793); results are the same.
 * Writes chunk [start,end) byte ranges. Returns number of chunks, or <0 on error. If starts==NULL only counts.
 * ---------------------------------------------------------------------------------------------- */
int64_t oracle_split(const uint8_t *text, uint64_t n, const char *pattern, int jit, uint64_t *starts, uint64_t *ends,
                     uint64_t cap) {
    uint64_t k = 0;
    size_t plen = strlen(pattern);
    if (plen == 0) {
        if (starts && cap >= 1) {
            starts[0] = 0;
            ends[0] = n;
        }
        return 1;
    }
    uint32_t options = PCRE2_UTF | PCRE2_UCP; /* :407 */
    if (strstr(pattern, "(?i:")) options |= PCRE2_CASELESS; /* :413-415 */
    int err;
    PCRE2_SIZE erroff;
    pcre2_code_8 *code = pcre2_compile_8((PCRE2_SPTR8)pattern, plen, options, &err, &erroff, NULL);
    if (!code) return -1;
    if (jit) pcre2_jit_compile_8(code, PCRE2_JIT_COMPLETE);
    pcre2_match_data_8 *md = pcre2_match_data_create_from_pattern_8(code, NULL);
    PCRE2_SIZE offset = 0;
    int64_t rc_out = 0;
    for (;;) {
        int rc = pcre2_match_8(code, text, n, offset, PCRE2_NO_UTF_CHECK, md, NULL);
        if (rc < 0) {
            if (rc == PCRE2_ERROR_NOMATCH) break;
            rc_out = -2;
            break;
        }
        PCRE2_SIZE *ov = pcre2_get_ovector_pointer_8(md);
        PCRE2_SIZE s = ov[0], e = ov[1];
        if (s == e) { /* :529-533 */
            if (offset >= n) break;
            offset++;
            continue;
        }
        if (starts) {
            if (k >= cap) {
                rc_out = -3;
                break;
            }
            starts[k] = s;
            ends[k] = e;
        }
        k++;
        offset = e;
    }
    pcre2_match_data_free_8(md);
    pcre2_code_free_8(code);
    return rc_out < 0 ? rc_out : (int64_t)k;
}

/* ------------------------------------------------------------------------------------------------
 * oracle_train_rescan: literal restatement of the train loop.
 * tokens/off/weight: flattened chunks (off has n_chunks+1 entries). Working copy is private.
 * merges_out: 2*(vocab_size-256) u32; counts_out (optional): count of the chosen pair when chosen
 * (the 'had C occurrences' figure of Tokenizer.h:576). Returns 0.
 * ---------------------------------------------------------------------------------------------- */
static void count_all(table_t *t, const tok_t *tok, const uint64_t *off, const uint32_t *len, const uint32_t *w,
                      uint64_t n_chunks) {
    /* calculate_freqs, Tokenizer.h:127-146: chunks in order, left to right, +1 per adjacent pair */
    for (uint64_t c = 0; c < n_chunks; c++) {
        const tok_t *p = tok + off[c];
        for (uint32_t i = 0; i + 1 < len[c]; i++) table_add(t, p[i], p[i + 1], (int32_t)w[c]);
    }
}

/* merge (Tokenizer.h:162-199): rewrite only. Returns new length. */
static uint32_t merge_plain(tok_t *p, uint32_t len, tok_t a, tok_t b, tok_t id) {
    uint32_t o = 0, i = 0;
    while (i < len) {
        if (i + 1 < len && p[i] == a && p[i + 1] == b) {
            p[o++] = id;
            i += 2;
        } else
            p[o++] = p[i++];
    }
    return o;
}

/* merge_incremental (Tokenizer.h:202-306): rewrite + <=5 count updates per hit, weighted.
 * x is the CURRENT left neighbour (may be an id written by the previous hit), y the current right one. */
static uint32_t merge_incr(tok_t *p, uint32_t len, tok_t a, tok_t b, tok_t id, table_t *t, int32_t w) {
    if (len < 2) return len; /* :217-220 */
    uint32_t o = 0, i = 0;
    while (i < len) {
        if (i + 1 < len && p[i] == a && p[i + 1] == b) {
            if (table_find(t, ((uint64_t)a << 32) | b) >= 0) table_add(t, a, b, -w); /* :240-246 */
            if (o > 0) {                                                             /* :248-261 */
                tok_t x = p[o - 1];
                if (table_find(t, ((uint64_t)x << 32) | a) >= 0) table_add(t, x, a, -w);
                table_add(t, x, id, w);
            }
            if (i + 2 < len) { /* :263-280 */
                tok_t y = p[i + 2];
                if (table_find(t, ((uint64_t)b << 32) | y) >= 0) table_add(t, b, y, -w);
                table_add(t, id, y, w);
            }
            p[o++] = id;
            i += 2;
        } else
            p[o++] = p[i++];
    }
    return o;
}

int oracle_train_rescan(const uint32_t *tokens, uint64_t n_tokens, const uint64_t *off, uint64_t n_chunks,
                        const uint32_t *weight, uint32_t vocab_size, int mode, uint32_t *merges_out,
                        int32_t *counts_out, uint32_t *n_merges_out) {
    tok_t *tok = (tok_t *)malloc((n_tokens ? n_tokens : 1) * sizeof(tok_t));
    memcpy(tok, tokens, n_tokens * sizeof(tok_t));
    uint32_t *len = (uint32_t *)malloc((n_chunks ? n_chunks : 1) * sizeof(uint32_t));
    for (uint64_t c = 0; c < n_chunks; c++) len[c] = (uint32_t)(off[c + 1] - off[c]);
    table_t t;
    table_init(&t);
    count_all(&t, tok, off, len, weight, n_chunks); /* :552 */
    uint32_t nm = 0;
    for (uint32_t id = 256; id < vocab_size; id++) { /* :557 */
        int64_t top = table_top(&t, mode);           /* :558 */
        if (top < 0) break;                          /* :586-588 */
        tok_t a = (tok_t)(t.e[top].key >> 32), b = (tok_t)t.e[top].key;
        merges_out[2 * nm] = a; /* :578 */
        merges_out[2 * nm + 1] = b;
        if (counts_out) counts_out[nm] = t.e[top].count;
        nm++;
        for (uint64_t c = 0; c < n_chunks; c++) { /* merge_chunks :309-320 */
            if (mode == MODE_FIRST)
                len[c] = merge_plain(tok + off[c], len[c], a, b, id);
            else
                len[c] = merge_incr(tok + off[c], len[c], a, b, id, &t, (int32_t)weight[c]);
        }
        if (mode == MODE_FIRST) { /* :581-585 fresh table, fresh insertion order */
            table_clear(&t);
            count_all(&t, tok, off, len, weight, n_chunks);
        }
    }
    *n_merges_out = nm;
    table_free(&t);
    free(tok);
    free(len);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * oracle_train_indexed: same observable results as oracle_train_rescan, cost O(touched chunks) per merge.
 *
 *  - counts are maintained incrementally in BOTH modes (they are exact counts either way);
 *  - pair -> list of chunk ids that contained it when it was created or counted (append-only, validated on use);
 *  - LEXICAL key  = (count desc, pair asc);  exhaustion: entries are never erased, so when the best count is 0
 *    the smallest pair ever inserted wins, again and again;
 *  - FIRST key    = (count desc, first live occurrence asc) where occurrence position = (chunk index, original
 *    byte offset): equal to the insertion order of a fresh left-to-right recount. A pair whose
 *    count is 0 is absent from a fresh table, so best count 0 => stop (Tokenizer.h:586-588).
 *  - lazy max-heap: every key change pushes a fresh node; a popped node is accepted only if it still equals
 *    the pair's true key.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t count;
    uint64_t tie; /* LEXICAL: pair key; FIRST: first-occurrence position key */
    int64_t ent;
} hnode_t;

typedef struct {
    hnode_t *a;
    uint64_t n, cap;
} heap_t;

static inline int hless(const hnode_t *x, const hnode_t *y) { /* x ranks before y */
    if (x->count != y->count) return x->count > y->count;
    if (x->tie != y->tie) return x->tie < y->tie;
    return x->ent < y->ent;
}
static void heap_push(heap_t *h, hnode_t v) {
    if (h->n == h->cap) {
        h->cap = h->cap ? h->cap * 2 : 1024;
        h->a = (hnode_t *)realloc(h->a, h->cap * sizeof(hnode_t));
    }
    uint64_t i = h->n++;
    while (i > 0) {
        uint64_t p = (i - 1) / 2;
        if (!hless(&v, &h->a[p])) break;
        h->a[i] = h->a[p];
        i = p;
    }
    h->a[i] = v;
}
static hnode_t heap_pop(heap_t *h) {
    hnode_t top = h->a[0], v = h->a[--h->n];
    uint64_t i = 0;
    for (;;) {
        uint64_t l = 2 * i + 1, r = l + 1, m = i;
        const hnode_t *best = &v;
        if (l < h->n && hless(&h->a[l], best)) {
            best = &h->a[l];
            m = l;
        }
        if (r < h->n && hless(&h->a[r], best)) {
            best = &h->a[r];
            m = r;
        }
        if (m == i) break;
        h->a[i] = h->a[m];
        i = m;
    }
    if (h->n) h->a[i] = v;
    return top;
}

typedef struct {
    uint32_t *v;
    uint32_t n, cap, head;
} u32vec_t;
static void vec_push(u32vec_t *v, uint32_t x) {
    if (v->n == v->cap) {
        v->cap = v->cap ? v->cap * 2 : 4;
        v->v = (uint32_t *)realloc(v->v, v->cap * sizeof(uint32_t));
    }
    v->v[v->n++] = x;
}

typedef struct {
    table_t t;
    u32vec_t *where; /* per entry: chunk ids (ascending; duplicates possible) */
    uint64_t cap_where;
    uint8_t *dirty; /* per entry: key changed during the current merge */
    u32vec_t dirty_list;
} index_t;

static int64_t idx_add(index_t *ix, tok_t a, tok_t b, int32_t delta, uint32_t chunk, int record) {
    int64_t e = table_add(&ix->t, a, b, delta);
    if ((uint64_t)e >= ix->cap_where) {
        uint64_t nc = ix->cap_where * 2;
        while ((uint64_t)e >= nc) nc *= 2;
        ix->where = (u32vec_t *)realloc(ix->where, nc * sizeof(u32vec_t));
        memset(ix->where + ix->cap_where, 0, (nc - ix->cap_where) * sizeof(u32vec_t));
        ix->dirty = (uint8_t *)realloc(ix->dirty, nc);
        memset(ix->dirty + ix->cap_where, 0, nc - ix->cap_where);
        ix->cap_where = nc;
    }
    if (record) {
        u32vec_t *w = &ix->where[e];
        if (w->n == 0 || w->v[w->n - 1] != chunk) vec_push(w, chunk);
    }
    if (!ix->dirty[e]) {
        ix->dirty[e] = 1;
        vec_push(&ix->dirty_list, (uint32_t)e);
    }
    return e;
}

/* first live occurrence of entry e's pair: (chunk << 32) | original byte offset, or UINT64_MAX if none.
 * pos[] holds the original offset of each live token (order-preserving under merges). */
static uint64_t first_occurrence(index_t *ix, int64_t e, const tok_t *tok, const uint32_t *pos, const uint64_t *off,
                                 const uint32_t *len) {
    u32vec_t *w = &ix->where[e];
    tok_t a = (tok_t)(ix->t.e[e].key >> 32), b = (tok_t)ix->t.e[e].key;
    while (w->head < w->n) {
        uint32_t c = w->v[w->head];
        const tok_t *p = tok + off[c];
        for (uint32_t i = 0; i + 1 < len[c]; i++)
            if (p[i] == a && p[i + 1] == b) return ((uint64_t)c << 32) | pos[off[c] + i];
        w->head++; /* chunk no longer contains the pair: occurrence sets only shrink after creation */
    }
    return UINT64_MAX;
}

int oracle_train_indexed(const uint32_t *tokens, uint64_t n_tokens, const uint64_t *off, uint64_t n_chunks,
                         const uint32_t *weight, uint32_t vocab_size, int mode, uint32_t *merges_out,
                         int32_t *counts_out, uint32_t *n_merges_out) {
    tok_t *tok = (tok_t *)malloc((n_tokens ? n_tokens : 1) * sizeof(tok_t));
    memcpy(tok, tokens, n_tokens * sizeof(tok_t));
    uint32_t *pos = (uint32_t *)malloc((n_tokens ? n_tokens : 1) * sizeof(uint32_t));
    uint32_t *len = (uint32_t *)malloc((n_chunks ? n_chunks : 1) * sizeof(uint32_t));
    uint32_t *stamp = (uint32_t *)calloc(n_chunks ? n_chunks : 1, sizeof(uint32_t));
    for (uint64_t c = 0; c < n_chunks; c++) {
        len[c] = (uint32_t)(off[c + 1] - off[c]);
        for (uint32_t i = 0; i < len[c]; i++) pos[off[c] + i] = i;
    }
    index_t ix;
    memset(&ix, 0, sizeof ix);
    table_init(&ix.t);
    ix.cap_where = 1024;
    ix.where = (u32vec_t *)calloc(ix.cap_where, sizeof(u32vec_t));
    ix.dirty = (uint8_t *)calloc(ix.cap_where, 1);
    heap_t heap;
    memset(&heap, 0, sizeof heap);

    for (uint64_t c = 0; c < n_chunks; c++) {
        const tok_t *p = tok + off[c];
        for (uint32_t i = 0; i + 1 < len[c]; i++) idx_add(&ix, p[i], p[i + 1], (int32_t)weight[c], (uint32_t)c, 1);
    }
    uint64_t min_key_ever = UINT64_MAX; /* LEXICAL exhaustion */

    uint32_t nm = 0;
    for (uint32_t id = 256; id < vocab_size; id++) {
        /* flush dirty entries into the heap with their true keys */
        for (uint32_t k = 0; k < ix.dirty_list.n; k++) {
            int64_t e = ix.dirty_list.v[k];
            ix.dirty[e] = 0;
            if (ix.t.e[e].key < min_key_ever) min_key_ever = ix.t.e[e].key;
            hnode_t nd;
            nd.count = ix.t.e[e].count;
            nd.ent = e;
            if (mode == MODE_LEXICAL)
                nd.tie = ix.t.e[e].key;
            else {
                if (nd.count <= 0) continue; /* absent from a fresh table */
                nd.tie = first_occurrence(&ix, e, tok, pos, off, len);
            }
            heap_push(&heap, nd);
        }
        ix.dirty_list.n = 0;

        /* pop until a node matches its pair's true key */
        int64_t top = -1;
        int32_t top_count = 0;
        while (heap.n) {
            hnode_t nd = heap_pop(&heap);
            int32_t c = ix.t.e[nd.ent].count;
            if (mode == MODE_LEXICAL) {
                if (c != nd.count) continue; /* stale; a fresh node exists */
                top = nd.ent;
                top_count = c;
                heap_push(&heap, nd); /* entries are never erased */
                break;
            }
            if (c != nd.count || c <= 0) continue;
            uint64_t f = first_occurrence(&ix, nd.ent, tok, pos, off, len);
            if (f != nd.tie) { /* first occurrence moved right: re-queue with the true key */
                nd.tie = f;
                heap_push(&heap, nd);
                continue;
            }
            top = nd.ent;
            top_count = c;
            heap_push(&heap, nd);
            break;
        }
        if (top < 0) break; /* FIRST: empty table (:586-588); LEXICAL: no pair was ever inserted */
        uint64_t key = ix.t.e[top].key;
        if (mode == MODE_LEXICAL && top_count <= 0) key = min_key_ever; /* all counts are 0: smallest key wins */
        tok_t a = (tok_t)(key >> 32), b = (tok_t)key;
        merges_out[2 * nm] = a;
        merges_out[2 * nm + 1] = b;
        if (counts_out) counts_out[nm] = top_count < 0 ? 0 : top_count;
        nm++;
        if (top_count <= 0) continue; /* nothing to rewrite */

        u32vec_t *w = &ix.where[top];
        uint32_t n_where = w->n; /* the vector may be reallocated by idx_add, so index it */
        for (uint32_t k = 0; k < n_where; k++) {
            uint32_t c = ix.where[top].v[k];
            if (stamp[c] == id) continue;
            stamp[c] = id;
            tok_t *p = tok + off[c];
            uint32_t *q = pos + off[c];
            uint32_t L = len[c], o = 0, i = 0;
            int32_t wt = (int32_t)weight[c];
            while (i < L) {
                if (i + 1 < L && p[i] == a && p[i + 1] == b) {
                    idx_add(&ix, a, b, -wt, c, 0);
                    if (o > 0) {
                        tok_t x = p[o - 1];
                        idx_add(&ix, x, a, -wt, c, 0);
                        idx_add(&ix, x, id, wt, c, 1);
                    }
                    if (i + 2 < L) {
                        tok_t y = p[i + 2];
                        idx_add(&ix, b, y, -wt, c, 0);
                        idx_add(&ix, id, y, wt, c, 1);
                    }
                    q[o] = q[i];
                    p[o++] = id;
                    i += 2;
                } else {
                    q[o] = q[i];
                    p[o++] = p[i++];
                }
            }
            len[c] = o;
        }
    }
    *n_merges_out = nm;
    for (uint64_t i = 0; i < ix.cap_where; i++) free(ix.where[i].v);
    free(ix.where);
    free(ix.dirty);
    free(ix.dirty_list.v);
    free(heap.a);
    table_free(&ix.t);
    free(tok);
    free(pos);
    free(len);
    free(stamp);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * oracle_encode: internal_internal_encode + internal_encode + flatten (Tokenizer.h:325-377, :714-717).
 * merges: n_merges pairs, pair i gets id 256+i, later duplicates overwrite (:835).
 * chunks are byte ranges [starts[c], ends[c]) of text; each byte widened to a token (:96-98).
 * out must hold at least sum(chunk lengths) tokens; out_off (optional) gets n_chunks+1 offsets.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    uint64_t *keys;
    uint32_t *vals;
    uint64_t cap;
} lookup_t;

static void lookup_build(lookup_t *lk, const uint32_t *merges, uint32_t n_merges) {
    lk->cap = 16;
    while (lk->cap < 4ull * (n_merges + 1)) lk->cap *= 2;
    lk->keys = (uint64_t *)malloc(lk->cap * sizeof(uint64_t));
    lk->vals = (uint32_t *)malloc(lk->cap * sizeof(uint32_t));
    memset(lk->keys, 0xff, lk->cap * sizeof(uint64_t));
    for (uint32_t i = 0; i < n_merges; i++) {
        uint64_t key = ((uint64_t)merges[2 * i] << 32) | merges[2 * i + 1], h = mix64(key) & (lk->cap - 1);
        while (lk->keys[h] != UINT64_MAX && lk->keys[h] != key) h = (h + 1) & (lk->cap - 1);
        lk->keys[h] = key;
        lk->vals[h] = 256 + i; /* merges_lookup[pair] = id, last writer wins */
    }
}
static inline int lookup_get(const lookup_t *lk, tok_t a, tok_t b, tok_t *id) {
    uint64_t key = ((uint64_t)a << 32) | b, h = mix64(key) & (lk->cap - 1);
    while (lk->keys[h] != UINT64_MAX) {
        if (lk->keys[h] == key) {
            *id = lk->vals[h];
            return 1;
        }
        h = (h + 1) & (lk->cap - 1);
    }
    return 0;
}

int oracle_encode(const uint32_t *merges, uint32_t n_merges, const uint8_t *text, const uint64_t *starts,
                  const uint64_t *ends, uint64_t n_chunks, uint32_t *out, uint64_t *out_off, uint64_t *n_out) {
    lookup_t lk;
    lookup_build(&lk, merges, n_merges);
    uint64_t o = 0;
    for (uint64_t c = 0; c < n_chunks; c++) {
        uint64_t len = ends[c] - starts[c];
        tok_t *p = out + o;
        for (uint64_t i = 0; i < len; i++) p[i] = text[starts[c] + i];
        if (out_off) out_off[c] = o;
        for (;;) { /* one pass per iteration; recursion at :362-366 */
            if (len < 2) break; /* :326-328 */
            uint64_t w = 0, i = 0, merged = 0;
            while (i < len) {
                tok_t id;
                if (i + 1 < len && lookup_get(&lk, p[i], p[i + 1], &id)) {
                    p[w++] = id;
                    i += 2;
                    merged++;
                } else
                    p[w++] = p[i++];
            }
            len = w;
            if (!merged) break;
        }
        o += len;
    }
    if (out_off) out_off[n_chunks] = o;
    *n_out = o;
    free(lk.keys);
    free(lk.vals);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * oracle_decode: Tokenizer.h:725-751. vocab rebuilt from merges as load() does (:844-861).
 * specials: n_special (id, byte string) pairs, checked FIRST (:733-736); ids >= vocab size skipped (:739-742).
 * Call with out==NULL to get the needed size in *n_out.
 * ---------------------------------------------------------------------------------------------- */
int oracle_decode(const uint32_t *merges, uint32_t n_merges, const uint32_t *special_ids,
                  const uint8_t *special_bytes, const uint64_t *special_off, uint32_t n_special, const uint32_t *ids,
                  uint64_t n_ids, uint8_t *out, uint64_t *n_out) {
    uint32_t V = 256 + n_merges;
    uint64_t *voff = (uint64_t *)malloc(((uint64_t)V + 1) * sizeof(uint64_t));
    uint64_t total = 256;
    uint32_t *vlen = (uint32_t *)malloc((uint64_t)V * sizeof(uint32_t));
    for (uint32_t i = 0; i < 256; i++) vlen[i] = 1;
    for (uint32_t i = 0; i < n_merges; i++) {
        uint32_t a = merges[2 * i], b = merges[2 * i + 1];
        if (a >= 256 + i || b >= 256 + i) { /* the reference would index out of range: undefined */
            free(voff);
            free(vlen);
            return -1;
        }
        vlen[256 + i] = vlen[a] + vlen[b];
        total += vlen[256 + i];
    }
    uint8_t *vb = (uint8_t *)malloc(total);
    voff[0] = 0;
    for (uint32_t i = 0; i < V; i++) voff[i + 1] = voff[i] + vlen[i];
    for (uint32_t i = 0; i < 256; i++) vb[i] = (uint8_t)i;
    for (uint32_t i = 0; i < n_merges; i++) {
        uint32_t a = merges[2 * i], b = merges[2 * i + 1];
        memcpy(vb + voff[256 + i], vb + voff[a], vlen[a]);
        memcpy(vb + voff[256 + i] + vlen[a], vb + voff[b], vlen[b]);
    }
    uint64_t o = 0;
    for (uint64_t k = 0; k < n_ids; k++) {
        uint32_t id = ids[k];
        int sp = -1;
        for (uint32_t s = 0; s < n_special; s++)
            if (special_ids[s] == id) sp = (int)s; /* map: a later duplicate id replaced the earlier one */
        if (sp >= 0) {
            uint64_t l = special_off[sp + 1] - special_off[sp];
            if (out) memcpy(out + o, special_bytes + special_off[sp], l);
            o += l;
            continue;
        }
        if (id >= V) continue;
        if (out) memcpy(out + o, vb + voff[id], vlen[id]);
        o += vlen[id];
    }
    *n_out = o;
    free(voff);
    free(vlen);
    free(vb);
    return 0;
}
