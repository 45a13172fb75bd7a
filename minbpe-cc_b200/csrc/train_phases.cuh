// train_phases.cuh -- the BPE train merge loop as barrier-separated phases.
//
// Replaces (file:line under /root/reference/code/include):
//   PairCount<T> + both implementations          PairCount.h:27-47, :101-181, :227-279   -> pair table (Slot[])
//   calculate_freqs                              Tokenizer.h:127-146                     -> initial histogram (train_kernels.cu)
//   get_top_pair_count + comparators             PairCount.h:66-74, :159, :195-207, :262 -> sel_* phases
//   merge / merge_incremental / merge_chunks     Tokenizer.h:162-320                     -> hits / mutate / seg_* phases
//   train driver loop                            Tokenizer.h:557-589                     -> one "step" = sel .. fin
//
// Design (B200-first, not a translation):
//   * The deduplicated corpus lives in HBM as one 16-byte Node per token position {tok, nxt, prv, weight};
//     positions never move, merged-away tokens are unlinked (no compaction, no rescans).
//   * The pair table is open-addressed, one 32-byte Slot per pair (one DRAM sector per probe hit).
//   * Every adjacent-pair occurrence of a pair (p, q) comes into existence in the merge step that creates
//     max(p, q) (or is in the initial text), and after that the set only shrinks. So each pair owns ONE
//     contiguous segment of an occurrence arena, written once; stale entries are filtered on use.
//   * arg-max runs over a small candidate list (all pairs with count >= theta). The best count never
//     increases and new pairs never exceed it, so the list stays complete until its best falls below theta;
//     then a full-grid scan rebuilds it with a lower theta.
//   * first-occurrence tie-break: Slot.first = min live position of the pair (flat position order ==
//     (first-appearance rank of the chunk, offset in chunk) order == insertion order of the reference's fresh
//     recount). It is invalidated when that occurrence dies and recomputed from the pair's segment
//     only when the pair ties for the best count.
//
// Each phase is a function of (ctx, tid, nth): thread `tid` of `nth` handles items tid, tid+nth, ...
// Phases only communicate through memory + atomics, and a barrier separates consecutive phases. The same
// code is driven three ways:  full-grid kernels (one launch per phase), one persistent CTA with
// __syncthreads() between phases, and -- in tests only -- a sequential host loop.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MB_HD __host__ __device__ __forceinline__
#else
#define MB_HD inline
#endif

namespace mbpe {

constexpr uint32_t NIL = 0xFFFFFFFFu;      // no neighbour
constexpr uint32_t DEAD = 0xFFFFFFFFu;     // Node.tok of an unlinked position
constexpr uint64_t EMPTY_KEY = ~0ull;      // free slot
constexpr uint32_t NO_FIRST = 0xFFFFFFFFu; // Slot.first unknown
constexpr int32_t CMAX_NONE = INT32_MIN;
constexpr int HIST_BUCKETS = 256;
// pair-table load factor limit MB_LOAD_NUM / MB_LOAD_DEN: linear probing's TAIL (the slowest of a few hundred
// concurrent probes, each step a dependent DRAM access) sets the step latency, so the table is kept sparse
#ifndef MB_LOAD_NUM
#define MB_LOAD_NUM 3
#define MB_LOAD_DEN 20
#endif

// count -> bucket: floor(log2 v) and the next three mantissa bits (monotone in v); bucket -> its smallest count
MB_HD uint32_t hist_bucket(uint32_t v) {
    uint32_t e = 0;
    for (uint32_t t = v; t > 1; t >>= 1) e++;
    uint32_t m = e >= 3 ? (v >> (e - 3)) & 7u : (v << (3 - e)) & 7u;
    return e * 8 + m;
}
MB_HD uint32_t hist_floor(uint32_t bkt) {
    uint32_t e = bkt >> 3, m = bkt & 7;
    return e >= 3 ? ((8u + m) << (e - 3)) : (((8u + m) >> (3 - e)) + ((((8u + m) & ((1u << (3 - e)) - 1)) != 0) ? 1u : 0u));
}

enum Status : int32_t {
    ST_RUN = 0,          // keep going
    ST_DONE = 1,         // n_target merges done
    ST_EXHAUSTED = 2,    // best count <= 0: FIRST stops (Tokenizer.h:586-588), LEXICAL replays
    ST_NEED_REBUILD = 3, // candidate list no longer holds the best pair
    ST_NEED_GROW = 4,    // pair table too full for the next step
    ST_BIG_MERGE = 5,    // selected pair has more occurrences than one CTA should walk
    ST_FAILED = 6,       // sharded resident kernel: a peer did not answer in time / an exchange buffer overflowed
};

struct Node {
    uint32_t tok, nxt, prv, wt;
};
static_assert(sizeof(Node) == 16, "Node is one 16-byte vector load");

struct Slot {
    uint64_t key;   // (a << 32) | b; numeric order == lexicographic pair order
    int32_t cnt;    // weighted count (PairCount.h:57 'int count')
    uint32_t len;   // occurrences recorded at creation (segment length)
    uint32_t first; // min live position, or NO_FIRST
    uint32_t seg;   // segment start in the arena
    uint32_t fill;  // segment fill cursor
    uint32_t pad;   // index in the candidate list + 1, 0 = not a candidate (kept by seg_alloc / rebuild_collect)
};
static_assert(sizeof(Slot) == 32, "Slot is one 32-byte sector");

// one count delta exchanged between ranks: -(w) for an occurrence that died, +(local count) for a pair born here
struct XRec {
    uint64_t key;
    int32_t delta;
    uint32_t pos; // global position: first occurrence (births) / the occurrence that died (deaths)
};
static_assert(sizeof(XRec) == 16, "exchange record");

struct Ctl {
    // loop state
    uint32_t step;     // merges recorded so far
    uint32_t n_target; // vocab_size - 256
    int32_t mode;      // MBPE_MODE_*
    int32_t status;
    // selection
    int32_t cmax;
    uint32_t n_fix;
    uint64_t best_tie;
    uint32_t best_slot;
    uint32_t best_cand; // resident CTA with the candidate mirror: index of best_slot in the candidate list, else NIL
    uint32_t a, b, new_id;
    uint32_t seg, seg_len;
    int32_t theta;
    uint32_t n_live; // candidates still >= theta, counted by sel_max
    // per-step lists
    uint32_t n_hit, n_rec, n_newp;
    uint32_t n_xrec, n_newp_own; // sharded: records to send; newp entries created by this rank's own occurrences
    uint32_t n_cand;
    uint32_t selected; // 1 between sel_commit and fin: (a, b, seg...) of the current step are valid
    // table / arena
    uint32_t n_pairs;
    uint32_t arena_cursor;
    uint32_t big_limit; // segment length above which the persistent CTA yields
    uint32_t big_count; // sharded resident CTA: best COUNT above which it yields (the count is the same on every rank, a
                        // local segment length is not; live local occurrences <= count bounds the exchange volume); 0 = off
    uint32_t xstep;     // sharded resident CTA: exchanges done so far (tag and buffer parity of the next one)
    uint32_t cand_limit; // candidate-list length that triggers a rebuild ...
    uint32_t cand_base;  // ... unless the list was already that long right after the last rebuild (massive ties)
    uint64_t min_key_ever;
    // rebuild scratch
    int32_t gmax;
    uint32_t n_positive;
    uint32_t hist[HIST_BUCKETS]; // pairs per count bucket: 8 sub-buckets per power of two
    // accounting (SURVEY 8(d) rescan-volume figure)
    uint64_t live_tokens;
    uint64_t rescan_bytes;
    uint64_t n_small_steps;
    // resident-CTA cycle counters (thread 0, clock64): select, hits, mutate+seg_alloc, seg_fill, fin, steps, total
    uint64_t prof[8];
    uint64_t dbg[4];
    uint64_t hh_steps[6], hh_cycles[6], hh_occ[6]; // hits phase by segment length class (debug)
    uint32_t pt_max[6];      // MBPE_PROFILE_HITS: slowest thread of the current step at checkpoints A..F
    uint32_t pf_mode;        // occurrence walk: ask L2 for the sectors around an occurrence while its own node is fetched (0 = off)
    uint32_t pt_pad1;
    uint64_t pt_sum[6][6];   // summed per segment length class
};

struct Ctx {
    Node *node;
    uint32_t n_pos;
    Slot *slot;
    uint32_t cap_mask; // slots - 1
    uint32_t *occ;     // occurrence arena
    uint32_t arena_cap;
    uint32_t *hit;      // positions of this step's merges
    uint32_t *hit_j;    // ... the position of each merge's second token and of the token after it, as the occurrence
    uint32_t *hit_y;    //     walk saw them: the rewrite (mutate) then needs no loads at all
    uint32_t *rec_slot; // this step's new-occurrence records
    uint32_t *rec_pos;
    uint32_t *newp;     // slots created this step
    uint32_t *cand;     // candidate slot list (capacity cand_cap = slots / 2 >= number of pairs)
    uint32_t *fix;      // FIRST mode: tied candidates whose first position must be recomputed (cand_cap)
    uint32_t cand_cap;
    Ctl *ctl;
    uint32_t *merges_out; // device copy, 2 * n_target
    int32_t *counts_out;
    // sharded training (one rank = one GPU = a contiguous share of the unique chunks):
    XRec *xrec;        // this step's count deltas for the other ranks (nullptr = single GPU)
    uint32_t xrec_cap;
    uint32_t pos_base; // global position of local position 0 (first-occurrence keys are global positions)
    // resident CTA: shared-memory mirror of the candidates (index = position in the candidate list), so
    // that the selection reads no global memory at all. m_cnt follows every count change (pair_dec_at finds the
    // index in Slot::pad); key / len / seg never change once a pair's birth step is over. nullptr elsewhere.
    int32_t *m_cnt;
    uint32_t *m_first; // FIRST mode: Slot::first of the candidate (dynamic: dies with its occurrence, recomputed on a tie)
    uint64_t *m_key;
    uint32_t *m_len;
    uint32_t *m_seg;
    uint32_t m_cap;
};

// ---------------------------------------------------------------------------------------------------------
// atomics: CUDA intrinsics on the device; plain sequential updates under the host test driver
// ---------------------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
#define MB_ON_DEVICE 1
#else
#define MB_ON_DEVICE 0
#endif

MB_HD uint32_t a_add(uint32_t *p, uint32_t v) {
#if MB_ON_DEVICE
    return atomicAdd(p, v);
#else
    uint32_t o = *p;
    *p = o + v;
    return o;
#endif
}
MB_HD int32_t a_add(int32_t *p, int32_t v) {
#if MB_ON_DEVICE
    return atomicAdd(p, v);
#else
    int32_t o = *p;
    *p = o + v;
    return o;
#endif
}
MB_HD uint64_t a_add(uint64_t *p, uint64_t v) {
#if MB_ON_DEVICE
    return (uint64_t)atomicAdd((unsigned long long *)p, (unsigned long long)v);
#else
    uint64_t o = *p;
    *p = o + v;
    return o;
#endif
}
MB_HD void a_min(uint32_t *p, uint32_t v) {
#if MB_ON_DEVICE
    atomicMin(p, v);
#else
    if (v < *p) *p = v;
#endif
}
MB_HD void a_min(uint64_t *p, uint64_t v) {
#if MB_ON_DEVICE
    atomicMin((unsigned long long *)p, (unsigned long long)v);
#else
    if (v < *p) *p = v;
#endif
}
MB_HD void a_max(int32_t *p, int32_t v) {
#if MB_ON_DEVICE
    atomicMax(p, v);
#else
    if (v > *p) *p = v;
#endif
}
MB_HD uint64_t a_cas(uint64_t *p, uint64_t expect, uint64_t v) {
#if MB_ON_DEVICE
    return (uint64_t)atomicCAS((unsigned long long *)p, (unsigned long long)expect, (unsigned long long)v);
#else
    uint64_t o = *p;
    if (o == expect) *p = v;
    return o;
#endif
}
template <bool S, class T>
MB_HD T ld_tok(const T *p) { // a field of a corpus node, same rule as ld_node
#if MB_ON_DEVICE
    return S ? *p : __ldcg(p);
#else
    return *p;
#endif
}
// loads of words that other threads update with atomics in an earlier phase: go to L2, not a stale L1 line
template <class T>
MB_HD T ld_l2(const T *p) {
#if MB_ON_DEVICE
    return __ldcg(p);
#else
    return *p;
#endif
}
// Corpus nodes: the resident CTA (S == true) is the only writer while it runs (plain stores from the same SM
// keep its L1 coherent, and L1 is invalidated at every kernel boundary), so it may use L1-cached loads: the
// neighbours of an occurrence sit in the same 128-byte line as the occurrence itself.
template <bool S = false>
MB_HD Node ld_node(const Node *p) {
#if MB_ON_DEVICE
    uint4 v = S ? *reinterpret_cast<const uint4 *>(p) : __ldcg(reinterpret_cast<const uint4 *>(p));
    Node n;
    n.tok = v.x;
    n.nxt = v.y;
    n.prv = v.z;
    n.wt = v.w;
    return n;
#else
    return *p;
#endif
}

// Control block / step lists: in the resident CTA (S == true) they live in shared memory, or in global memory
// written only by this CTA with plain stores, so plain loads are coherent; otherwise they are global words
// updated by atomics from other SMs and must be read at L2.
template <bool S, class T>
MB_HD T ctl_ld(const T *p) {
#if MB_ON_DEVICE
    if (S) return *reinterpret_cast<const volatile T *>(p);
    return __ldcg(p);
#else
    return *p;
#endif
}
// opportunistic warp aggregation of "claim one slot of a list": one atomic per warp instead of one per lane
MB_HD uint32_t claim_one(uint32_t *counter) {
#if MB_ON_DEVICE
    const unsigned m = __activemask();
    const unsigned lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    uint32_t base = 0;
    if ((int)lane == leader) base = atomicAdd(counter, (uint32_t)__popc(m));
    base = __shfl_sync(m, base, leader);
    return base + __popc(m & ((1u << lane) - 1));
#else
    uint32_t o = *counter;
    *counter = o + 1;
    return o;
#endif
}

MB_HD uint64_t pair_key(uint32_t a, uint32_t b) { return ((uint64_t)a << 32) | b; }
MB_HD uint32_t hash_key(uint64_t k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 29;
    k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 32;
    return (uint32_t)k;
}

// find an existing key (exact counts guarantee presence for every decrement)
MB_HD uint32_t slot_find(const Ctx &c, uint64_t key) {
    uint32_t s = hash_key(key) & c.cap_mask;
    for (;;) {
        uint64_t k = ld_l2(&c.slot[s].key);
        if (k == key) return s;
        if (k == EMPTY_KEY) return NIL;
        s = (s + 1) & c.cap_mask;
    }
}
// find or claim. *created is set for the claiming thread only
MB_HD uint32_t slot_upsert(const Ctx &c, uint64_t key, bool *created) {
    uint32_t s = hash_key(key) & c.cap_mask;
    *created = false;
    for (;;) {
        uint64_t k = ld_l2(&c.slot[s].key);
        if (k == key) return s;
        if (k == EMPTY_KEY) {
            uint64_t old = a_cas(&c.slot[s].key, EMPTY_KEY, key);
            if (old == EMPTY_KEY) {
                *created = true;
                return s;
            }
            if (old == key) return s;
        }
        s = (s + 1) & c.cap_mask;
    }
}

// probe continuation: `s` = home slot of `key`, `k0` = the key word already loaded from it. Lets a thread start
// several probes (independent loads in flight) before resolving any of them.
// After the home slot the walk continues FOUR slots per round trip (independent loads, same or adjacent 128-byte
// line), examined in probe order: the result is the one plain linear probing gives, but a chain of length L costs
// 1 + ceil((L - 1) / 4) dependent memory accesses instead of L.
MB_HD uint32_t slot_find_from(const Ctx &c, uint64_t key, uint32_t s, uint64_t k0) {
    if (k0 == key) return s;
    if (k0 == EMPTY_KEY) return NIL;
    for (;;) {
        uint64_t k[4];
        for (int i = 0; i < 4; i++) k[i] = ld_l2(&c.slot[(s + 1 + i) & c.cap_mask].key);
        for (int i = 0; i < 4; i++) {
            if (k[i] == key) return (s + 1 + i) & c.cap_mask;
            if (k[i] == EMPTY_KEY) return NIL;
        }
        s = (s + 4) & c.cap_mask;
    }
}
MB_HD uint32_t slot_upsert_from(const Ctx &c, uint64_t key, uint32_t s, uint64_t k0, bool *created) {
    *created = false;
    for (;;) { // slot s holds k0 (possibly stale if it was EMPTY: the CAS answers that)
        if (k0 == key) return s;
        if (k0 == EMPTY_KEY) {
            uint64_t old = a_cas(&c.slot[s].key, EMPTY_KEY, key);
            if (old == EMPTY_KEY) {
                *created = true;
                return s;
            }
            if (old == key) return s;
        }
        uint64_t k[4];
        for (int i = 0; i < 4; i++) k[i] = ld_l2(&c.slot[(s + 1 + i) & c.cap_mask].key);
        int i = 0;
        for (; i < 4; i++) {
            if (k[i] == key) return (s + 1 + i) & c.cap_mask;
            if (k[i] == EMPTY_KEY) break; // try to claim it: back to the top with this slot
        }
        if (i < 4) {
            s = (s + 1 + i) & c.cap_mask;
            k0 = EMPTY_KEY;
        } else {
            s = (s + 4) & c.cap_mask; // k[3] was neither the key nor empty: continue after it
            k0 = k[3];
        }
    }
}

// upsert when the claim of the home slot (CAS EMPTY -> key) was already issued: `old` is its answer
MB_HD uint32_t upsert_after_claim(const Ctx &c, uint64_t key, uint32_t home, uint64_t first_key, bool tried, uint64_t old,
                                  bool *created) {
    if (!tried) return slot_upsert_from(c, key, home, first_key, created);
    *created = (old == EMPTY_KEY);
    if (old == EMPTY_KEY || old == key) return home;
    const uint32_t s = (home + 1) & c.cap_mask; // somebody else took the home slot with another key
    return slot_upsert_from(c, key, s, ld_l2(&c.slot[s].key), created);
}

// ---------------------------------------------------------------------------------------------------------
// count updates
// ---------------------------------------------------------------------------------------------------------
#define MB_G(field) ctl_ld<S>(&g->field)
#define MB_L(ptr) ctl_ld<S>(ptr) /* step lists: hit, rec_*, newp, cand, fix */

// -(p,q) x w for the occurrence whose first token sits at position pairpos
// (hint_slot, hint_pad): Slot::pad of hint_slot if the caller already loaded it next to the key (NIL: no hint)
MB_HD void pair_dec_at(const Ctx &c, int32_t mode, uint32_t s, uint32_t w, uint32_t pairpos, uint32_t hint_slot = NIL,
                       uint32_t hint_pad = 0);
MB_HD void pair_dec(const Ctx &c, int32_t mode, uint32_t p, uint32_t q, uint32_t w, uint32_t pairpos) {
    // reference: decrement only if present (Tokenizer.h:250-256); with exact counts it always is
    pair_dec_at(c, mode, slot_find(c, pair_key(p, q)), w, pairpos);
}
MB_HD void pair_dec_at(const Ctx &c, int32_t mode, uint32_t s, uint32_t w, uint32_t pairpos, uint32_t hint_slot,
                       uint32_t hint_pad) {
    if (s == NIL) return;
    a_add(&c.slot[s].cnt, -(int32_t)w);
    uint32_t ci = 0; // place in the resident CTA's candidate mirror + 1
    if (c.m_cnt) {
        ci = (s == hint_slot) ? hint_pad : ld_l2(&c.slot[s].pad);
        if (ci > c.m_cap) ci = 0;
        if (ci) a_add(&c.m_cnt[ci - 1], -(int32_t)w);
    }
    if (mode == 0 && ld_l2(&c.slot[s].first) == c.pos_base + pairpos) {
        c.slot[s].first = NO_FIRST;
        if (ci) c.m_first[ci - 1] = NO_FIRST;
    }
    if (c.xrec) { // the other ranks hold the same pair with the same global count: tell them
        uint32_t r = claim_one(&c.ctl->n_xrec);
        if (r < c.xrec_cap) {
            XRec x;
            x.key = ld_l2(&c.slot[s].key);
            x.delta = -(int32_t)w;
            x.pos = c.pos_base + pairpos;
            c.xrec[r] = x;
        }
    }
}
MB_HD void pair_inc_at(const Ctx &c, int32_t mode, uint32_t s, bool created, uint64_t key, uint32_t w, uint32_t pairpos);
// +(p,q) x w for a new occurrence at pairpos; q or p is this step's new id, so the pair is born in this step
MB_HD void pair_inc(const Ctx &c, int32_t mode, uint32_t p, uint32_t q, uint32_t w, uint32_t pairpos) {
    bool created;
    uint64_t key = pair_key(p, q);
    uint32_t s = slot_upsert(c, key, &created);
    pair_inc_at(c, mode, s, created, key, w, pairpos);
}
MB_HD void pair_inc_at(const Ctx &c, int32_t mode, uint32_t s, bool created, uint64_t key, uint32_t w, uint32_t pairpos) {
    if (created) {
        // n_pairs is advanced by n_newp in phase_fin; the all-time smallest key only changes when a smaller one
        // shows up, so the (contended, CAS-loop) 64-bit atomic is behind a plain read
        c.newp[claim_one(&c.ctl->n_newp)] = s;
        // (a stale read can only be too LARGE -- the minimum never grows -- so it never skips a needed update)
        if (key < *reinterpret_cast<const volatile uint64_t *>(&c.ctl->min_key_ever)) a_min(&c.ctl->min_key_ever, key);
    }
    // cnt (low word) += w and len (high word) += 1 in one 64-bit atomic: both only grow during the birth step
    a_add(reinterpret_cast<uint64_t *>(&c.slot[s].cnt), ((uint64_t)1 << 32) | (uint64_t)w);
    if (mode == 0) a_min(&c.slot[s].first, c.pos_base + pairpos);
    uint32_t r = claim_one(&c.ctl->n_rec);
    c.rec_slot[r] = s;
    c.rec_pos[r] = pairpos;
}

// ---------------------------------------------------------------------------------------------------------
// selection: get_top_pair_count. Every sel_* phase is a no-op unless status == ST_RUN and nothing is selected
// yet, so a driver may enqueue them without looking at the status first.
// ---------------------------------------------------------------------------------------------------------
template <bool S>
MB_HD bool selecting(const Ctl *g) { return MB_G(status) == ST_RUN && MB_G(selected) == 0; }

// sel_max: best count among candidates (+ how many are still >= theta)
template <bool S = false>
MB_HD void phase_sel_max(const Ctx &c, uint32_t tid, uint32_t nth) {
    Ctl *g = c.ctl;
    if (!selecting<S>(g)) return;
    uint32_t n = MB_G(n_cand);
    int32_t theta = MB_G(theta), best = CMAX_NONE;
    uint32_t live = 0;
    for (uint32_t i = tid; i < n; i += nth) {
        int32_t v = ld_l2(&c.slot[MB_L(&c.cand[i])].cnt);
        if (v > best) best = v;
        live += (v >= theta);
    }
    if (best != CMAX_NONE) a_max(&g->cmax, best);
    if (live) a_add(&g->n_live, live);
}
// sel_tie: among candidates holding the best count, the smallest tie-break key. FIRST-mode pairs whose first
// position is unknown go to the fix list instead.
template <bool S = false>
MB_HD void phase_sel_tie(const Ctx &c, uint32_t tid, uint32_t nth) {
    Ctl *g = c.ctl;
    if (!selecting<S>(g)) return;
    int32_t cmax = MB_G(cmax);
    if (cmax == CMAX_NONE || cmax < MB_G(theta)) return; // sel_check turns this into ST_NEED_REBUILD
    uint32_t n = MB_G(n_cand);
    int32_t mode = MB_G(mode);
    for (uint32_t i = tid; i < n; i += nth) {
        uint32_t s = MB_L(&c.cand[i]);
        if (ld_l2(&c.slot[s].cnt) != cmax) continue;
        if (mode == 1) {
            a_min(&g->best_tie, ld_l2(&c.slot[s].key)); // PairCount.h:195-207
        } else {
            uint32_t f = ld_l2(&c.slot[s].first);
            if (f == NO_FIRST)
                c.fix[a_add(&g->n_fix, 1u)] = s;
            else
                a_min(&g->best_tie, (uint64_t)f); // PairCount.h:66-74 via SURVEY H1
        }
    }
}
// sel_check (one thread, after sel_tie): candidate list exhausted?
template <bool S = false>
MB_HD void phase_sel_check(const Ctx &c) {
    Ctl *g = c.ctl;
    if (!selecting<S>(g)) return;
    int32_t cmax = MB_G(cmax);
    if (cmax == CMAX_NONE || cmax < MB_G(theta)) g->status = ST_NEED_REBUILD;
}
// sel_fix_scan: recompute Slot.first of every pair on the fix list from its segment (live occurrences only).
// Short segments: one thread per pair. Long segments: all threads stride over the segment together.
MB_HD void fix_scan_range(const Ctx &c, uint32_t s, uint32_t k0, uint32_t stride) {
    uint64_t key = ld_l2(&c.slot[s].key);
    uint32_t p = (uint32_t)(key >> 32), q = (uint32_t)key;
    uint32_t seg = ld_l2(&c.slot[s].seg), len = ld_l2(&c.slot[s].len);
    uint32_t best = NO_FIRST;
    for (uint32_t k = k0; k < len; k += stride) {
        uint32_t pos = ld_l2(&c.occ[seg + k]);
        if (pos >= best) continue;
        Node n0 = ld_node(&c.node[pos]);
        if (n0.tok != p || n0.nxt == NIL) continue;
        if (ld_l2(&c.node[n0.nxt].tok) != q) continue;
        best = pos;
    }
    if (best != NO_FIRST) a_min(&c.slot[s].first, c.pos_base + best);
}
constexpr uint32_t FIX_SOLO_LEN = 128;
template <bool S = false>
MB_HD void phase_sel_fix_scan(const Ctx &c, uint32_t tid, uint32_t nth) {
    Ctl *g = c.ctl;
    if (!selecting<S>(g)) return;
    uint32_t nf = MB_G(n_fix);
    for (uint32_t f = tid; f < nf; f += nth) {
        uint32_t s = MB_L(&c.fix[f]);
        if (ld_l2(&c.slot[s].len) <= FIX_SOLO_LEN) fix_scan_range(c, s, 0, 1);
    }
    for (uint32_t f = 0; f < nf; f++) {
        uint32_t s = MB_L(&c.fix[f]);
        if (ld_l2(&c.slot[s].len) > FIX_SOLO_LEN) fix_scan_range(c, s, tid, nth);
    }
}
template <bool S = false>
MB_HD void phase_sel_fix_tie(const Ctx &c, uint32_t tid, uint32_t nth) {
    Ctl *g = c.ctl;
    if (!selecting<S>(g)) return;
    uint32_t nf = MB_G(n_fix);
    for (uint32_t f = tid; f < nf; f += nth) a_min(&g->best_tie, (uint64_t)ld_l2(&c.slot[MB_L(&c.fix[f])].first));
}
// sel_pick: the unique candidate matching (cmax, best_tie)
template <bool S = false>
MB_HD void phase_sel_pick(const Ctx &c, uint32_t tid, uint32_t nth) {
    Ctl *g = c.ctl;
    if (!selecting<S>(g)) return;
    uint32_t n = MB_G(n_cand);
    int32_t cmax = MB_G(cmax), mode = MB_G(mode);
    uint64_t tie = MB_G(best_tie);
    for (uint32_t i = tid; i < n; i += nth) {
        uint32_t s = MB_L(&c.cand[i]);
        if (ld_l2(&c.slot[s].cnt) != cmax) continue;
        uint64_t t = mode == 1 ? ld_l2(&c.slot[s].key) : (uint64_t)ld_l2(&c.slot[s].first);
        if (t == tie) g->best_slot = s;
    }
}
// sel_commit (one thread): record the merge (Tokenizer.h:578) and decide how the step is executed.
// persistent != 0: called from the resident CTA, which hands segments longer than big_limit to the grid.
template <bool S = false>
MB_HD void phase_sel_commit(const Ctx &c, int persistent) {
    Ctl *g = c.ctl;
    if (!selecting<S>(g)) return;
    uint32_t s = MB_G(best_slot);
    uint64_t key = ld_l2(&c.slot[s].key);
    uint32_t step = MB_G(step), seg_len = ld_l2(&c.slot[s].len);
    g->seg_len = seg_len;
    g->best_cand = NIL; // chosen through the table, not through the resident CTA's candidate mirror
    // every occurrence can create two pairs; keep the load factor under 0.6 after the step
    uint64_t need = (uint64_t)MB_G(n_pairs) + 2ull * seg_len + 64;
    if (need * MB_LOAD_DEN > ((uint64_t)c.cap_mask + 1) * MB_LOAD_NUM) {
        g->status = ST_NEED_GROW; // slots move: the step is re-selected after the rehash
        return;
    }
    g->a = (uint32_t)(key >> 32);
    g->b = (uint32_t)key;
    g->new_id = 256 + step;
    g->seg = ld_l2(&c.slot[s].seg);
    g->seg_len = seg_len;
    c.merges_out[2 * step] = (uint32_t)(key >> 32);
    c.merges_out[2 * step + 1] = (uint32_t)key;
    c.counts_out[step] = MB_G(cmax);
    g->selected = 1;
    if (persistent && (seg_len > MB_G(big_limit) || (MB_G(big_count) && (uint32_t)MB_G(cmax) > MB_G(big_count)))) g->status = ST_BIG_MERGE;
}

// ---------------------------------------------------------------------------------------------------------
// hits: find the live occurrences of (a, b), decide the count deltas from the OLD neighbourhood (read-only on
// the corpus), and record what to rewrite. Net effect == merge_incremental's (Tokenizer.h:239-280): exact
// counts of the rewritten text; -(a,b) itself is folded into "count(a,b) := 0" in phase_fin.
// ---------------------------------------------------------------------------------------------------------
#if defined(MBPE_PROFILE_HITS) && MB_ON_DEVICE
#define MB_PT(i) atomicMax(&g->pt_max[i], (uint32_t)(clock64() - pt_t0))
#else
#define MB_PT(i)
#endif
template <bool S = false>
MB_HD void phase_hits(const Ctx &c, uint32_t tid, uint32_t nth) {
    Ctl *g = c.ctl;
#if defined(MBPE_PROFILE_HITS) && MB_ON_DEVICE
    const long long pt_t0 = clock64();
#endif
    const uint32_t a = MB_G(a), b = MB_G(b), id = MB_G(new_id), seg = MB_G(seg), len = MB_G(seg_len);
    const int32_t mode = MB_G(mode);
    const uint32_t pf = MB_G(pf_mode);
    for (uint32_t k = tid; k < len; k += nth) {
        uint32_t pos = ld_l2(&c.occ[seg + k]);
#if MB_ON_DEVICE
        // The walk reads the occurrence's node, then -- one dependent round trip after the other -- its right neighbour, the
        // neighbours of both, and theirs. They are the next live nodes of the same chunk, a few nodes away: ask L2 for the
        // sectors (2 nodes each) around the occurrence NOW, so that only the first of those round trips goes to HBM.
        if (pf & 255u) {
            const uint32_t pfn = pf & 255u;
            const uint32_t s0 = (pos >> 1) > pfn ? (pos >> 1) - pfn : 0u, s1 = min((pos >> 1) + pfn, (c.n_pos - 1) >> 1);
            for (uint32_t sct = s0; sct <= s1; sct++) asm volatile("prefetch.global.L2 [%0];" ::"l"(&c.node[2 * sct]));
        }
#endif
        Node p = ld_node<S>(&c.node[pos]);
        if (p.tok != a || p.nxt == NIL) continue; // stale record
        uint32_t j = p.nxt;
        Node q = ld_node<S>(&c.node[j]);
        if (q.tok != b) continue;
        MB_PT(0);
        const uint32_t w = p.wt;
        if (a != b) {
            {
                const uint32_t hk = claim_one(&g->n_hit);
                c.hit[hk] = pos;
                c.hit_j[hk] = j;
                c.hit_y[hk] = q.nxt;
            }
            // Gather the whole neighbourhood first, then start all four table probes, then resolve them: the
            // loads of each stage are independent, so the dependent chain is ~7 round trips instead of ~15.
            const bool has_l = p.prv != NIL, has_r = q.nxt != NIL;
            Node x = p, y = q;
            if (has_l) x = ld_node<S>(&c.node[p.prv]);
            if (has_r) y = ld_node<S>(&c.node[q.nxt]);
            const bool need_xx = has_l && x.tok == b && x.prv != NIL;
            const bool need_yy = has_r && y.tok == a && y.nxt != NIL;
            uint32_t xx = DEAD, yy = DEAD;
            if (need_xx) xx = ld_tok<S>(&c.node[x.prv].tok);
            if (need_yy) yy = ld_tok<S>(&c.node[y.nxt].tok);
            MB_PT(1);
            // x is the tail of another occurrence ("abab"): that occurrence's right side covers this gap
            const bool do_l = has_l && !(need_xx && xx == a);
            const bool head = need_yy && yy == b; // y starts another occurrence: the new right pair is (id, id)
            const uint64_t kd1 = pair_key(x.tok, a), ki1 = pair_key(x.tok, id);
            const uint64_t kd2 = pair_key(b, y.tok), ki2 = pair_key(id, head ? id : y.tok);
            const uint32_t hd1 = hash_key(kd1) & c.cap_mask, hi1 = hash_key(ki1) & c.cap_mask;
            const uint32_t hd2 = hash_key(kd2) & c.cap_mask, hi2 = hash_key(ki2) & c.cap_mask;
            uint64_t fd1 = 0, fi1 = 0, fd2 = 0, fi2 = 0;
            uint32_t pd1 = 0, pd2 = 0; // candidate index of the home slots (same sector as the key: no extra latency)
            if (do_l) {
                fd1 = ld_l2(&c.slot[hd1].key);
                fi1 = ld_l2(&c.slot[hi1].key);
                if (c.m_cnt) pd1 = ld_l2(&c.slot[hd1].pad);
            }
            if (has_r) {
                fd2 = ld_l2(&c.slot[hd2].key);
                fi2 = ld_l2(&c.slot[hi2].key);
                if (c.m_cnt) pd2 = ld_l2(&c.slot[hd2].pad);
            }
#if MB_ON_DEVICE
            // A home slot that holds another pair means one more round trip for that lookup (the next four slots), and the
            // four lookups are resolved one after the other: ask L2 for all the continuations now (pf_mode bit 8).
            if (pf & 256u) {
                if (do_l && fd1 != kd1 && fd1 != EMPTY_KEY) asm volatile("prefetch.global.L2 [%0];" ::"l"(&c.slot[(hd1 + 1) & c.cap_mask]));
                if (do_l && fi1 != ki1 && fi1 != EMPTY_KEY) asm volatile("prefetch.global.L2 [%0];" ::"l"(&c.slot[(hi1 + 1) & c.cap_mask]));
                if (has_r && fd2 != kd2 && fd2 != EMPTY_KEY) asm volatile("prefetch.global.L2 [%0];" ::"l"(&c.slot[(hd2 + 1) & c.cap_mask]));
                if (has_r && fi2 != ki2 && fi2 != EMPTY_KEY) asm volatile("prefetch.global.L2 [%0];" ::"l"(&c.slot[(hi2 + 1) & c.cap_mask]));
            }
#endif
            MB_PT(2);
            // both claims of empty home slots go out before either answer is needed
            uint64_t o1 = 0, o2 = 0;
            const bool try1 = do_l && fi1 == EMPTY_KEY, try2 = has_r && fi2 == EMPTY_KEY;
            if (try1) o1 = a_cas(&c.slot[hi1].key, EMPTY_KEY, ki1);
            if (try2) o2 = a_cas(&c.slot[hi2].key, EMPTY_KEY, ki2);
            MB_PT(3);
            if (do_l) {
                bool created;
                pair_dec_at(c, mode, slot_find_from(c, kd1, hd1, fd1), w, p.prv, hd1, pd1);
                uint32_t s1 = upsert_after_claim(c, ki1, hi1, fi1, try1, o1, &created);
                pair_inc_at(c, mode, s1, created, ki1, w, p.prv);
            }
            MB_PT(4);
            if (has_r) {
                bool created;
                pair_dec_at(c, mode, slot_find_from(c, kd2, hd2, fd2), w, j, hd2, pd2);
                uint32_t s2 = upsert_after_claim(c, ki2, hi2, fi2, try2, o2, &created);
                pair_inc_at(c, mode, s2, created, ki2, w, pos);
            }
            MB_PT(5);
        } else {
            // a == b: left-to-right non-overlapping rule (Tokenizer.h:176-191). Only the start of a run of a's
            // acts; it walks its run and merges the 1st, 3rd, 5th... pair.
            if (p.prv != NIL) {
                uint32_t xt = ld_tok<S>(&c.node[p.prv].tok);
                if (xt == a) continue;
                pair_dec(c, mode, xt, a, w, p.prv);
                pair_inc(c, mode, xt, id, w, p.prv);
            }
            uint32_t cur = pos, second = j;
            for (;;) {
                uint32_t r = ld_tok<S>(&c.node[second].nxt);
                {
                    const uint32_t hk = claim_one(&g->n_hit);
                    c.hit[hk] = cur;
                    c.hit_j[hk] = second;
                    c.hit_y[hk] = r;
                }
                if (r == NIL) break;
                Node nr = ld_node<S>(&c.node[r]);
                if (nr.tok != a) { // run ended right after this pair
                    pair_dec(c, mode, a, nr.tok, w, second);
                    pair_inc(c, mode, id, nr.tok, w, cur);
                    break;
                }
                bool head = (nr.nxt != NIL && ld_tok<S>(&c.node[nr.nxt].tok) == a);
                pair_inc(c, mode, id, head ? id : a, w, cur);
                if (!head) break; // one trailing a stays; its right-hand pair is untouched
                cur = r;
                second = nr.nxt;
            }
        }
    }
}

// mutate: rewrite the corpus (merge, Tokenizer.h:182-183). Field-wise stores only: different threads own
// different fields of a shared neighbour node.
template <bool S = false>
MB_HD void phase_mutate(const Ctx &c, uint32_t tid, uint32_t nth) {
    Ctl *g = c.ctl;
    const uint32_t id = MB_G(new_id), n = MB_G(n_hit);
    for (uint32_t h = tid; h < n; h += nth) {
        const uint32_t pos = MB_L(&c.hit[h]), j = MB_L(&c.hit_j[h]), y = MB_L(&c.hit_y[h]); // the corpus is read-only
        c.node[pos].tok = id;                                                                  // between the two phases
        c.node[pos].nxt = y;
        c.node[j].tok = DEAD;
        if (y != NIL) c.node[y].prv = pos;
    }
}

// seg_alloc: give every pair born in this step its arena segment; join the candidate list if it qualifies
template <bool S = false>
MB_HD void phase_seg_alloc(const Ctx &c, uint32_t tid, uint32_t nth) {
    Ctl *g = c.ctl;
    uint32_t n = MB_G(n_newp);
    int32_t theta = MB_G(theta);
    for (uint32_t i = tid; i < n; i += nth) {
        uint32_t s = MB_L(&c.newp[i]);
        const uint32_t len = ld_l2(&c.slot[s].len), seg = a_add(&g->arena_cursor, len);
        const int32_t cnt = ld_l2(&c.slot[s].cnt);
        c.slot[s].seg = seg;
        if (cnt >= theta) {
            uint32_t k = a_add(&g->n_cand, 1u);
            if (k < c.cand_cap) c.cand[k] = s; // cannot overflow: cand_cap >= number of pairs
            c.slot[s].pad = k + 1;
            if (c.m_cnt && k < c.m_cap) { // counts and length of a new pair are final here (the hits phase is over)
                c.m_cnt[k] = cnt;
                c.m_first[k] = ld_l2(&c.slot[s].first);
                c.m_key[k] = ld_l2(&c.slot[s].key);
                c.m_len[k] = len;
                c.m_seg[k] = seg;
            }
        }
    }
}
// seg_fill: scatter this step's occurrence records into their pair's segment
template <bool S = false>
MB_HD void phase_seg_fill(const Ctx &c, uint32_t tid, uint32_t nth) {
    Ctl *g = c.ctl;
    uint32_t n = MB_G(n_rec);
    for (uint32_t r = tid; r < n; r += nth) {
        uint32_t s = MB_L(&c.rec_slot[r]);
        uint32_t k = a_add(&c.slot[s].fill, 1u);
        c.occ[ld_l2(&c.slot[s].seg) + k] = MB_L(&c.rec_pos[r]);
    }
}
// fin (one thread): close the step
template <bool S = false>
MB_HD void phase_fin(const Ctx &c) {
    Ctl *g = c.ctl;
    uint64_t live = MB_G(live_tokens);
    uint32_t n_hit = MB_G(n_hit), n_cand = MB_G(n_cand), step = MB_G(step) + 1;
    g->n_pairs = MB_G(n_pairs) + MB_G(n_newp); // every slot created in this step is on the newp list
    // B_train(m) = 4 T_m + 4 T_m + 4 T_{m+1} + 16 P_m
    g->rescan_bytes = MB_G(rescan_bytes) + 8ull * live + 4ull * (live - n_hit) + 16ull * MB_G(n_pairs);
    g->live_tokens = live - n_hit;
    c.slot[MB_G(best_slot)].cnt = 0; // every occurrence of (a,b) was merged or destroyed
    if (c.m_cnt) { // best_cand: the winner's place in the candidate list when the mirror-based selection chose it
        const uint32_t ci = MB_G(best_cand) != NIL ? MB_G(best_cand) + 1 : ld_l2(&c.slot[MB_G(best_slot)].pad);
        if (ci != 0 && ci <= c.m_cap) c.m_cnt[ci - 1] = 0;
    }
    g->best_cand = NIL;
    g->step = step;
    g->selected = 0;
    g->n_hit = 0;
    g->n_rec = 0;
    g->n_newp = 0;
    g->n_xrec = 0;
    g->n_newp_own = 0;
    g->cmax = CMAX_NONE;
    g->best_tie = ~0ull;
    g->n_fix = 0;
    int32_t st = (step >= MB_G(n_target)) ? ST_DONE : ST_RUN;
    // list is mostly dead weight, or longer than the resident CTA keeps in registers: rebuild (full-grid scan,
    // fresh theta) before the next selection
    if (st == ST_RUN && (n_cand > 2 * MB_G(n_live) + 1024 || n_cand > c.cand_cap ||
                         (n_cand > MB_G(cand_limit) && n_cand > 2 * MB_G(cand_base))))
        st = ST_NEED_REBUILD; // n_cand > cand_cap: appends were dropped, only a table scan restores the list
    g->n_live = 0;
    g->status = st;
}
// reset of the selection scratch when a step is (re-)selected after a rebuild / grow
template <bool S = false>
MB_HD void phase_sel_reset(const Ctx &c) {
    Ctl *g = c.ctl;
    g->cmax = CMAX_NONE;
    g->best_tie = ~0ull;
    g->n_fix = 0;
    g->n_live = 0;
    g->selected = 0;
    g->best_cand = NIL;
    g->cand_base = MB_G(n_cand);
    if (MB_G(status) != ST_EXHAUSTED) g->status = ST_RUN;
}
// a big merge taken over by the grid: same step, status back to RUN
template <bool S = false>
MB_HD void phase_take_big(const Ctx &c) {
    Ctl *g = c.ctl;
    if (MB_G(status) == ST_BIG_MERGE) g->status = ST_RUN;
}

// ---------------------------------------------------------------------------------------------------------
// candidate rebuild (full scans of the table)
// ---------------------------------------------------------------------------------------------------------
template <bool S = false>
MB_HD void phase_rebuild_reset(const Ctx &c) {
    Ctl *g = c.ctl;
    g->gmax = 0;
    g->n_positive = 0;
    for (int i = 0; i < HIST_BUCKETS; i++) g->hist[i] = 0;
}
template <bool S = false>
MB_HD void phase_rebuild_hist(const Ctx &c, uint32_t tid, uint32_t nth) {
    Ctl *g = c.ctl;
    uint32_t cap = c.cap_mask + 1;
    for (uint32_t s = tid; s < cap; s += nth) {
        if (ld_l2(&c.slot[s].key) == EMPTY_KEY) continue;
        int32_t v = ld_l2(&c.slot[s].cnt);
        if (v <= 0) continue;
        a_max(&g->gmax, v);
        a_add(&g->n_positive, 1u);
        a_add(&g->hist[hist_bucket((uint32_t)v)], 1u);
    }
}
// one thread: theta = the largest bucket floor with at least `want` pairs at or above it (or 1)
template <bool S = false>
MB_HD void phase_rebuild_theta(const Ctx &c, uint32_t want) {
    Ctl *g = c.ctl;
    g->n_cand = 0;
    if (MB_G(gmax) <= 0) {
        g->status = ST_EXHAUSTED;
        return;
    }
    uint32_t acc = 0;
    int32_t theta = 1;
    for (int bkt = HIST_BUCKETS - 1; bkt >= 0; bkt--) {
        acc += MB_G(hist[bkt]);
        if (acc >= want) {
            theta = (int32_t)hist_floor((uint32_t)bkt);
            break;
        }
    }
    g->theta = theta < 1 ? 1 : theta;
}
template <bool S = false>
MB_HD void phase_rebuild_collect(const Ctx &c, uint32_t tid, uint32_t nth) {
    Ctl *g = c.ctl;
    if (MB_G(status) == ST_EXHAUSTED) return;
    uint32_t cap = c.cap_mask + 1;
    int32_t theta = MB_G(theta);
    for (uint32_t s = tid; s < cap; s += nth) {
        if (ld_l2(&c.slot[s].key) == EMPTY_KEY) continue;
        uint32_t idx = 0; // Slot::pad = place in the new list + 1, 0 for everything that is not on it
        if (ld_l2(&c.slot[s].cnt) >= theta) {
            const uint32_t k = a_add(&g->n_cand, 1u);
            c.cand[k] = s;
            idx = k + 1;
        }
        if (ld_l2(&c.slot[s].pad) != idx) c.slot[s].pad = idx;
    }
}

// ---------------------------------------------------------------------------------------------------------
// table growth: re-insert every slot of `old` into the (larger, cleared) table of c
// ---------------------------------------------------------------------------------------------------------
MB_HD void phase_rehash(const Ctx &c, const Slot *old, uint32_t old_cap, uint32_t tid, uint32_t nth) {
    for (uint32_t s = tid; s < old_cap; s += nth) {
        Slot v = old[s];
        if (v.key == EMPTY_KEY) continue;
        bool created;
        uint32_t d = slot_upsert(c, v.key, &created);
        c.slot[d].cnt = v.cnt;
        c.slot[d].len = v.len;
        c.slot[d].first = v.first;
        c.slot[d].seg = v.seg;
        c.slot[d].fill = v.fill;
    }
}
MB_HD void phase_clear_slots(Slot *slot, uint32_t cap, uint32_t tid, uint32_t nth) {
    for (uint32_t s = tid; s < cap; s += nth) {
        Slot v;
        v.key = EMPTY_KEY;
        v.cnt = 0;
        v.len = 0;
        v.first = NO_FIRST;
        v.seg = 0;
        v.fill = 0;
        v.pad = 0;
        slot[s] = v;
    }
}

// ---------------------------------------------------------------------------------------------------------
// initial build: calculate_freqs (Tokenizer.h:127-146) + occurrence index
// ---------------------------------------------------------------------------------------------------------
// nodes from the flattened chunk list: chunk c = positions [off[c], off[c+1])
MB_HD uint64_t chunk_of(const uint64_t *off, uint64_t n_chunks, uint64_t i) {
    uint64_t lo = 0, hi = n_chunks; // largest c with off[c] <= i (empty chunks share an offset: take the last)
    while (hi - lo > 1) {
        uint64_t mid = (lo + hi) >> 1;
        if (off[mid] <= i)
            lo = mid;
        else
            hi = mid;
    }
    return lo;
}
MB_HD void phase_init_nodes(const Ctx &c, const uint32_t *tokens, const uint64_t *off, const uint32_t *weight,
                            uint64_t n_chunks, uint32_t tid, uint32_t nth) {
    for (uint32_t i = tid; i < c.n_pos; i += nth) {
        uint64_t ch = chunk_of(off, n_chunks, i);
        Node n;
        n.tok = tokens[i];
        n.prv = (i == off[ch]) ? NIL : i - 1;
        n.nxt = (i + 1 == off[ch + 1]) ? NIL : i + 1;
        n.wt = weight ? weight[ch] : 1u;
        c.node[i] = n;
    }
}
// count of one pair: w = summed weight, n_occ = occurrences, first = smallest position (the CUDA build
// pre-aggregates these per CTA in shared memory before touching the table in HBM)
MB_HD void count_one(const Ctx &c, uint64_t key, uint32_t w, uint32_t n_occ, uint32_t first) {
    bool created;
    uint32_t s = slot_upsert(c, key, &created);
    if (created) {
        a_add(&c.ctl->n_pairs, 1u);
        a_min(&c.ctl->min_key_ever, key);
    }
    a_add(reinterpret_cast<uint64_t *>(&c.slot[s].cnt), ((uint64_t)n_occ << 32) | (uint64_t)w);
    a_min(&c.slot[s].first, first);
}
MB_HD void phase_init_count(const Ctx &c, uint32_t tid, uint32_t nth) {
    for (uint32_t i = tid; i < c.n_pos; i += nth) {
        Node n = c.node[i];
        if (n.nxt == NIL) continue;
        count_one(c, pair_key(n.tok, c.node[n.nxt].tok), n.wt, 1u, c.pos_base + i);
    }
}
MB_HD void phase_init_alloc(const Ctx &c, uint32_t tid, uint32_t nth) {
    uint32_t cap = c.cap_mask + 1;
    for (uint32_t s = tid; s < cap; s += nth) {
        if (c.slot[s].key == EMPTY_KEY) continue;
        c.slot[s].seg = a_add(&c.ctl->arena_cursor, c.slot[s].len);
    }
}
MB_HD void phase_init_fill(const Ctx &c, uint32_t tid, uint32_t nth) {
    for (uint32_t i = tid; i < c.n_pos; i += nth) {
        Node n = c.node[i];
        if (n.nxt == NIL) continue;
        uint32_t s = slot_find(c, pair_key(n.tok, c.node[n.nxt].tok));
        uint32_t k = a_add(&c.slot[s].fill, 1u);
        c.occ[c.slot[s].seg + k] = i;
    }
}

} // namespace mbpe

// ---------------------------------------------------------------------------------------------------------
// sharded training: every rank keeps the WHOLE pair table with GLOBAL counts (so every rank selects the same
// pair with no broadcast) and only its own share of the corpus. After the local occurrence walk each rank
// sends its count deltas; applying everybody else's deltas makes the replicas equal again.
// ---------------------------------------------------------------------------------------------------------
// births: one record per pair created by this rank's occurrences in this step, with its local count so far
namespace mbpe {
template <bool S = false>
MB_HD void phase_export_births(const Ctx &c, uint32_t tid, uint32_t nth) {
    Ctl *g = c.ctl;
    const uint32_t n = MB_G(n_newp);
    for (uint32_t i = tid; i < n; i += nth) {
        const uint32_t s = ld_l2(&c.newp[i]);
        uint32_t r = a_add(&g->n_xrec, 1u);
        if (r < c.xrec_cap) {
            XRec x;
            x.key = ld_l2(&c.slot[s].key);
            x.delta = ld_l2(&c.slot[s].cnt);
            x.pos = ld_l2(&c.slot[s].first);
            c.xrec[r] = x;
        }
    }
    if (tid == 0) g->n_newp_own = n;
}
// FIRST mode: after the local sel_fix_scan every rank knows ITS first live occurrence of each tied pair; the
// global first is the minimum over ranks (positions are global)
template <bool S = false>
MB_HD void phase_export_fix(const Ctx &c, uint32_t tid, uint32_t nth) {
    Ctl *g = c.ctl;
    const uint32_t n = MB_G(n_fix);
    for (uint32_t i = tid; i < n; i += nth) {
        const uint32_t s = ld_l2(&c.fix[i]);
        XRec x;
        x.key = ld_l2(&c.slot[s].key);
        x.delta = 0;
        x.pos = ld_l2(&c.slot[s].first);
        c.xrec[i] = x;
    }
    if (tid == 0) g->n_xrec = n;
}
// initial histogram: every occupied slot is a birth
MB_HD void phase_export_all(const Ctx &c, uint32_t tid, uint32_t nth) {
    Ctl *g = c.ctl;
    const uint32_t cap = c.cap_mask + 1;
    for (uint32_t s = tid; s < cap; s += nth) {
        const uint64_t key = c.slot[s].key;
        if (key == EMPTY_KEY) continue;
        uint32_t r = a_add(&g->n_xrec, 1u);
        if (r < c.xrec_cap) {
            XRec x;
            x.key = key;
            x.delta = c.slot[s].cnt;
            x.pos = c.slot[s].first;
            c.xrec[r] = x;
        }
    }
}
// one thread, after the initial exchange: slots created by foreign records count as pairs; lists start empty
MB_HD void phase_after_init_exchange(const Ctx &c) {
    Ctl *g = c.ctl;
    g->n_pairs = ld_l2(&g->n_pairs) + ld_l2(&g->n_newp);
    g->n_newp = 0;
    g->n_xrec = 0;
    g->n_newp_own = 0;
}
// a record another rank wrote into this rank's memory while this kernel runs (resident sharded CTA: peer stores over
// NVLink land in L2; an L1 line from the exchange before last must not answer)
MB_HD XRec ld_xrec(const XRec *p) {
#if MB_ON_DEVICE
    const uint4 v = __ldcv(reinterpret_cast<const uint4 *>(p));
    XRec x;
    x.key = ((uint64_t)v.y << 32) | v.x;
    x.delta = (int32_t)v.z;
    x.pos = v.w;
    return x;
#else
    return *p;
#endif
}
// all[r * stride .. + counts[r]) = records of rank r; the own block is skipped
template <bool S = false>
MB_HD void phase_apply_foreign(const Ctx &c, const XRec *all, const uint32_t *counts, uint32_t stride, uint32_t world,
                               uint32_t my_rank, uint32_t tid, uint32_t nth) {
    Ctl *g = c.ctl;
    const int32_t mode = MB_G(mode);
    for (uint32_t r = 0; r < world; r++) {
        if (r == my_rank) continue;
        const uint32_t n = counts[r];
        for (uint32_t i = tid; i < n; i += nth) {
            const XRec x = ld_xrec(&all[(uint64_t)r * stride + i]);
            bool created;
            const uint32_t s = slot_upsert(c, x.key, &created);
            if (created) {
                c.newp[a_add(&g->n_newp, 1u)] = s; // gets an (empty) segment and a candidate check like local births
                a_min(&g->min_key_ever, x.key);
            }
            if (x.delta) {
                a_add(&c.slot[s].cnt, x.delta);
                if (c.m_cnt) { // resident CTA with the candidate mirror: the count it selects from follows (Slot::pad = place + 1)
                    uint32_t ci = ld_l2(&c.slot[s].pad);
                    if (ci > c.m_cap) ci = 0;
                    if (ci) a_add(&c.m_cnt[ci - 1], x.delta);
                }
            }
            if (mode == 0) {
                if (x.delta >= 0) // birth, or (delta == 0) a rank's recomputed first live occurrence
                    a_min(&c.slot[s].first, x.pos);
                else if (ld_l2(&c.slot[s].first) == x.pos)
                    c.slot[s].first = NO_FIRST;
            }
        }
    }
}
} // namespace mbpe

// ---------------------------------------------------------------------------------------------------------
// phase functors: what a backend launches (kernel names in ncu read k_par<mbpe::PhHits> etc.)
// ---------------------------------------------------------------------------------------------------------
namespace mbpe {
#define MB_PHASE_PAR(NAME, FN)                                                          \
    template <bool S = false>                                                           \
    struct NAME##T {                                                                    \
        Ctx c;                                                                          \
        MB_HD void operator()(uint32_t tid, uint32_t nth) const { FN<S>(c, tid, nth); } \
    };                                                                                  \
    using NAME = NAME##T<false>;
#define MB_PHASE_ONE(NAME, FN)                           \
    template <bool S = false>                            \
    struct NAME##T {                                     \
        Ctx c;                                           \
        MB_HD void operator()() const { FN<S>(c); }      \
    };                                                   \
    using NAME = NAME##T<false>;
MB_PHASE_PAR(PhSelMax, phase_sel_max)
MB_PHASE_PAR(PhSelTie, phase_sel_tie)
MB_PHASE_ONE(PhSelCheck, phase_sel_check)
MB_PHASE_PAR(PhSelFixScan, phase_sel_fix_scan)
MB_PHASE_PAR(PhSelFixTie, phase_sel_fix_tie)
MB_PHASE_PAR(PhSelPick, phase_sel_pick)
MB_PHASE_PAR(PhHits, phase_hits)
MB_PHASE_PAR(PhMutate, phase_mutate)
MB_PHASE_PAR(PhSegAlloc, phase_seg_alloc)
MB_PHASE_PAR(PhSegFill, phase_seg_fill)
MB_PHASE_ONE(PhFin, phase_fin)
MB_PHASE_ONE(PhSelReset, phase_sel_reset)
MB_PHASE_ONE(PhTakeBig, phase_take_big)
MB_PHASE_ONE(PhRebuildReset, phase_rebuild_reset)
MB_PHASE_PAR(PhRebuildHist, phase_rebuild_hist)
MB_PHASE_PAR(PhRebuildCollect, phase_rebuild_collect)
struct PhInitCount { Ctx c; MB_HD void operator()(uint32_t tid, uint32_t nth) const { phase_init_count(c, tid, nth); } };
struct PhInitAlloc { Ctx c; MB_HD void operator()(uint32_t tid, uint32_t nth) const { phase_init_alloc(c, tid, nth); } };
struct PhInitFill { Ctx c; MB_HD void operator()(uint32_t tid, uint32_t nth) const { phase_init_fill(c, tid, nth); } };
template <bool S = false>
struct PhSelCommitT {
    Ctx c;
    int persistent;
    MB_HD void operator()() const { phase_sel_commit<S>(c, persistent); }
};
using PhSelCommit = PhSelCommitT<false>;
struct PhRebuildTheta {
    Ctx c;
    uint32_t want;
    MB_HD void operator()() const { phase_rebuild_theta<false>(c, want); }
};
MB_PHASE_PAR(PhExportBirths, phase_export_births)
MB_PHASE_PAR(PhExportFix, phase_export_fix)
struct PhResetXrec {
    Ctx c;
    MB_HD void operator()() const { c.ctl->n_xrec = 0; }
};
struct PhExportAll {
    Ctx c;
    MB_HD void operator()(uint32_t tid, uint32_t nth) const { phase_export_all(c, tid, nth); }
};
struct PhAfterInitExchange {
    Ctx c;
    MB_HD void operator()() const { phase_after_init_exchange(c); }
};
template <bool S = false>
struct PhApplyForeignT {
    Ctx c;
    const XRec *all;
    const uint32_t *counts;
    uint32_t stride, world, my_rank;
    MB_HD void operator()(uint32_t tid, uint32_t nth) const {
        phase_apply_foreign<S>(c, all, counts, stride, world, my_rank, tid, nth);
    }
};
using PhApplyForeign = PhApplyForeignT<false>;
struct PhRehash {
    Ctx c;
    const Slot *old;
    uint32_t old_cap;
    MB_HD void operator()(uint32_t tid, uint32_t nth) const { phase_rehash(c, old, old_cap, tid, nth); }
};
struct PhClearSlots {
    Slot *slot;
    uint32_t cap;
    MB_HD void operator()(uint32_t tid, uint32_t nth) const { phase_clear_slots(slot, cap, tid, nth); }
};
struct PhInitNodes {
    Ctx c;
    const uint32_t *tokens;
    const uint64_t *off;
    const uint32_t *weight;
    uint64_t n_chunks;
    MB_HD void operator()(uint32_t tid, uint32_t nth) const { phase_init_nodes(c, tokens, off, weight, n_chunks, tid, nth); }
};

// The resident program: steps back to back until something needs the grid (status != ST_RUN).
// Exec supplies the barrier: on the device  par(f) = f(threadIdx.x, blockDim.x); __syncthreads();
// under the host test driver it is a sequential loop over tid.
template <bool S, class Exec>
MB_HD void persistent_program(const Ctx &c, Exec &ex) {
    for (;;) {
        if (ex.load(&c.ctl->status) != ST_RUN) return;
        if (ex.load(&c.ctl->selected) == 0) {
            ex.par(PhSelMaxT<S>{c});
            ex.par(PhSelTieT<S>{c});
            ex.one(PhSelCheckT<S>{c});
            if (ex.load(&c.ctl->n_fix) != 0) {
                ex.par(PhSelFixScanT<S>{c});
                ex.par(PhSelFixTieT<S>{c});
            }
            ex.par(PhSelPickT<S>{c});
            ex.one(PhSelCommitT<S>{c, 1});
            if (ex.load(&c.ctl->status) != ST_RUN) return;
        }
        ex.par(PhHitsT<S>{c});
        ex.par2(PhMutateT<S>{c}, PhSegAllocT<S>{c}); // disjoint data: corpus nodes vs. new slots
        ex.par(PhSegFillT<S>{c});
        ex.one(PhFinT<S>{c});
    }
}

// The resident program of SHARDED training: like persistent_program, plus one exchange of count deltas per merge
// (and one more in FIRST mode when a tied pair lost its first occurrence). Exec also supplies
//   bool select(c)           optional fast selection (device: fused_select in LEXICAL mode); false = not done
//   bool exchange_apply(c)   send c.xrec[0 .. n_xrec) to every other rank, receive theirs, apply them (phase_apply_foreign);
//                            false = failed (status is ST_FAILED)
// Every rank runs the same sequence of steps and leaves the program at the same step: all decisions below are
// functions of the replicated pair table only (never of a rank's local segment lengths).
template <bool S, class Exec>
MB_HD void persistent_program_sharded(const Ctx &c, Exec &ex) {
    for (;;) {
        if (ex.load(&c.ctl->status) != ST_RUN) return;
        if (ex.load(&c.ctl->selected) == 0) {
            if (!ex.select(c)) {
                ex.par(PhSelMaxT<S>{c});
                ex.par(PhSelTieT<S>{c});
                ex.one(PhSelCheckT<S>{c});
                if (ex.load(&c.ctl->status) == ST_RUN && ex.load(&c.ctl->n_fix) != 0) {
                    ex.par(PhSelFixScanT<S>{c});
                    ex.par(PhExportFixT<S>{c});
                    if (!ex.exchange_apply(c)) return;
                    ex.one(PhResetXrec{c});
                    ex.par(PhSelFixTieT<S>{c});
                }
                ex.par(PhSelPickT<S>{c});
                ex.one(PhSelCommitT<S>{c, 1});
            }
            if (ex.load(&c.ctl->status) != ST_RUN) return;
        }
        ex.par(PhHitsT<S>{c});
        ex.par(PhExportBirthsT<S>{c});
        if (!ex.exchange_apply(c)) return;
        ex.par2(PhMutateT<S>{c}, PhSegAllocT<S>{c});
        ex.par(PhSegFillT<S>{c});
        ex.one(PhFinT<S>{c});
    }
}
} // namespace mbpe
