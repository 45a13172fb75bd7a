// train_driver.hpp -- host-side orchestration of the train merge loop (Tokenizer.h:551-589), written once
// against a small backend interface. The product backend (train_cuda.cu) launches the phases as sm_100a
// kernels; tests/emu/ has a sequential host backend used ONLY by the CPU unit tests to check the phase
// logic and this orchestration against the oracle. There is no CPU execution path in the product library.
//
// Backend interface:
//   void *alloc(size_t bytes);  void release(void *p);
//   void upload(void *dst, const void *src, size_t n);  void download(void *dst, const void *src, size_t n);  // sync
//   template <class F> void par(const F &f, uint64_t n_items);   // f(tid, nth) over enough threads for n_items
//   template <class F> void one(const F &f);                     // f() on one thread
//   void init_count(const Ctx &c);                                // calculate_freqs: fill the pair table
//   void persistent(const Ctx &c);                                // persistent_program until status != ST_RUN
//   uint64_t launches() const;
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "train_phases.cuh"

namespace mbpe {

struct TrainConfig {
    uint32_t vocab_size;
    int mode;            // 0 first, 1 lexical
    int engine;          // 0 stepwise, 1 persistent
    uint32_t big_limit;  // persistent CTA yields merges with longer segments to the grid
    uint32_t cand_want;  // candidate list target size at a rebuild
    uint32_t cand_limit; // rebuild when the candidate list grows past this (0 = 4096)
    uint32_t init_slots; // initial pair-table capacity (power of two), 0 = sized for 65536 byte pairs
    uint32_t pf_mode;    // occurrence walk: sectors on either side of an occurrence asked into L2 ahead of the walk (0 = off)
};

struct TrainOutcome {
    uint32_t n_merges;   // merges recorded on the device before stop
    int32_t final_status;
    uint64_t min_key_ever;
    uint64_t n_pairs, table_slots, n_big, n_rebuilds, n_grows, rescan_bytes;
    uint64_t prof[8];
};

inline uint32_t next_pow2_u32(uint64_t v) {
    uint32_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

template <class BE>
class TrainLoop {
  public:
    TrainLoop(BE &be) : be_(be) { memset(&c_, 0, sizeof c_); }
    ~TrainLoop() { free_all(); }

    // d_tokens/d_off/d_weight: resident inputs (weight may be null). h_merges: 2*(vocab-256), h_counts: vocab-256.
    int run(const uint32_t *d_tokens, const uint64_t *d_off, const uint32_t *d_weight, uint64_t n_tokens,
            uint64_t n_chunks, const TrainConfig &cfg, uint32_t *h_merges, int32_t *h_counts, TrainOutcome *out) {
        const uint32_t n_target = cfg.vocab_size - 256;
        memset(out, 0, sizeof *out);
        if (n_target == 0) return 0;
        c_.n_pos = (uint32_t)n_tokens;
        const uint64_t np = n_tokens ? n_tokens : 1;
        c_.node = (Node *)be_.alloc(np * sizeof(Node));
        c_.arena_cap = (uint32_t)(3 * np + 16);
        c_.occ = (uint32_t *)be_.alloc((uint64_t)c_.arena_cap * 4);
        c_.hit = (uint32_t *)be_.alloc((np + 16) * 4);
        c_.hit_j = (uint32_t *)be_.alloc((np + 16) * 4);
        c_.hit_y = (uint32_t *)be_.alloc((np + 16) * 4);
        c_.rec_slot = (uint32_t *)be_.alloc((np + 16) * 4);
        c_.rec_pos = (uint32_t *)be_.alloc((np + 16) * 4);
        c_.newp = (uint32_t *)be_.alloc((np + 16) * 4);
        c_.ctl = (Ctl *)be_.alloc(sizeof(Ctl));
        c_.merges_out = (uint32_t *)be_.alloc((uint64_t)n_target * 8);
        c_.counts_out = (int32_t *)be_.alloc((uint64_t)n_target * 4);
        // initial pairs are byte pairs (tokens < 256 inside multi-token chunks): at most 65536 distinct
        uint64_t init_pairs = n_tokens < 65536 ? n_tokens : 65536;
        alloc_table(cfg.init_slots ? next_pow2_u32(cfg.init_slots) : next_pow2_u32(4 * init_pairs + 1024));

        Ctl h;
        memset(&h, 0, sizeof h);
        h.n_target = n_target;
        h.mode = cfg.mode;
        h.status = ST_NEED_REBUILD;
        h.cmax = CMAX_NONE;
        h.best_tie = ~0ull;
        h.best_cand = NIL;
        h.theta = 1;
        h.big_limit = cfg.big_limit;
        h.pf_mode = cfg.pf_mode;
        h.cand_limit = cfg.cand_limit ? cfg.cand_limit : 4096;
        h.min_key_ever = ~0ull;
        h.live_tokens = n_tokens;
        be_.upload(c_.ctl, &h, sizeof h);

        be_.par(PhInitNodes{c_, d_tokens, d_off, d_weight, n_chunks}, n_tokens);
        be_.init_count(c_);
        be_.par(PhInitAlloc{c_}, (uint64_t)c_.cap_mask + 1);
        be_.par(PhInitFill{c_}, n_tokens);

        for (;;) {
            be_.download(&h, c_.ctl, sizeof h);
            if (h.status == ST_DONE || h.status == ST_EXHAUSTED) break;
            switch (h.status) {
            case ST_NEED_REBUILD:
                rebuild(cfg);
                break;
            case ST_NEED_GROW:
                grow(h);
                rebuild(cfg);
                break;
            case ST_BIG_MERGE:
                be_.one(PhTakeBig{c_});
                apply_grid(h.seg_len);
                select_grid(1, h.n_cand);
                out->n_big++;
                break;
            case ST_RUN:
                if (cfg.engine == 1) {
                    be_.persistent(c_);
                } else if (!h.selected) {
                    select_grid(0, h.n_cand);
                } else {
                    apply_grid(h.seg_len);
                    select_grid(0, h.n_cand);
                }
                break;
            default:
                return -1;
            }
        }
        out->n_merges = h.step;
        out->final_status = h.status;
        out->min_key_ever = h.min_key_ever;
        out->n_pairs = h.n_pairs;
        out->table_slots = (uint64_t)c_.cap_mask + 1;
        out->n_rebuilds = n_rebuilds_;
        out->n_grows = n_grows_;
        out->rescan_bytes = h.rescan_bytes;
        for (int i = 0; i < 8; i++) out->prof[i] = h.prof[i];
        if (getenv("MBPE_DEBUG")) {
            const char *names[6] = {"<=32", "<=256", "<=1024", "<=2048", "<=8192", ">8192"};
            for (int i = 0; i < 6; i++)
                fprintf(stderr, "[mbpe] hits seg_len %-7s steps %8llu  cycles %12llu (%7.0f/step)  occurrences %10llu\n", names[i],
                        (unsigned long long)h.hh_steps[i], (unsigned long long)h.hh_cycles[i],
                        h.hh_steps[i] ? (double)h.hh_cycles[i] / h.hh_steps[i] : 0.0, (unsigned long long)h.hh_occ[i]);
            for (int i = 0; i < 6; i++)
                if (h.pt_sum[i][5])
                    fprintf(stderr, "[mbpe]   slowest thread, avg cycles at checkpoints (nodes, neighbours, probes out, claims out, left done, right done) %-7s: %.0f %.0f %.0f %.0f %.0f %.0f\n",
                            names[i], (double)h.pt_sum[i][0] / h.hh_steps[i], (double)h.pt_sum[i][1] / h.hh_steps[i],
                            (double)h.pt_sum[i][2] / h.hh_steps[i], (double)h.pt_sum[i][3] / h.hh_steps[i],
                            (double)h.pt_sum[i][4] / h.hh_steps[i], (double)h.pt_sum[i][5] / h.hh_steps[i]);
        }
        if (getenv("MBPE_DEBUG")) fprintf(stderr, "[mbpe] select dbg: sum_ncand=%llu generic_steps=%llu max_ncand=%llu max_select_cycles=%llu slow_selects=%llu\n", (unsigned long long)h.prof[7], (unsigned long long)h.dbg[0], (unsigned long long)h.dbg[1], (unsigned long long)h.dbg[2], (unsigned long long)h.dbg[3]);
        if (h.step) {
            be_.download(h_merges, c_.merges_out, (uint64_t)h.step * 8);
            if (h_counts) be_.download(h_counts, c_.counts_out, (uint64_t)h.step * 4);
        }
        free_all();
        return 0;
    }

  private:
    void alloc_table(uint32_t cap) {
        c_.slot = (Slot *)be_.alloc((uint64_t)cap * sizeof(Slot));
        c_.cap_mask = cap - 1;
        c_.cand_cap = cap; // >= number of pairs at any load factor
        c_.cand = (uint32_t *)be_.alloc((uint64_t)c_.cand_cap * 4);
        c_.fix = (uint32_t *)be_.alloc((uint64_t)c_.cand_cap * 4);
        be_.par(PhClearSlots{c_.slot, cap}, cap);
    }
    void select_grid(int persistent, uint32_t n_cand) {
        be_.par(PhSelMax{c_}, n_cand);
        be_.par(PhSelTie{c_}, n_cand);
        be_.one(PhSelCheck{c_});
        be_.par(PhSelFixScan{c_}, n_cand);
        be_.par(PhSelFixTie{c_}, n_cand);
        be_.par(PhSelPick{c_}, n_cand);
        be_.one(PhSelCommit{c_, persistent});
    }
    void apply_grid(uint32_t seg_len) {
        be_.par(PhHits{c_}, seg_len);
        be_.par(PhMutate{c_}, seg_len);
        be_.par(PhSegAlloc{c_}, 2ull * seg_len);
        be_.par(PhSegFill{c_}, 2ull * seg_len);
        be_.one(PhFin{c_});
    }
    void rebuild(const TrainConfig &cfg) {
        uint64_t cap = (uint64_t)c_.cap_mask + 1;
        be_.one(PhRebuildReset{c_});
        be_.par(PhRebuildHist{c_}, cap);
        be_.one(PhRebuildTheta{c_, cfg.cand_want});
        be_.par(PhRebuildCollect{c_}, cap);
        be_.one(PhSelReset{c_});
        n_rebuilds_++;
    }
    void grow(const Ctl &h) {
        Slot *old = c_.slot;
        uint32_t old_cap = c_.cap_mask + 1;
        uint64_t need = (uint64_t)h.n_pairs + 2ull * h.seg_len + 64; // same bound as phase_sel_commit
        uint64_t cap = (uint64_t)old_cap * 2; // stay as small as the load limit allows: L2 residency matters
        while (need * MB_LOAD_DEN > cap * MB_LOAD_NUM) cap *= 2;
        be_.release(c_.cand);
        be_.release(c_.fix);
        alloc_table((uint32_t)cap);
        be_.par(PhRehash{c_, old, old_cap}, old_cap);
        be_.release(old);
        n_grows_++;
    }
    void free_all() {
        void *ps[] = {c_.node, c_.occ, c_.hit, c_.hit_j, c_.hit_y, c_.rec_slot, c_.rec_pos, c_.newp, c_.ctl,
                      c_.merges_out, c_.counts_out, c_.slot, c_.cand, c_.fix};
        for (void *p : ps)
            if (p) be_.release(p);
        memset(&c_, 0, sizeof c_);
    }

    BE &be_;
    Ctx c_;
    uint64_t n_rebuilds_ = 0, n_grows_ = 0;
};

// ---------------------------------------------------------------------------------------------------------
// Sharded training: rank r of `world` owns a contiguous share of the unique chunks (local position
// i = global position pos_base + i) and a full replica of the pair table with GLOBAL counts. One step:
//   select (every rank, same answer)  ->  local occurrence walk + count deltas  ->  exchange of the deltas
//   (backend: NCCL all-gather over NVLink)  ->  apply the other ranks' deltas  ->  local rewrite + segments.
// Extra backend interface:
//   uint32_t world(), rank();
//   void exchange(const XRec *d_send, uint32_t n_send, const XRec **d_all, const uint32_t **d_counts, uint32_t *stride);
//   uint32_t resident_limit();            // 0: no resident program; else the largest best COUNT it takes (Ctl::big_count)
//   void persistent_sharded(const Ctx &c); // persistent_program_sharded until status != ST_RUN
// With engine 1 and a backend that has it, merges whose count is at most resident_limit() run in ONE resident CTA per
// rank that exchanges the deltas itself (CUDA: stores into the peers' memory over NVLink, no launch and no host
// round trip per merge); bigger merges, rebuilds and everything under engine 0 are driven from the host, one grid
// kernel per phase and one all-gather per exchange.
// The table is sized once for the worst case (no growth: the ranks' step sequences must stay identical).
// ---------------------------------------------------------------------------------------------------------
template <class BE>
class TrainLoopSharded {
  public:
    TrainLoopSharded(BE &be) : be_(be) { memset(&c_, 0, sizeof c_); }
    ~TrainLoopSharded() { free_all(); }

    int run(const uint32_t *d_tokens, const uint64_t *d_off, const uint32_t *d_weight, uint64_t n_tokens_local,
            uint64_t n_chunks_local, uint64_t pos_base, uint64_t n_tokens_global, const TrainConfig &cfg,
            uint32_t *h_merges, int32_t *h_counts, TrainOutcome *out) {
        const uint32_t n_target = cfg.vocab_size - 256;
        memset(out, 0, sizeof *out);
        if (n_target == 0) return 0;
        mode_ = cfg.mode;
        c_.n_pos = (uint32_t)n_tokens_local;
        c_.pos_base = (uint32_t)pos_base;
        const uint64_t np = n_tokens_local ? n_tokens_local : 1, ng = n_tokens_global ? n_tokens_global : 1;
        c_.node = (Node *)be_.alloc(np * sizeof(Node));
        c_.arena_cap = (uint32_t)(3 * np + 16);
        c_.occ = (uint32_t *)be_.alloc((uint64_t)c_.arena_cap * 4);
        c_.hit = (uint32_t *)be_.alloc((np + 16) * 4);
        c_.hit_j = (uint32_t *)be_.alloc((np + 16) * 4);
        c_.hit_y = (uint32_t *)be_.alloc((np + 16) * 4);
        c_.rec_slot = (uint32_t *)be_.alloc((np + 16) * 4);
        c_.rec_pos = (uint32_t *)be_.alloc((np + 16) * 4);
        c_.newp = (uint32_t *)be_.alloc((ng + 65536 + 16) * 4); // births of every rank land here
        // sized by the GLOBAL corpus: the padded all-gather reads max-over-ranks records from every rank's buffer
        c_.xrec_cap = (uint32_t)(2 * ng + 65536 + 64);
        c_.xrec = (XRec *)be_.alloc((uint64_t)c_.xrec_cap * sizeof(XRec));
        c_.ctl = (Ctl *)be_.alloc(sizeof(Ctl));
        c_.merges_out = (uint32_t *)be_.alloc((uint64_t)n_target * 8);
        c_.counts_out = (int32_t *)be_.alloc((uint64_t)n_target * 4);
        // every pair ever: <= 65536 byte pairs + 2 per merged occurrence (< n_tokens_global in total)
        const uint64_t max_pairs = 65536 + 2 * ng + 2 * np + 128;
        const uint32_t cap = next_pow2_u32(max_pairs * MB_LOAD_DEN / MB_LOAD_NUM + 1);
        c_.slot = (Slot *)be_.alloc((uint64_t)cap * sizeof(Slot));
        c_.cap_mask = cap - 1;
        c_.cand_cap = cap;
        c_.cand = (uint32_t *)be_.alloc((uint64_t)c_.cand_cap * 4);
        c_.fix = (uint32_t *)be_.alloc((uint64_t)c_.cand_cap * 4);
        be_.par(PhClearSlots{c_.slot, cap}, cap);

        Ctl h;
        memset(&h, 0, sizeof h);
        h.n_target = n_target;
        h.mode = cfg.mode;
        h.status = ST_NEED_REBUILD;
        h.cmax = CMAX_NONE;
        h.best_tie = ~0ull;
        h.best_cand = NIL;
        h.theta = 1;
        h.big_limit = ~0u;
        h.pf_mode = cfg.pf_mode;
        const uint32_t resident_limit = cfg.engine == 1 ? be_.resident_limit() : 0;
        resident_limit_ = resident_limit;
        h.big_count = resident_limit;
        h.xstep = be_.xstep();
        h.cand_limit = cfg.cand_limit ? cfg.cand_limit : 4096;
        h.min_key_ever = ~0ull;
        h.live_tokens = n_tokens_local;
        be_.upload(c_.ctl, &h, sizeof h);

        be_.par(PhInitNodes{c_, d_tokens, d_off, d_weight, n_chunks_local}, n_tokens_local);
        be_.init_count(c_);
        be_.par(PhExportAll{c_}, cap); // local histogram -> global histogram on every rank
        if (exchange_and_apply()) return -1;
        be_.one(PhAfterInitExchange{c_});
        be_.par(PhInitAlloc{c_}, cap);
        be_.par(PhInitFill{c_}, n_tokens_local);

        for (;;) {
            be_.download(&h, c_.ctl, sizeof h);
            if (h.status == ST_DONE || h.status == ST_EXHAUSTED) break;
            if (h.status == ST_NEED_REBUILD) {
                rebuild(cfg, cap);
            } else if (h.status == ST_FAILED) {
                return -2;
            } else if (h.status == ST_BIG_MERGE) { // the resident CTA handed this merge to the grid (same on every rank)
                be_.one(PhTakeBig{c_});
                if (step_grid(h)) return -1;
                out->n_big++;
            } else if (h.status == ST_RUN && resident_limit) {
                be_.persistent_sharded(c_);
            } else if (h.status == ST_RUN && !h.selected) {
                if (select_grid(h.n_cand)) return -1;
            } else if (h.status == ST_RUN) {
                if (step_grid(h)) return -1;
            } else {
                return -1; // ST_NEED_GROW cannot happen: fixed table
            }
        }
        be_.set_xstep(h.xstep);
        out->n_merges = h.step;
        out->final_status = h.status;
        out->min_key_ever = h.min_key_ever;
        out->n_pairs = h.n_pairs;
        out->table_slots = cap;
        out->n_rebuilds = n_rebuilds_;
        out->n_big = resident_limit ? out->n_big : n_exchanges_;
        out->rescan_bytes = h.rescan_bytes;
        for (int i = 0; i < 8; i++) out->prof[i] = h.prof[i];
        if (h.step) {
            be_.download(h_merges, c_.merges_out, (uint64_t)h.step * 8);
            if (h_counts) be_.download(h_counts, c_.counts_out, (uint64_t)h.step * 4);
        }
        free_all();
        return 0;
    }

  private:
    // one merge driven from the host: the selection is done (h.selected), grid kernels + one all-gather
    int step_grid(const Ctl &h) {
        be_.par(PhHits{c_}, h.seg_len);
        be_.par(PhExportBirths{c_}, 2ull * h.seg_len);
        if (exchange_and_apply()) return -1;
        be_.par(PhMutate{c_}, h.seg_len);
        be_.par(PhSegAlloc{c_}, 4096);
        be_.par(PhSegFill{c_}, 2ull * h.seg_len);
        be_.one(PhFin{c_});
        n_exchanges_++;
        return select_grid(h.n_cand);
    }
    int exchange_and_apply() {
        Ctl h;
        be_.download(&h, c_.ctl, sizeof h); // n_xrec of this rank
        if (h.n_xrec > c_.xrec_cap) return -1;
        const XRec *all = nullptr;
        const uint32_t *counts = nullptr;
        uint32_t stride = 0;
        be_.exchange(c_.xrec, h.n_xrec, &all, &counts, &stride);
        be_.par(PhApplyForeign{c_, all, counts, stride, be_.world(), be_.rank()}, (uint64_t)stride);
        return 0;
    }
    int select_grid(uint32_t n_cand) {
        be_.par(PhSelMax{c_}, n_cand);
        be_.par(PhSelTie{c_}, n_cand);
        be_.one(PhSelCheck{c_});
        if (mode_ == 0) { // FIRST: tied pairs whose first occurrence died need the minimum over all ranks
            Ctl h;
            be_.download(&h, c_.ctl, sizeof h);
            if (h.status == ST_RUN && !h.selected && h.n_fix) { // same replicas -> same decision on every rank
                be_.par(PhSelFixScan{c_}, n_cand);
                be_.par(PhExportFix{c_}, h.n_fix);
                if (exchange_and_apply()) return -1;
                be_.one(PhResetXrec{c_});
                be_.par(PhSelFixTie{c_}, n_cand);
            }
        }
        be_.par(PhSelPick{c_}, n_cand);
        be_.one(PhSelCommit{c_, resident_limit_ ? 1 : 0}); // (1: a merge above Ctl::big_count is marked ST_BIG_MERGE)
        return 0;
    }
    void rebuild(const TrainConfig &cfg, uint64_t cap) {
        be_.one(PhRebuildReset{c_});
        be_.par(PhRebuildHist{c_}, cap);
        be_.one(PhRebuildTheta{c_, cfg.cand_want});
        be_.par(PhRebuildCollect{c_}, cap);
        be_.one(PhSelReset{c_});
        n_rebuilds_++;
    }
    void free_all() {
        void *ps[] = {c_.node, c_.occ, c_.hit, c_.hit_j, c_.hit_y, c_.rec_slot, c_.rec_pos, c_.newp, c_.ctl, c_.merges_out,
                      c_.counts_out, c_.slot, c_.cand, c_.fix, c_.xrec};
        for (void *p : ps)
            if (p) be_.release(p);
        memset(&c_, 0, sizeof c_);
    }

    BE &be_;
    Ctx c_;
    uint64_t n_rebuilds_ = 0, n_exchanges_ = 0;
    uint32_t resident_limit_ = 0;
    int mode_ = 1;
};

// Host epilogue shared by every caller: turn the device outcome into the reference's observable merge list.
//  FIRST:   stop where the table ran empty (Tokenizer.h:586-588).
//  LEXICAL: entries are never erased, so once the best count is 0 the smallest pair ever inserted is
//           returned for every remaining id; no pair ever inserted => loop breaks at once.
inline uint32_t finish_merges(const TrainOutcome &o, uint32_t vocab_size, int mode, uint32_t *merges, int32_t *counts) {
    uint32_t n_target = vocab_size - 256, n = o.n_merges;
    if (mode == 1 && n < n_target && o.min_key_ever != ~0ull) {
        for (; n < n_target; n++) {
            merges[2 * n] = (uint32_t)(o.min_key_ever >> 32);
            merges[2 * n + 1] = (uint32_t)o.min_key_ever;
            if (counts) counts[n] = 0;
        }
    }
    return n;
}

} // namespace mbpe
