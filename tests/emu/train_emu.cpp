// tests/emu/train_emu.cpp -- TEST INFRASTRUCTURE. Sequential host backend for train_driver.hpp.
//
// Runs the SAME phase functions and the SAME orchestration as the CUDA product, but each "parallel" phase is
// a plain loop over thread ids (ascending, descending or shuffled), so the CPU unit tests can check the
// algorithm (delta rules, a==b runs, tie-breaks, candidate/rebuild/grow logic) against the oracle without a
// GPU. It is never linked into libminbpe_b200.so and is not a fallback: the product refuses to run without CUDA.
#include <stdint.h>
#include <stdlib.h>

#include <vector>

#include "../../minbpe-cc_b200/csrc/train_driver.hpp"

using namespace mbpe;

struct HostBE {
    uint32_t nth;
    int order; // 0 ascending, 1 descending, 2 shuffled per phase
    uint64_t n_launch = 0, rng = 0x9E3779B97F4A7C15ull;
    std::vector<uint32_t> perm;

    void *alloc(size_t n) { return calloc(n ? n : 1, 1); }
    void release(void *p) { free(p); }
    void upload(void *d, const void *s, size_t n) { memcpy(d, s, n); }
    void download(void *d, const void *s, size_t n) { memcpy(d, s, n); }
    uint64_t launches() const { return n_launch; }

    const std::vector<uint32_t> &ids() {
        perm.resize(nth);
        for (uint32_t i = 0; i < nth; i++) perm[i] = order == 1 ? nth - 1 - i : i;
        if (order == 2)
            for (uint32_t i = nth - 1; i > 0; i--) {
                rng ^= rng << 13, rng ^= rng >> 7, rng ^= rng << 17;
                std::swap(perm[i], perm[rng % (i + 1)]);
            }
        return perm;
    }
    template <class F>
    void par(const F &f, uint64_t) {
        n_launch++;
        for (uint32_t t : ids()) f(t, nth);
    }
    template <class F>
    void one(const F &f) {
        n_launch++;
        f();
    }
    void init_count(const Ctx &c) { par(PhInitCount{c}, c.n_pos); }

    struct Exec {
        HostBE *be;
        template <class F>
        void par(const F &f) { for (uint32_t t : be->ids()) f(t, be->nth); }
        template <class F, class G>
        void par2(const F &f, const G &g) {
            for (uint32_t t : be->ids()) {
                f(t, be->nth);
                g(t, be->nth);
            }
        }
        template <class F>
        void one(const F &f) { f(); }
        template <class T>
        T load(const T *p) { return *p; }
    };
    void persistent(const Ctx &c) {
        n_launch++;
        Exec ex{this};
        persistent_program<false>(c, ex);
    }
};

extern "C" int emu_train(const uint32_t *tokens, uint64_t n_tokens, const uint64_t *off, uint64_t n_chunks,
                         const uint32_t *weight, uint32_t vocab_size, int mode, int engine, uint32_t nth, int order,
                         uint32_t big_limit, uint32_t cand_want, uint32_t init_slots, uint32_t *merges_out, int32_t *counts_out,
                         uint32_t *n_merges_out, uint64_t *stats /* 8 */) {
    HostBE be{nth, order};
    TrainLoop<HostBE> loop(be);
    TrainConfig cfg{vocab_size, mode, engine, big_limit, cand_want, /*cand_limit*/ cand_want * 4 + 64, init_slots};
    TrainOutcome o;
    int rc = loop.run(tokens, off, weight, n_tokens, n_chunks, cfg, merges_out, counts_out, &o);
    if (rc) return rc;
    *n_merges_out = finish_merges(o, vocab_size, mode, merges_out, counts_out);
    if (stats) {
        stats[0] = o.n_pairs, stats[1] = o.table_slots, stats[2] = o.n_big, stats[3] = o.n_rebuilds;
        stats[4] = o.n_grows, stats[5] = o.rescan_bytes, stats[6] = be.launches(), stats[7] = o.n_merges;
    }
    return 0;
}
