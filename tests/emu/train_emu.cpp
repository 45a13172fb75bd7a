// tests/emu/train_emu.cpp -- TEST INFRASTRUCTURE. Sequential host backend for train_driver.hpp.
//
// Runs the SAME phase functions and the SAME orchestration as the CUDA product, but each "parallel" phase is
// a plain loop over thread ids (ascending, descending or shuffled), so the CPU unit tests can check the
// algorithm (delta rules, a==b runs, tie-breaks, candidate/rebuild/grow logic) against the oracle without a
// GPU. It is never linked into libminbpe_b200.so and is not a fallback: the product refuses to run without CUDA.
#include <stdint.h>
#include <stdlib.h>

#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "../../minbpe-cc_b200/csrc/train_driver.hpp"

using namespace mbpe;

// exchange between emulated ranks (threads of this process): a barrier and memcpy stand in for the NCCL all-gather
struct HostComm {
    uint32_t world;
    std::mutex mu;
    std::condition_variable cv;
    uint32_t arrived = 0, generation = 0;
    std::vector<const mbpe::XRec *> send;
    std::vector<uint32_t> n_send;
    explicit HostComm(uint32_t w) : world(w), send(w, nullptr), n_send(w, 0) {}
    void barrier() {
        std::unique_lock<std::mutex> lk(mu);
        uint32_t gen = generation;
        if (++arrived == world) {
            arrived = 0;
            generation++;
            cv.notify_all();
        } else {
            cv.wait(lk, [&] { return generation != gen; });
        }
    }
};

struct HostBE {
    uint32_t nth;
    int order; // 0 ascending, 1 descending, 2 shuffled per phase
    HostComm *comm = nullptr;
    uint32_t my_rank = 0;
    std::vector<mbpe::XRec> all_recs;
    std::vector<uint32_t> all_counts;
    uint32_t world() const { return comm ? comm->world : 1; }
    uint32_t rank() const { return my_rank; }
    void exchange(const mbpe::XRec *d_send, uint32_t n_send, const mbpe::XRec **d_all, const uint32_t **d_counts,
                  uint32_t *stride) {
        comm->send[my_rank] = d_send;
        comm->n_send[my_rank] = n_send;
        comm->barrier();
        uint32_t mx = 0;
        for (uint32_t r = 0; r < comm->world; r++) mx = std::max(mx, comm->n_send[r]);
        all_counts.assign(comm->n_send.begin(), comm->n_send.end());
        all_recs.assign((size_t)mx * comm->world + 1, mbpe::XRec{0, 0, 0});
        for (uint32_t r = 0; r < comm->world; r++)
            memcpy(all_recs.data() + (size_t)r * mx, comm->send[r], (size_t)comm->n_send[r] * sizeof(mbpe::XRec));
        comm->barrier(); // everybody has copied: send buffers may be reused
        *d_all = all_recs.data();
        *d_counts = all_counts.data();
        *stride = mx;
    }
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
    // sharded resident program: the exchange is the same barrier + memcpy as the host-driven one
    struct ExecSharded : Exec {
        bool select(const Ctx &) { return false; }
        bool exchange_apply(const Ctx &c) {
            const mbpe::XRec *all = nullptr;
            const uint32_t *counts = nullptr;
            uint32_t stride = 0;
            if (c.ctl->n_xrec > c.xrec_cap) {
                c.ctl->status = ST_FAILED;
                return false;
            }
            be->exchange(c.xrec, c.ctl->n_xrec, &all, &counts, &stride);
            par(PhApplyForeign{c, all, counts, stride, be->world(), be->rank()});
            c.ctl->xstep++;
            return true;
        }
    };
    uint32_t resident_count_limit = 0, xstep_ = 0;
    uint32_t resident_limit() const { return resident_count_limit; }
    uint32_t xstep() const { return xstep_; }
    void set_xstep(uint32_t v) { xstep_ = v; }
    void persistent_sharded(const Ctx &c) {
        n_launch++;
        ExecSharded ex{{this}};
        persistent_program_sharded<false>(c, ex);
    }
};

extern "C" int emu_train(const uint32_t *tokens, uint64_t n_tokens, const uint64_t *off, uint64_t n_chunks,
                         const uint32_t *weight, uint32_t vocab_size, int mode, int engine, uint32_t nth, int order,
                         uint32_t big_limit, uint32_t cand_want, uint32_t init_slots, uint32_t *merges_out, int32_t *counts_out,
                         uint32_t *n_merges_out, uint64_t *stats /* 8 */) {
    HostBE be;
    be.nth = nth;
    be.order = order;
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

// `world` emulated ranks, one host thread each; rank r owns a contiguous, token-balanced range of the chunks.
// Every rank must end with the same merge list; rank 0's is returned, rc 7 if any rank disagrees.
// resident_limit: 0 = every step driven from the "host" (engine 0); else merges whose count is at most that run in the
// resident program (engine 1) and the others fall back to the host-driven step.
extern "C" int emu_train_sharded(const uint32_t *tokens, uint64_t n_tokens, const uint64_t *off, uint64_t n_chunks,
                                 const uint32_t *weight, uint32_t vocab_size, int mode, uint32_t world, uint32_t nth,
                                 int order, uint32_t cand_want, uint32_t resident_limit, uint32_t *merges_out,
                                 int32_t *counts_out, uint32_t *n_merges_out) {
    HostComm comm(world);
    std::vector<uint64_t> first(world + 1, n_chunks);
    first[0] = 0;
    for (uint32_t r = 1; r < world; r++) {
        uint64_t target = n_tokens / world * r, c = first[r - 1];
        while (c < n_chunks && off[c] < target) c++;
        first[r] = c;
    }
    const uint32_t n_target = vocab_size > 256 ? vocab_size - 256 : 1;
    std::vector<std::vector<uint32_t>> m(world, std::vector<uint32_t>(2 * (size_t)n_target));
    std::vector<std::vector<int32_t>> cn(world, std::vector<int32_t>(n_target));
    std::vector<uint32_t> nm(world, 0);
    std::vector<int> rcs(world, 0);
    auto work = [&](uint32_t r) {
        HostBE be;
        be.nth = nth;
        be.order = order;
        be.comm = &comm;
        be.my_rank = r;
        be.rng += r * 0x1234567ull;
        be.resident_count_limit = resident_limit;
        const uint64_t c0 = first[r], c1 = first[r + 1], t0 = off[c0], t1 = off[c1];
        std::vector<uint64_t> loff(c1 - c0 + 1);
        for (uint64_t c = c0; c <= c1; c++) loff[c - c0] = off[c] - t0;
        TrainLoopSharded<HostBE> loop(be);
        TrainConfig cfg{vocab_size, mode, resident_limit ? 1 : 0, ~0u, cand_want, cand_want * 4 + 64, 0};
        TrainOutcome o;
        rcs[r] = loop.run(tokens + t0, loff.data(), weight + c0, t1 - t0, c1 - c0, t0, n_tokens, cfg, m[r].data(),
                          cn[r].data(), &o);
        if (rcs[r] == 0) nm[r] = finish_merges(o, vocab_size, mode, m[r].data(), cn[r].data());
    };
    std::vector<std::thread> th;
    for (uint32_t r = 1; r < world; r++) th.emplace_back(work, r);
    work(0);
    for (auto &t : th) t.join();
    for (uint32_t r = 0; r < world; r++) {
        if (rcs[r]) return rcs[r];
        if (nm[r] != nm[0] || memcmp(m[r].data(), m[0].data(), (size_t)nm[0] * 8) ||
            memcmp(cn[r].data(), cn[0].data(), (size_t)nm[0] * 4)) {
            if (getenv("EMU_DEBUG"))
                for (uint32_t i = 0; i < nm[0] && i < nm[r]; i++)
                    if (m[r][2 * i] != m[0][2 * i] || m[r][2 * i + 1] != m[0][2 * i + 1] || cn[r][i] != cn[0][i]) {
                        fprintf(stderr, "rank %u differs from rank 0 at merge %u: (%u,%u)x%d vs (%u,%u)x%d; n=%u vs %u\n", r, i,
                                m[r][2 * i], m[r][2 * i + 1], cn[r][i], m[0][2 * i], m[0][2 * i + 1], cn[0][i], nm[r], nm[0]);
                        break;
                    }
            return 7;
        }
    }
    *n_merges_out = nm[0];
    memcpy(merges_out, m[0].data(), (size_t)nm[0] * 8);
    memcpy(counts_out, cn[0].data(), (size_t)nm[0] * 4);
    return 0;
}
