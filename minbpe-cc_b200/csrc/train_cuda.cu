// train_cuda.cu -- sm_100a backend of the train merge loop + the PairCount seam.
//
// Kernels (names as they appear in ncu):
//   k_init_count            calculate_freqs (Tokenizer.h:127-146): weighted adjacent-pair histogram. 128-bit node
//                           loads, per-CTA shared-memory open-addressed pre-aggregation, then one 64-bit atomic per
//                           distinct pair per CTA into the HBM pair table.
//   k_par<Ph*> / k_one<Ph*> one barrier-separated phase of train_phases.cuh over the whole grid / one thread
//   k_persistent            persistent_program(): one resident 1024-thread CTA runs merge steps back to back
//                           (selection, hits, mutate, segment build) with __syncthreads() as the phase barrier;
//                           control block, step lists and a mirror of the candidate list live in shared memory
//   k_pc_*                  PairCount seam (PairCount.h:27-47): batched upsert, block-then-grid arg-max, lookup
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "train_driver.hpp"

namespace mbpe {

// ---------------------------------------------------------------------------------------------------------
// generic phase launchers
// ---------------------------------------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(256) k_par(const F f) {
    f(blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}
template <class F>
__global__ void __launch_bounds__(32) k_one(const F f) {
    if (threadIdx.x == 0) f();
}

// ---------------------------------------------------------------------------------------------------------
// k_persistent: the resident CTA. The control block and (for steps with <= PS_HIT occurrences) the step lists
// live in shared memory, so the per-step bookkeeping costs shared-memory latency; what remains on the critical
// path are the dependent L2/HBM round trips of the occurrence walk itself.
// ---------------------------------------------------------------------------------------------------------
constexpr int PERSISTENT_THREADS = 1024;
constexpr uint32_t PS_HIT = 2048;        // occurrences per step served from shared-memory lists
constexpr uint32_t PS_REC = 2 * PS_HIT;  // each occurrence creates at most two new-pair records
constexpr int PS_SEL = 4;                // candidates per thread held in registers by the fused selection

struct PersistSmem {
    Ctl ctl;
    uint32_t hit[PS_HIT];
    uint32_t hit_j[PS_HIT];
    uint32_t hit_y[PS_HIT];
    uint32_t rec_slot[PS_REC];
    uint32_t rec_pos[PS_REC];
    uint32_t newp[PS_REC];
    uint32_t cand[PS_SEL * PERSISTENT_THREADS]; // mirror of the candidate list while it fits
    // what the selection needs of every candidate (Ctx::m_*), so that it reads no global memory
    int32_t ccnt[PS_SEL * PERSISTENT_THREADS];
    uint32_t cfirst[PS_SEL * PERSISTENT_THREADS];
    uint32_t clen[PS_SEL * PERSISTENT_THREADS];
    uint32_t cseg[PS_SEL * PERSISTENT_THREADS];
    uint64_t ckey[PS_SEL * PERSISTENT_THREADS];
    // block reductions of the mirror-based selection: one word per warp, re-reduced by every warp (no atomics)
    int32_t red_max[PERSISTENT_THREADS / 32];
    uint32_t red_live[PERSISTENT_THREADS / 32];
    uint64_t red_tie[PERSISTENT_THREADS / 32];
};

__device__ __forceinline__ uint64_t warp_min_u64(uint64_t v) {
    for (int d = 16; d > 0; d >>= 1) {
        uint64_t o = __shfl_xor_sync(0xffffffffu, v, d);
        v = o < v ? o : v;
    }
    return v;
}

// get_top_pair_count for the resident CTA: same result as sel_max .. sel_commit, but each candidate's slot is
// fetched once (one 16-byte load: key, cnt, len) and kept in registers across the three reductions.
template <class SMEM> // PersistSmem / ShardSmem: only red_max, red_live, red_tie are touched (mirror path)
__device__ __forceinline__ void fused_select(const Ctx &c, SMEM *red) {
    Ctl *g = c.ctl; // shared memory
    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint32_t n = g->n_cand;
    const int32_t theta = g->theta, mode = g->mode;
    if (n > PS_SEL * PERSISTENT_THREADS) { // long candidate list (massive ties at low counts): generic phases
        phase_sel_max<true>(c, tid, PERSISTENT_THREADS);
        __syncthreads();
        phase_sel_tie<true>(c, tid, PERSISTENT_THREADS);
        __syncthreads();
        if (tid == 0) phase_sel_check<true>(c);
        __syncthreads();
        if (g->n_fix) {
            phase_sel_fix_scan<true>(c, tid, PERSISTENT_THREADS);
            __syncthreads();
            phase_sel_fix_tie<true>(c, tid, PERSISTENT_THREADS);
            __syncthreads();
        }
        phase_sel_pick<true>(c, tid, PERSISTENT_THREADS);
        __syncthreads();
        if (tid == 0) phase_sel_commit<true>(c, 1);
        __syncthreads();
        return;
    }
    if (c.m_cnt) { // LEXICAL mode with the candidate mirror: everything below comes from shared memory
        // The selection is issue-bound (32 warps run the same code), so warp w owns candidates [128 w, 128 w + 128) and
        // warps without candidates only keep the barriers company; partial results travel through one word per
        // warp and are reduced again by whoever needs them.
        const uint32_t warp = tid >> 5, i0 = warp * (PS_SEL * 32) + lane;
        const bool owner = warp * (PS_SEL * 32) < n; // warp-uniform
        int32_t cv[PS_SEL], best = CMAX_NONE;
        if (owner) {
            uint32_t live = 0;
#pragma unroll
            for (int k = 0; k < PS_SEL; k++) {
                const uint32_t i = i0 + k * 32;
                cv[k] = i < n ? c.m_cnt[i] : CMAX_NONE;
                best = cv[k] > best ? cv[k] : best;
                live += (i < n && cv[k] >= theta);
            }
            best = __reduce_max_sync(0xffffffffu, best);
            live = __reduce_add_sync(0xffffffffu, live);
            if (lane == 0) {
                red->red_max[warp] = best;
                red->red_live[warp] = live;
            }
        } else if (lane == 0) {
            red->red_max[warp] = CMAX_NONE;
            red->red_live[warp] = 0;
        }
        __syncthreads();
        static_assert(PERSISTENT_THREADS / 32 == 32, "one partial result per lane");
        int32_t cmax = CMAX_NONE;
        if (owner || warp == 0) cmax = __reduce_max_sync(0xffffffffu, red->red_max[lane]);
        if (warp == 0) {
            const uint32_t all_live = __reduce_add_sync(0xffffffffu, red->red_live[lane]);
            if (lane == 0) {
                g->cmax = cmax;
                g->n_live = all_live;
                if (cmax == CMAX_NONE || cmax < theta) g->status = ST_NEED_REBUILD;
            }
        }
        // tie-break value of candidate i: LEXICAL the pair itself, FIRST its first live position (PairCount.h:66-74,
        // :195-207). FIRST: a best-count pair whose first occurrence died goes to the fix list; after the rescan of those
        // pairs' segments the mirror is refreshed and the reduction runs once more.
        const bool tied_warp = owner && cmax != CMAX_NONE && cmax >= theta && best == cmax;
        uint64_t mine = ~0ull;
        if (tied_warp) {
#pragma unroll
            for (int k = 0; k < PS_SEL; k++) {
                const uint32_t i = i0 + k * 32;
                if (i < n && cv[k] == cmax) {
                    uint64_t t;
                    if (mode == 1) {
                        t = c.m_key[i];
                    } else {
                        const uint32_t f = c.m_first[i];
                        if (f == NO_FIRST) c.fix[atomicAdd(&g->n_fix, 1u)] = c.cand[i];
                        t = f == NO_FIRST ? ~0ull : (uint64_t)f;
                    }
                    mine = t < mine ? t : mine;
                }
            }
            mine = warp_min_u64(mine);
        }
        if (lane == 0) red->red_tie[warp] = mine;
        __syncthreads();
        if (mode == 0 && g->status == ST_RUN && g->n_fix) { // uniform: n_fix was complete at the barrier
            phase_sel_fix_scan<true>(c, tid, PERSISTENT_THREADS); // Slot::first of the fix list, from their segments
            __syncthreads();
            mine = ~0ull;
            if (tied_warp) {
#pragma unroll
                for (int k = 0; k < PS_SEL; k++) {
                    const uint32_t i = i0 + k * 32;
                    if (i < n && cv[k] == cmax) {
                        uint32_t f = c.m_first[i];
                        if (f == NO_FIRST) {
                            f = __ldcg(&c.slot[c.cand[i]].first);
                            c.m_first[i] = f;
                        }
                        const uint64_t t = f == NO_FIRST ? ~0ull : (uint64_t)f;
                        mine = t < mine ? t : mine;
                    }
                }
                mine = warp_min_u64(mine);
            }
            if (lane == 0) red->red_tie[warp] = mine;
            __syncthreads();
        }
        if (g->status != ST_RUN) return; // (written by thread 0 before the barrier; uniform)
        if (owner && best == cmax) {
            const uint64_t win = warp_min_u64(red->red_tie[lane]);
#pragma unroll
            for (int k = 0; k < PS_SEL; k++) {
                const uint32_t i = i0 + k * 32;
                if (i < n && cv[k] == cmax && (mode == 1 ? c.m_key[i] : (uint64_t)c.m_first[i]) == win) { // exactly one
                    const uint64_t key = c.m_key[i]; // thread: keys are unique, and so are first positions
                    const uint32_t seg_len = c.m_len[i], step = g->step;
                    g->best_tie = win;
                    g->best_slot = c.cand[i];
                    g->best_cand = i;
                    g->seg_len = seg_len;
                    const uint64_t need = (uint64_t)g->n_pairs + 2ull * seg_len + 64;
                    if (need * MB_LOAD_DEN > ((uint64_t)c.cap_mask + 1) * MB_LOAD_NUM) {
                        g->status = ST_NEED_GROW;
                    } else {
                        g->a = (uint32_t)(key >> 32);
                        g->b = (uint32_t)key;
                        g->new_id = 256 + step;
                        g->seg = c.m_seg[i];
                        c.merges_out[2 * step] = (uint32_t)(key >> 32);
                        c.merges_out[2 * step + 1] = (uint32_t)key;
                        c.counts_out[step] = cmax;
                        g->selected = 1;
                        if (seg_len > g->big_limit || (g->big_count && (uint32_t)cmax > g->big_count)) g->status = ST_BIG_MERGE;
                    }
                }
            }
        }
        __syncthreads();
        return;
    }
    uint32_t slot_id[PS_SEL];
    uint4 head[PS_SEL]; // {key.lo, key.hi, cnt, len}
    uint4 tail[PS_SEL]; // {first, seg, fill, pad}
    int32_t best = CMAX_NONE;
    uint32_t live = 0;
#pragma unroll
    for (int k = 0; k < PS_SEL; k++) {
        uint32_t i = tid + k * PERSISTENT_THREADS;
        slot_id[k] = i < n ? ctl_ld<true>(&c.cand[i]) : NIL; // shared-memory mirror (or this CTA's own global list)
    }
#pragma unroll
    for (int k = 0; k < PS_SEL; k++) {
        if (slot_id[k] != NIL) {
            const uint4 *sp = reinterpret_cast<const uint4 *>(&c.slot[slot_id[k]]);
            head[k] = __ldcg(sp);
            tail[k] = __ldcg(sp + 1);
            int32_t v = (int32_t)head[k].z;
            best = v > best ? v : best;
            live += (v >= theta);
        }
    }
    best = __reduce_max_sync(0xffffffffu, best);
    live = __reduce_add_sync(0xffffffffu, live);
    if (lane == 0) {
        if (best != CMAX_NONE) atomicMax(&g->cmax, best);
        if (live) atomicAdd(&g->n_live, live);
    }
    __syncthreads();
    const int32_t cmax = g->cmax;
    if (cmax == CMAX_NONE || cmax < theta) {
        __syncthreads();
        if (tid == 0) g->status = ST_NEED_REBUILD;
        __syncthreads();
        return;
    }
    uint64_t tie[PS_SEL];
    uint64_t mine = ~0ull;
#pragma unroll
    for (int k = 0; k < PS_SEL; k++) {
        tie[k] = ~0ull;
        if (slot_id[k] != NIL && (int32_t)head[k].z == cmax) {
            if (mode == 1) {
                tie[k] = ((uint64_t)head[k].y << 32) | head[k].x;
            } else {
                uint32_t f = tail[k].x;
                if (f == NO_FIRST)
                    c.fix[atomicAdd(&g->n_fix, 1u)] = slot_id[k];
                else
                    tie[k] = f;
            }
            mine = tie[k] < mine ? tie[k] : mine;
        }
    }
    mine = warp_min_u64(mine);
    if (lane == 0 && mine != ~0ull) atomicMin((unsigned long long *)&g->best_tie, (unsigned long long)mine);
    __syncthreads();
    if (g->n_fix) { // FIRST mode, a tied pair lost its first occurrence: recompute from its segment
        phase_sel_fix_scan<true>(c, tid, PERSISTENT_THREADS);
        __syncthreads();
        phase_sel_fix_tie<true>(c, tid, PERSISTENT_THREADS);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < PS_SEL; k++)
            if (slot_id[k] != NIL && (int32_t)head[k].z == cmax && tie[k] == ~0ull)
                tie[k] = __ldcg(&c.slot[slot_id[k]].first);
    }
    const uint64_t win = g->best_tie;
#pragma unroll
    for (int k = 0; k < PS_SEL; k++)
        if (slot_id[k] != NIL && (int32_t)head[k].z == cmax && tie[k] == win) { // exactly one thread
            // phase_sel_commit, with the slot already in registers (no further loads on the critical path)
            const uint32_t seg_len = head[k].w, step = g->step;
            g->best_slot = slot_id[k];
            g->seg_len = seg_len;
            const uint64_t need = (uint64_t)g->n_pairs + 2ull * seg_len + 64;
            if (need * MB_LOAD_DEN > ((uint64_t)c.cap_mask + 1) * MB_LOAD_NUM) {
                g->status = ST_NEED_GROW;
            } else {
                g->a = head[k].y;
                g->b = head[k].x;
                g->new_id = 256 + step;
                g->seg = tail[k].y;
                c.merges_out[2 * step] = head[k].y;
                c.merges_out[2 * step + 1] = head[k].x;
                c.counts_out[step] = cmax;
                g->selected = 1;
                if (seg_len > g->big_limit || (g->big_count && (uint32_t)cmax > g->big_count)) g->status = ST_BIG_MERGE;
            }
        }
    __syncthreads();
}

__global__ void __launch_bounds__(PERSISTENT_THREADS, 1) k_persistent(const Ctx cg) {
    extern __shared__ __align__(16) unsigned char ps_raw[];
    PersistSmem *sm = reinterpret_cast<PersistSmem *>(ps_raw);
    const uint32_t tid = threadIdx.x;
    constexpr uint32_t CTL_WORDS = sizeof(Ctl) / 4;
    if (tid < CTL_WORDS) reinterpret_cast<uint32_t *>(&sm->ctl)[tid] = __ldcg(reinterpret_cast<const uint32_t *>(cg.ctl) + tid);
    __syncthreads();
    Ctx c = cg;
    c.ctl = &sm->ctl;
    Ctl *g = &sm->ctl;
    const uint32_t n_cand_in = g->n_cand;
    const bool cand_in_smem = n_cand_in <= PS_SEL * PERSISTENT_THREADS;
    if (cand_in_smem) { // the candidate list lives in shared memory while this launch runs
        for (uint32_t i = tid; i < n_cand_in; i += PERSISTENT_THREADS) sm->cand[i] = __ldcg(&cg.cand[i]);
        c.cand = sm->cand;
        c.cand_cap = PS_SEL * PERSISTENT_THREADS; // appends past it are dropped and phase_fin asks for a rebuild
        __syncthreads();
        if (!cg.xrec && cg.m_cap) { // single GPU (m_cap != 0: the host allows it): mirror the candidates' slots
            for (uint32_t i = tid; i < n_cand_in; i += PERSISTENT_THREADS) {
                const uint32_t s = sm->cand[i];
                const uint4 *sp = reinterpret_cast<const uint4 *>(&cg.slot[s]);
                const uint4 head = __ldcg(sp), tail = __ldcg(sp + 1); // {key.lo, key.hi, cnt, len} {first, seg, fill, pad}
                sm->ccnt[i] = (int32_t)head.z;
                sm->cfirst[i] = tail.x;
                sm->ckey[i] = ((uint64_t)head.y << 32) | head.x;
                sm->clen[i] = head.w;
                sm->cseg[i] = tail.y;
                if (tail.w != i + 1) cg.slot[s].pad = i + 1; // (kept by seg_alloc / rebuild_collect; cheap insurance)
            }
            c.m_cnt = sm->ccnt;
            c.m_first = sm->cfirst;
            c.m_key = sm->ckey;
            c.m_len = sm->clen;
            c.m_seg = sm->cseg;
            c.m_cap = PS_SEL * PERSISTENT_THREADS;
            __syncthreads();
        }
    }
    long long t0 = clock64();
    const long long t_enter = t0;
    auto lap = [&](int i) { // thread 0 only: cycles since the previous lap go to counter i
        if (tid == 0) {
            long long t1 = clock64();
            g->prof[i] += (uint64_t)(t1 - t0);
            t0 = t1;
        }
    };
    for (;;) {
        if (g->status != ST_RUN) break;
        if (g->selected == 0) {
            if (tid == 0) { // debug counters: candidate-list length seen by the selection
                g->prof[7] += g->n_cand;
                if (g->n_cand > PS_SEL * PERSISTENT_THREADS) g->dbg[0] += 1;
                if (g->n_cand > g->dbg[1]) g->dbg[1] = g->n_cand;
            }
            long long ts = clock64();
            fused_select(c, sm);
            if (tid == 0) {
                uint64_t d = (uint64_t)(clock64() - ts);
                if (d > g->dbg[2]) g->dbg[2] = d;
                if (d > 20000) g->dbg[3] += 1;
            }
            lap(0);
            if (g->status != ST_RUN) break;
        }
        Ctx w = c; // small steps keep their lists in shared memory
        if (g->seg_len <= PS_HIT) {
            w.hit = sm->hit;
            w.hit_j = sm->hit_j;
            w.hit_y = sm->hit_y;
            w.rec_slot = sm->rec_slot;
            w.rec_pos = sm->rec_pos;
            w.newp = sm->newp;
        }
        const long long th0 = clock64();
        phase_hits<true>(w, tid, PERSISTENT_THREADS);
        __syncthreads();
        if (tid == 0) {
            const uint32_t L = g->seg_len;
            const int cls = L <= 32 ? 0 : L <= 256 ? 1 : L <= 1024 ? 2 : L <= 2048 ? 3 : L <= 8192 ? 4 : 5;
            g->hh_steps[cls] += 1;
            g->hh_cycles[cls] += (uint64_t)(clock64() - th0);
            g->hh_occ[cls] += L;
            for (int i = 0; i < 6; i++) {
                g->pt_sum[cls][i] += g->pt_max[i];
                g->pt_max[i] = 0;
            }
        }
        lap(1);
        phase_mutate<true>(w, tid, PERSISTENT_THREADS);   // corpus nodes
        phase_seg_alloc<true>(w, tid, PERSISTENT_THREADS); // new slots: disjoint data, same barrier
        __syncthreads();
        lap(2);
        phase_seg_fill<true>(w, tid, PERSISTENT_THREADS);
        __syncthreads();
        lap(3);
        if (tid == 0) {
            phase_fin<true>(w);
            g->prof[5] += 1;
        }
        __syncthreads();
        lap(4);
    }
    if (tid == 0) g->prof[6] += (uint64_t)(clock64() - t_enter);
    __syncthreads();
    if (cand_in_smem) {
        const uint32_t n_out = min(g->n_cand, (uint32_t)(PS_SEL * PERSISTENT_THREADS));
        for (uint32_t i = tid; i < n_out; i += PERSISTENT_THREADS) cg.cand[i] = sm->cand[i];
    }
    if (tid < CTL_WORDS) reinterpret_cast<uint32_t *>(cg.ctl)[tid] = reinterpret_cast<const uint32_t *>(&sm->ctl)[tid];
}

// ---------------------------------------------------------------------------------------------------------
// k_persistent_sharded: the resident CTA of SHARDED training (one per rank = per GPU). Same step structure as
// k_persistent -- the device transcription of persistent_program_sharded() in train_phases.cuh, which the CPU test tier
// runs on emulated ranks -- plus the exchange of the step's count deltas with the other ranks done by the CTA itself:
// every rank has mapped the other ranks' inboxes (cudaIpc), so the records leave as plain 16-byte stores over NVLink,
// followed by a system-scope fence and one flag word per peer {exchange number, record count}; then the CTA polls the
// flag words the peers write into ITS memory and applies their records. No kernel launch, no host round trip and no
// NCCL call per merge. Two buffers per (sender, receiver) by exchange parity: a rank can be at most one exchange ahead
// of a peer (it needs the peer's records of exchange k to finish k), so a buffer is never overwritten before it was read.
// ---------------------------------------------------------------------------------------------------------
constexpr uint32_t XCH_MAX_WORLD = 8;
struct PeerLinks {
    XRec *send_to[XCH_MAX_WORLD];               // peer p's inbox block for THIS rank: [parity][cap] (mapped peer memory; self: null)
    unsigned long long *flag_to[XCH_MAX_WORLD]; // peer p's flag words for this rank: [parity]
    const XRec *inbox;                          // this rank's inbox: [source rank][parity][cap]
    unsigned long long *flags;                  // this rank's flag words: [source rank][parity]
    uint32_t cap, world, rank;
};

struct ShardSmem {
    Ctl ctl;
    uint32_t hit[PS_HIT];
    uint32_t hit_j[PS_HIT];
    uint32_t hit_y[PS_HIT];
    uint32_t rec_slot[PS_REC];
    uint32_t rec_pos[PS_REC];
    uint32_t cand[PS_SEL * PERSISTENT_THREADS];
    uint32_t counts[XCH_MAX_WORLD];
    // LEXICAL mode: the candidate mirror of k_persistent (see PersistSmem); the other ranks' records update it in
    // phase_apply_foreign through Slot::pad, like this rank's own decrements do
    int32_t ccnt[PS_SEL * PERSISTENT_THREADS];
    uint32_t cfirst[PS_SEL * PERSISTENT_THREADS];
    uint32_t clen[PS_SEL * PERSISTENT_THREADS];
    uint32_t cseg[PS_SEL * PERSISTENT_THREADS];
    uint64_t ckey[PS_SEL * PERSISTENT_THREADS];
    int32_t red_max[PERSISTENT_THREADS / 32];
    uint32_t red_live[PERSISTENT_THREADS / 32];
    uint64_t red_tie[PERSISTENT_THREADS / 32];
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// send c.xrec[0 .. n_xrec) to every peer, receive theirs, apply them. All threads of the CTA. false: ST_FAILED.
__device__ __forceinline__ bool xch_apply(const Ctx &c, ShardSmem *sm, const PeerLinks &pl) {
    Ctl *g = c.ctl; // shared memory
    const uint32_t tid = threadIdx.x;
    const uint32_t n = g->n_xrec, xs = g->xstep, par = xs & 1u, tag = xs + 1u;
    if (n > pl.cap || n > c.xrec_cap) { // cannot happen while Ctl::big_count <= cap / 4; never send a truncated list
        __syncthreads();
        if (tid == 0) g->status = ST_FAILED;
        __syncthreads();
        return false;
    }
    const uint4 *src = reinterpret_cast<const uint4 *>(c.xrec);
    for (uint32_t p = 0; p < pl.world; p++) {
        if (p == pl.rank) continue;
        uint4 *dst = reinterpret_cast<uint4 *>(pl.send_to[p] + (size_t)par * pl.cap);
        for (uint32_t i = tid; i < n; i += PERSISTENT_THREADS) dst[i] = __ldcg(&src[i]);
    }
    __threadfence_system(); // every thread's records are visible system-wide before the flags go out
    __syncthreads();
    if (tid < pl.world) {
        uint32_t cnt = 0;
        if (tid != pl.rank) {
            st_release_sys(pl.flag_to[tid] + par, ((unsigned long long)tag << 32) | n);
            const unsigned long long *f = pl.flags + tid * 2 + par;
            const long long t0 = clock64();
            for (;;) {
                const unsigned long long v = ld_acquire_sys(f);
                if ((uint32_t)(v >> 32) == tag) {
                    cnt = (uint32_t)v;
                    break;
                }
                if (clock64() - t0 > 8000000000ll) { // ~4 s: the peer is gone (or the ranks' step sequences diverged)
                    g->status = ST_FAILED;
                    break;
                }
            }
        }
        sm->counts[tid] = cnt;
    }
    __syncthreads();
    if (g->status == ST_FAILED) return false;
    phase_apply_foreign<true>(c, pl.inbox + (size_t)par * pl.cap, sm->counts, 2 * pl.cap, pl.world, pl.rank, tid, PERSISTENT_THREADS);
    if (tid == 0) g->xstep = xs + 1u;
    __syncthreads();
    return true;
}

__global__ void __launch_bounds__(PERSISTENT_THREADS, 1) k_persistent_sharded(const Ctx cg, const PeerLinks pl) {
    extern __shared__ __align__(16) unsigned char ps_raw[];
    ShardSmem *sm = reinterpret_cast<ShardSmem *>(ps_raw);
    const uint32_t tid = threadIdx.x;
    constexpr uint32_t CTL_WORDS = sizeof(Ctl) / 4;
    if (tid < CTL_WORDS) reinterpret_cast<uint32_t *>(&sm->ctl)[tid] = __ldcg(reinterpret_cast<const uint32_t *>(cg.ctl) + tid);
    __syncthreads();
    Ctx c = cg;
    c.ctl = &sm->ctl;
    c.m_cnt = nullptr;
    c.m_cap = 0;
    Ctl *g = &sm->ctl;
    const uint32_t n_cand_in = g->n_cand;
    const bool cand_in_smem = n_cand_in <= PS_SEL * PERSISTENT_THREADS;
    if (cand_in_smem) {
        for (uint32_t i = tid; i < n_cand_in; i += PERSISTENT_THREADS) sm->cand[i] = __ldcg(&cg.cand[i]);
        c.cand = sm->cand;
        c.cand_cap = PS_SEL * PERSISTENT_THREADS; // appends past it are dropped and phase_fin asks for a rebuild
        __syncthreads();
        if (g->mode == 1 && cg.m_cap) { // LEXICAL (m_cap != 0: the host allows it): mirror the candidates' slots, as k_persistent does
            for (uint32_t i = tid; i < n_cand_in; i += PERSISTENT_THREADS) {
                const uint32_t s = sm->cand[i];
                const uint4 *sp = reinterpret_cast<const uint4 *>(&cg.slot[s]);
                const uint4 head = __ldcg(sp), tail = __ldcg(sp + 1); // {key.lo, key.hi, cnt, len} {first, seg, fill, pad}
                sm->ccnt[i] = (int32_t)head.z;
                sm->cfirst[i] = tail.x;
                sm->ckey[i] = ((uint64_t)head.y << 32) | head.x;
                sm->clen[i] = head.w;
                sm->cseg[i] = tail.y;
                if (tail.w != i + 1) cg.slot[s].pad = i + 1;
            }
            c.m_cnt = sm->ccnt;
            c.m_first = sm->cfirst;
            c.m_key = sm->ckey;
            c.m_len = sm->clen;
            c.m_seg = sm->cseg;
            c.m_cap = PS_SEL * PERSISTENT_THREADS;
            __syncthreads();
        }
    }
    const long long t_enter = clock64();
    for (;;) {
        if (g->status != ST_RUN) break;
        if (g->selected == 0) {
            if (g->mode == 1) {
                fused_select(c, sm); // LEXICAL: a function of the replicated table only
            } else { // FIRST: a tied pair that lost its first occurrence needs the minimum over ALL ranks' occurrences
                phase_sel_max<true>(c, tid, PERSISTENT_THREADS);
                __syncthreads();
                phase_sel_tie<true>(c, tid, PERSISTENT_THREADS);
                __syncthreads();
                if (tid == 0) phase_sel_check<true>(c);
                __syncthreads();
                if (g->status == ST_RUN && g->n_fix) { // same replicas: same decision on every rank
                    phase_sel_fix_scan<true>(c, tid, PERSISTENT_THREADS);
                    __syncthreads();
                    phase_export_fix<true>(c, tid, PERSISTENT_THREADS);
                    __syncthreads();
                    if (!xch_apply(c, sm, pl)) break;
                    if (tid == 0) g->n_xrec = 0;
                    __syncthreads();
                    phase_sel_fix_tie<true>(c, tid, PERSISTENT_THREADS);
                    __syncthreads();
                }
                phase_sel_pick<true>(c, tid, PERSISTENT_THREADS);
                __syncthreads();
                if (tid == 0) phase_sel_commit<true>(c, 1);
                __syncthreads();
            }
            if (g->status != ST_RUN) break;
        }
        Ctx w = c; // small steps keep their lists in shared memory (newp stays global: the other ranks' births land there too)
        if (g->seg_len <= PS_HIT) {
            w.hit = sm->hit;
            w.hit_j = sm->hit_j;
            w.hit_y = sm->hit_y;
            w.rec_slot = sm->rec_slot;
            w.rec_pos = sm->rec_pos;
        }
        phase_hits<true>(w, tid, PERSISTENT_THREADS);
        __syncthreads();
        phase_export_births<true>(w, tid, PERSISTENT_THREADS);
        __syncthreads();
        if (!xch_apply(w, sm, pl)) break;
        phase_mutate<true>(w, tid, PERSISTENT_THREADS);
        phase_seg_alloc<true>(w, tid, PERSISTENT_THREADS);
        __syncthreads();
        phase_seg_fill<true>(w, tid, PERSISTENT_THREADS);
        __syncthreads();
        if (tid == 0) {
            phase_fin<true>(w);
            g->prof[5] += 1;
        }
        __syncthreads();
    }
    __syncthreads();
    if (tid == 0) g->prof[6] += (uint64_t)(clock64() - t_enter);
    __syncthreads();
    if (cand_in_smem) {
        const uint32_t n_out = min(g->n_cand, (uint32_t)(PS_SEL * PERSISTENT_THREADS));
        for (uint32_t i = tid; i < n_out; i += PERSISTENT_THREADS) cg.cand[i] = sm->cand[i];
    }
    if (tid < CTL_WORDS) reinterpret_cast<uint32_t *>(cg.ctl)[tid] = reinterpret_cast<const uint32_t *>(&sm->ctl)[tid];
}

// ---------------------------------------------------------------------------------------------------------
// k_init_count: K1, the weighted adjacent-pair histogram
// ---------------------------------------------------------------------------------------------------------
constexpr int IC_THREADS = 256;
constexpr int IC_SLOTS = 4096; // shared-memory table: 4096 * (8 + 4 + 4 + 4) B = 80 KB
constexpr int IC_ITEMS = 16;   // positions per thread per tile
struct IcSmem {
    unsigned long long key[IC_SLOTS];
    uint32_t w[IC_SLOTS];
    uint32_t n[IC_SLOTS];
    uint32_t first[IC_SLOTS];
    uint32_t used;
};

__device__ __forceinline__ void ic_flush(const Ctx &c, IcSmem *sm) {
    __syncthreads();
    for (int s = threadIdx.x; s < IC_SLOTS; s += blockDim.x) {
        unsigned long long k = sm->key[s];
        if (k != EMPTY_KEY) {
            count_one(c, k, sm->w[s], sm->n[s], sm->first[s]);
            sm->key[s] = EMPTY_KEY;
            sm->w[s] = 0;
            sm->n[s] = 0;
            sm->first[s] = NO_FIRST;
        }
    }
    if (threadIdx.x == 0) sm->used = 0;
    __syncthreads();
}

__global__ void __launch_bounds__(IC_THREADS) k_init_count(const Ctx c) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    IcSmem *sm = reinterpret_cast<IcSmem *>(smem_raw);
    for (int s = threadIdx.x; s < IC_SLOTS; s += blockDim.x) {
        sm->key[s] = EMPTY_KEY;
        sm->w[s] = 0;
        sm->n[s] = 0;
        sm->first[s] = NO_FIRST;
    }
    if (threadIdx.x == 0) sm->used = 0;
    __syncthreads();

    const uint32_t tile = IC_THREADS * IC_ITEMS;
    const uint4 *nodes = reinterpret_cast<const uint4 *>(c.node);
    for (uint64_t base = (uint64_t)blockIdx.x * tile; base < c.n_pos; base += (uint64_t)gridDim.x * tile) {
#pragma unroll 4
        for (int it = 0; it < IC_ITEMS; it++) {
            uint64_t i = base + (uint64_t)it * IC_THREADS + threadIdx.x; // coalesced 16-byte loads
            if (i >= c.n_pos) break;
            uint4 nd = __ldg(&nodes[i]); // {tok, nxt, prv, wt}
            if (nd.y == NIL) continue;
            uint32_t tok2 = __ldg(&nodes[nd.y]).x; // the neighbour: same or next 128-byte line
            unsigned long long key = pair_key(nd.x, tok2);
            uint32_t s = hash_key(key) & (IC_SLOTS - 1);
            bool done = false;
            for (int probe = 0; probe < 16; probe++) {
                unsigned long long k = sm->key[s];
                if (k == EMPTY_KEY) {
                    k = atomicCAS(&sm->key[s], EMPTY_KEY, key);
                    if (k == EMPTY_KEY) {
                        atomicAdd(&sm->used, 1u);
                        k = key;
                    }
                }
                if (k == key) {
                    atomicAdd(&sm->w[s], nd.w);
                    atomicAdd(&sm->n[s], 1u);
                    atomicMin(&sm->first[s], c.pos_base + (uint32_t)i);
                    done = true;
                    break;
                }
                s = (s + 1) & (IC_SLOTS - 1);
            }
            if (!done) count_one(c, key, nd.w, 1u, c.pos_base + (uint32_t)i); // crowded neighbourhood: straight to HBM
        }
        __syncthreads();
        if (sm->used > IC_SLOTS / 2) ic_flush(c, sm); // block-uniform: read after the barrier
    }
    ic_flush(c, sm);
}

// ---------------------------------------------------------------------------------------------------------
// CUDA backend for TrainLoop
// ---------------------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------------------
// NCCL, resolved at run time (libnccl.so.2: the copy already loaded by the host process -- e.g. torch's -- or the
// system one). Only the sharded trainer needs it; single-GPU use never touches it.
// ---------------------------------------------------------------------------------------------------------
struct NcclApi {
    typedef struct ncclComm *comm_t;
    struct unique_id {
        char internal[128];
    };
    int (*GetUniqueId)(unique_id *) = nullptr;
    int (*CommInitRank)(comm_t *, int, unique_id, int) = nullptr;
    int (*CommDestroy)(comm_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int /*dtype*/, comm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
    static NcclApi &get() {
        static NcclApi api;
        static bool tried = false;
        if (!tried) {
            tried = true;
            void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
            if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
            if (h) {
                api.GetUniqueId = (int (*)(unique_id *))dlsym(h, "ncclGetUniqueId");
                api.CommInitRank = (int (*)(comm_t *, int, unique_id, int))dlsym(h, "ncclCommInitRank");
                api.CommDestroy = (int (*)(comm_t))dlsym(h, "ncclCommDestroy");
                api.AllGather = (int (*)(const void *, void *, size_t, int, comm_t, cudaStream_t))dlsym(h, "ncclAllGather");
                api.GetErrorString = (const char *(*)(int))dlsym(h, "ncclGetErrorString");
                api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather;
            }
        }
        return api;
    }
};
constexpr int NCCL_UINT8 = 1; // ncclUint8 (nccl.h ncclDataType_t)

struct CudaBE {
    cudaStream_t stream = nullptr;
    // sharded training
    NcclApi::comm_t nccl = nullptr;
    uint32_t comm_world = 1, comm_rank = 0;
    XRec *d_all = nullptr;
    uint64_t all_cap = 0;
    uint32_t *d_counts = nullptr, *d_my_count = nullptr;
    uint32_t world() const { return comm_world; }
    uint32_t rank() const { return comm_rank; }
    // resident sharded program (k_persistent_sharded): peer inboxes mapped by mbpe_comm_create, null = not available
    const PeerLinks *links = nullptr;
    uint32_t *xstep_ptr = nullptr; // exchanges done so far on this communicator (continues from run to run)
    uint32_t resident_limit() const { return links ? links->cap / 4 : 0; } // <= 4 records per live local occurrence <= count
    uint32_t xstep() const { return xstep_ptr ? *xstep_ptr : 0; }
    void set_xstep(uint32_t v) {
        if (xstep_ptr) *xstep_ptr = v;
    }
    void persistent_sharded(const Ctx &c) {
        if (err != cudaSuccess) return;
        note(cudaFuncSetAttribute(k_persistent_sharded, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ShardSmem)),
             "smem attr");
        Ctx cl = c;
        static const bool no_mirror = getenv("MBPE_NO_CAND_MIRROR") != nullptr || getenv("MBPE_NO_SHARD_MIRROR") != nullptr;
        cl.m_cap = no_mirror ? 0 : 1; // request the shared-memory candidate mirror (lexical mode; the kernel sets the real capacity)
        k_persistent_sharded<<<1, PERSISTENT_THREADS, sizeof(ShardSmem), stream>>>(cl, *links);
        n_launch++;
        note(cudaGetLastError(), "k_persistent_sharded launch");
    }
    // all-gather of this step's count deltas over NVLink: counts first (they size the padded second gather)
    void exchange(const XRec *d_send, uint32_t n_send, const XRec **out_all, const uint32_t **out_counts, uint32_t *stride) {
        *out_all = d_all;
        *out_counts = d_counts;
        *stride = 0;
        if (err != cudaSuccess) return;
        NcclApi &api = NcclApi::get();
        if (!d_counts) {
            note(cudaMalloc(&d_counts, comm_world * 4), "cudaMalloc counts");
            note(cudaMalloc(&d_my_count, 4), "cudaMalloc count");
        }
        note(cudaMemcpyAsync(d_my_count, &n_send, 4, cudaMemcpyHostToDevice, stream), "H2D count");
        if (api.AllGather(d_my_count, d_counts, 4, NCCL_UINT8, nccl, stream) != 0) note(cudaErrorUnknown, "ncclAllGather counts");
        std::vector<uint32_t> h(comm_world);
        note(cudaMemcpyAsync(h.data(), d_counts, comm_world * 4, cudaMemcpyDeviceToHost, stream), "D2H counts");
        note(cudaStreamSynchronize(stream), "sync");
        uint32_t mx = 0;
        for (uint32_t v : h) mx = std::max(mx, v);
        if (mx == 0) {
            *out_counts = d_counts;
            return;
        }
        if ((uint64_t)mx * comm_world > all_cap) {
            if (d_all) note(cudaFree(d_all), "cudaFree");
            all_cap = (uint64_t)mx * comm_world * 2;
            note(cudaMalloc(&d_all, all_cap * sizeof(XRec)), "cudaMalloc exchange buffer");
        }
        // every rank sends `mx` records (its own count + padding): all send buffers are sized for the global worst case
        if (api.AllGather(d_send, d_all, (size_t)mx * sizeof(XRec), NCCL_UINT8, nccl, stream) != 0)
            note(cudaErrorUnknown, "ncclAllGather records");
        n_launch += 2;
        *out_all = d_all;
        *out_counts = d_counts;
        *stride = mx;
    }

    int sms = 148;
    uint64_t n_launch = 0;
    cudaError_t err = cudaSuccess;
    const char *err_what = "";
    Ctl *pinned = nullptr; // staging for the per-iteration control read
    size_t l2_window_max = 0, l2_persist_bytes = 0;

    void note(cudaError_t e, const char *what) {
        if (e != cudaSuccess && err == cudaSuccess) {
            err = e;
            err_what = what;
        }
    }
    void *alloc(size_t n) {
        void *p = nullptr;
        if (err == cudaSuccess) note(cudaMallocAsync(&p, n ? n : 1, stream), "cudaMallocAsync");
        return p;
    }
    void release(void *p) {
        if (p) note(cudaFreeAsync(p, stream), "cudaFreeAsync");
    }
    void upload(void *d, const void *s, size_t n) {
        if (err != cudaSuccess) return;
        note(cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, stream), "H2D");
        note(cudaStreamSynchronize(stream), "sync");
    }
    void download(void *d, const void *s, size_t n) {
        if (err != cudaSuccess) { // make the driver loop terminate
            if (n == sizeof(Ctl)) reinterpret_cast<Ctl *>(d)->status = ST_DONE;
            return;
        }
        if (n == sizeof(Ctl) && pinned) {
            note(cudaMemcpyAsync(pinned, s, n, cudaMemcpyDeviceToHost, stream), "D2H ctl");
            note(cudaStreamSynchronize(stream), "sync");
            memcpy(d, pinned, n);
        } else {
            note(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, stream), "D2H");
            note(cudaStreamSynchronize(stream), "sync");
        }
        if (err != cudaSuccess && n == sizeof(Ctl)) reinterpret_cast<Ctl *>(d)->status = ST_DONE;
    }
    unsigned grid_for(uint64_t n_items) const {
        uint64_t blocks = (n_items + 255) / 256;
        uint64_t cap = (uint64_t)sms * 8; // 8 resident 256-thread CTAs per SM
        return (unsigned)std::max<uint64_t>(1, std::min(blocks, cap));
    }
    template <class F>
    void par(const F &f, uint64_t n_items) {
        if (err != cudaSuccess) return;
        k_par<F><<<grid_for(n_items), 256, 0, stream>>>(f);
        n_launch++;
        note(cudaGetLastError(), "k_par launch");
    }
    template <class F>
    void one(const F &f) {
        if (err != cudaSuccess) return;
        k_one<F><<<1, 32, 0, stream>>>(f);
        n_launch++;
        note(cudaGetLastError(), "k_one launch");
    }
    void init_count(const Ctx &c) {
        if (err != cudaSuccess) return;
        // the attribute belongs to the current device's context: set it on every launch (it is cheap), so a process
        // that trains on device 0 and then on device 1 works
        note(cudaFuncSetAttribute(k_init_count, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(IcSmem)), "smem attr");
        uint64_t tiles = ((uint64_t)c.n_pos + IC_THREADS * IC_ITEMS - 1) / (IC_THREADS * IC_ITEMS);
        unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(tiles, (uint64_t)sms * 2));
        k_init_count<<<grid, IC_THREADS, sizeof(IcSmem), stream>>>(c);
        n_launch++;
        note(cudaGetLastError(), "k_init_count launch");
    }
    void persistent(const Ctx &c) {
        if (err != cudaSuccess) return;
        note(cudaFuncSetAttribute(k_persistent, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PersistSmem)),
             "smem attr"); // per device, see init_count
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(1);
        cfg.blockDim = dim3(PERSISTENT_THREADS);
        cfg.dynamicSmemBytes = sizeof(PersistSmem);
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        cfg.attrs = attr;
        cfg.numAttrs = 0;
        if (l2_window_max > 0) { // keep the pair table (the randomly probed structure) resident in L2
            size_t bytes = ((size_t)c.cap_mask + 1) * sizeof(Slot);
            attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
            attr[0].val.accessPolicyWindow.base_ptr = c.slot;
            attr[0].val.accessPolicyWindow.num_bytes = std::min(bytes, l2_window_max);
            attr[0].val.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)l2_persist_bytes / (double)std::min(bytes, l2_window_max));
            attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            cfg.numAttrs = 1;
        }
        Ctx cl = c;
        static const bool no_mirror = getenv("MBPE_NO_CAND_MIRROR") != nullptr;
        cl.m_cap = no_mirror ? 0 : 1; // request the shared-memory candidate mirror (the kernel sets the real capacity)
        note(cudaLaunchKernelEx(&cfg, k_persistent, cl), "k_persistent launch");
        n_launch++;
        note(cudaGetLastError(), "k_persistent launch");
    }
    uint64_t launches() const { return n_launch; }
};

} // namespace mbpe

using namespace mbpe;

// ---------------------------------------------------------------------------------------------------------
// C ABI: trainer
// ---------------------------------------------------------------------------------------------------------
struct mbpe_trainer {
    int device = 0;
    uint32_t *d_tokens = nullptr;
    uint64_t *d_off = nullptr;
    uint32_t *d_weight = nullptr;
    uint64_t n_tokens = 0, n_chunks = 0;
    Ctl *pinned_ctl = nullptr;
    bool pooled = false; // d_* came from the stream-ordered pool (mbpe_trainer_create_device)
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
};

// one pinned control block is kept between trainers: cudaMallocHost / cudaFreeHost per trainer cost milliseconds
static std::mutex g_pinned_mu;
static Ctl *g_pinned_spare = nullptr;
static Ctl *pinned_ctl_take() {
    {
        std::lock_guard<std::mutex> lk(g_pinned_mu);
        if (g_pinned_spare) {
            Ctl *p = g_pinned_spare;
            g_pinned_spare = nullptr;
            return p;
        }
    }
    Ctl *p = nullptr;
    return cudaMallocHost(&p, sizeof(Ctl)) == cudaSuccess ? p : nullptr;
}
static void pinned_ctl_give(Ctl *p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(g_pinned_mu);
        if (!g_pinned_spare) {
            g_pinned_spare = p;
            return;
        }
    }
    cudaFreeHost(p);
}

extern "C" int mbpe_trainer_create(const uint32_t *tokens, uint64_t n_tokens, const uint64_t *chunk_off,
                                   uint64_t n_chunks, const uint32_t *chunk_weight, int device, mbpe_trainer **out) {
    if (!out || (!tokens && n_tokens) || !chunk_off) return set_error(MBPE_E_INVALID, "null argument");
    *out = nullptr;
    if (n_tokens >= (1ull << 30)) return set_error(MBPE_E_INVALID, "n_tokens must be < 2^30 per trainer");
    if (chunk_off[0] != 0 || chunk_off[n_chunks] != n_tokens)
        return set_error(MBPE_E_INVALID, "chunk_off must start at 0 and end at n_tokens");
    for (uint64_t c = 0; c < n_chunks; c++) {
        if (chunk_off[c + 1] < chunk_off[c]) return set_error(MBPE_E_INVALID, "chunk_off not monotonic");
        if (chunk_off[c + 1] - chunk_off[c] >= 2)
            for (uint64_t i = chunk_off[c]; i < chunk_off[c + 1]; i++)
                if (tokens[i] >= 256)
                    return set_error(MBPE_E_INVALID, "token >= 256 inside a multi-token chunk (new ids would collide)");
    }
    int rc = use_device(device);
    if (rc) return rc;
    mbpe_trainer *t = new mbpe_trainer();
    t->device = device;
    t->n_tokens = n_tokens;
    t->n_chunks = n_chunks;
    MB_CUDA(cudaMalloc(&t->d_tokens, std::max<uint64_t>(n_tokens, 1) * 4));
    MB_CUDA(cudaMalloc(&t->d_off, (n_chunks + 1) * 8));
    MB_CUDA(cudaMemcpy(t->d_tokens, tokens, n_tokens * 4, cudaMemcpyHostToDevice));
    MB_CUDA(cudaMemcpy(t->d_off, chunk_off, (n_chunks + 1) * 8, cudaMemcpyHostToDevice));
    if (chunk_weight) {
        MB_CUDA(cudaMalloc(&t->d_weight, std::max<uint64_t>(n_chunks, 1) * 4));
        MB_CUDA(cudaMemcpy(t->d_weight, chunk_weight, n_chunks * 4, cudaMemcpyHostToDevice));
    }
    t->pinned_ctl = pinned_ctl_take();
    if (!t->pinned_ctl) return set_error(MBPE_E_CUDA, "out of pinned memory");
    for (auto &e : t->ev) MB_CUDA(cudaEventCreate(&e));
    // keep freed blocks in the stream-ordered pool so repeated runs do not go back to the driver
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    *out = t;
    return MBPE_OK;
}

// takes the corpus buffers over (the struct is cleared): nothing is copied or allocated
extern "C" int mbpe_trainer_create_device(mbpe_device_corpus *c, mbpe_trainer **out) {
    if (!out || !c || !c->d_off) return set_error(MBPE_E_INVALID, "null argument");
    *out = nullptr;
    if (c->n_tokens >= (1ull << 30)) return set_error(MBPE_E_INVALID, "n_tokens must be < 2^30 per trainer");
    int rc = use_device(c->device);
    if (rc) return rc;
    mbpe_trainer *t = new mbpe_trainer();
    t->device = c->device;
    t->n_tokens = c->n_tokens;
    t->n_chunks = c->n_unique;
    t->d_tokens = c->d_tokens;
    t->d_off = c->d_off;
    t->d_weight = c->d_weight;
    t->pooled = true;
    c->d_tokens = nullptr, c->d_off = nullptr, c->d_weight = nullptr;
    c->n_tokens = c->n_unique = c->n_chunks = 0;
    t->pinned_ctl = pinned_ctl_take();
    if (!t->pinned_ctl) {
        mbpe_trainer_destroy(t);
        return set_error(MBPE_E_CUDA, "out of pinned memory");
    }
    for (auto &e : t->ev) MB_CUDA(cudaEventCreate(&e));
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, t->device) == cudaSuccess) {
        uint64_t keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    *out = t;
    return MBPE_OK;
}

extern "C" void mbpe_trainer_destroy(mbpe_trainer *t) {
    if (!t) return;
    cudaSetDevice(t->device);
    if (t->pooled) { // buffers taken over from a device corpus: back to the stream-ordered pool
        if (t->d_tokens) cudaFreeAsync(t->d_tokens, nullptr);
        if (t->d_off) cudaFreeAsync(t->d_off, nullptr);
        if (t->d_weight) cudaFreeAsync(t->d_weight, nullptr);
    } else {
        cudaFree(t->d_tokens);
        cudaFree(t->d_off);
        cudaFree(t->d_weight);
    }
    pinned_ctl_give(t->pinned_ctl);
    for (auto &e : t->ev)
        if (e) cudaEventDestroy(e);
    delete t;
}

static uint32_t env_u32(const char *name, uint32_t dflt) {
    const char *v = getenv(name);
    return v && *v ? (uint32_t)strtoul(v, nullptr, 10) : dflt;
}

extern "C" int mbpe_trainer_run(mbpe_trainer *t, uint32_t vocab_size, int mode, int engine, void *stream,
                                uint32_t *merges_out, int32_t *counts_out, uint32_t *n_merges_out,
                                mbpe_train_stats *stats) {
    if (!t || !merges_out || !n_merges_out) return set_error(MBPE_E_INVALID, "null argument");
    if (vocab_size < 256) return set_error(MBPE_E_INVALID, "vocab_size must be >= 256 (Tokenizer.h:492)");
    if (mode != MBPE_MODE_FIRST && mode != MBPE_MODE_LEXICAL) return set_error(MBPE_E_INVALID, "bad mode");
    if (engine != MBPE_ENGINE_STEPWISE && engine != MBPE_ENGINE_PERSISTENT) return set_error(MBPE_E_INVALID, "bad engine");
    int rc = use_device(t->device);
    if (rc) return rc;
    CudaBE be;
    be.stream = (cudaStream_t)stream;
    be.sms = sm_count(t->device);
    be.pinned = t->pinned_ctl;
    if (!getenv("MBPE_NO_L2_PERSIST")) {
        int max_persist = 0, max_window = 0;
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, t->device);
        cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, t->device);
        if (max_persist > 0 && max_window > 0 &&
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)max_persist) == cudaSuccess) {
            be.l2_window_max = (size_t)max_window;
            be.l2_persist_bytes = (size_t)max_persist;
        } else {
            cudaGetLastError();
        }
    }
    TrainConfig cfg{vocab_size, mode, engine, env_u32("MBPE_BIG_LIMIT", 16384), env_u32("MBPE_CAND_WANT", 1024),
                    env_u32("MBPE_CAND_LIMIT", PS_SEL * PERSISTENT_THREADS), env_u32("MBPE_INIT_SLOTS", 0), env_u32("MBPE_TRAIN_PF", 0)};
    TrainOutcome o;
    *n_merges_out = 0;
    MB_CUDA(cudaEventRecord(t->ev[0], be.stream));
    int drc;
    {
        TrainLoop<CudaBE> loop(be);
        drc = loop.run(t->d_tokens, t->d_off, t->d_weight, t->n_tokens, t->n_chunks, cfg, merges_out, counts_out, &o);
    }
    MB_CUDA(cudaEventRecord(t->ev[1], be.stream));
    MB_CUDA(cudaStreamSynchronize(be.stream));
    if (be.err != cudaSuccess) return cuda_fail(be.err, be.err_what, __FILE__, __LINE__);
    if (drc) return set_error(MBPE_E_CUDA, "train loop reached an unknown state");
    *n_merges_out = finish_merges(o, vocab_size, mode, merges_out, counts_out);
    if (stats) {
        float ms = 0;
        cudaEventElapsedTime(&ms, t->ev[0], t->ev[1]);
        memset(stats, 0, sizeof *stats);
        stats->gpu_ms = ms;
        stats->n_positions = t->n_tokens;
        stats->n_pairs = o.n_pairs;
        stats->table_slots = o.table_slots;
        stats->n_launches = be.launches();
        stats->n_big_merges = o.n_big;
        stats->n_rebuilds = o.n_rebuilds;
        stats->n_grows = o.n_grows;
        stats->rescan_bytes = o.rescan_bytes;
        for (int i = 0; i < 8; i++) stats->resident_cycles[i] = o.prof[i];
    }
    return MBPE_OK;
}

extern "C" int mbpe_train(const uint32_t *tokens, uint64_t n_tokens, const uint64_t *chunk_off, uint64_t n_chunks,
                          const uint32_t *chunk_weight, uint32_t vocab_size, int mode, uint32_t *merges_out,
                          int32_t *counts_out, uint32_t *n_merges_out) {
    mbpe_trainer *t = nullptr;
    int rc = mbpe_trainer_create(tokens, n_tokens, chunk_off, n_chunks, chunk_weight, 0, &t);
    if (rc) return rc;
    rc = mbpe_trainer_run(t, vocab_size, mode, MBPE_ENGINE_PERSISTENT, nullptr, merges_out, counts_out, n_merges_out,
                          nullptr);
    mbpe_trainer_destroy(t);
    return rc;
}

// ---------------------------------------------------------------------------------------------------------
// C ABI: sharded training over NCCL (one process per GPU)
// ---------------------------------------------------------------------------------------------------------
struct mbpe_comm {
    NcclApi::comm_t nccl = nullptr;
    int rank = 0, world = 1, device = 0;
    // resident exchange (k_persistent_sharded): this rank's inbox + flag words, and the peers' mapped into this process
    XRec *d_inbox = nullptr;
    unsigned long long *d_flags = nullptr;
    void *peer_inbox[XCH_MAX_WORLD] = {}, *peer_flags[XCH_MAX_WORLD] = {};
    PeerLinks links{};
    bool resident_ok = false;
    uint32_t xstep = 0;
};

extern "C" int mbpe_comm_unique_id(uint8_t *id_out) {
    if (!id_out) return set_error(MBPE_E_INVALID, "null argument");
    NcclApi &api = NcclApi::get();
    if (!api.ok) return set_error(MBPE_E_NO_DEVICE, "libnccl.so.2 could not be loaded");
    NcclApi::unique_id id;
    if (api.GetUniqueId(&id) != 0) return set_error(MBPE_E_CUDA, "ncclGetUniqueId failed");
    memcpy(id_out, id.internal, 128);
    return MBPE_OK;
}

// all-gather of `bytes` host bytes per rank through the communicator (setup only)
static int comm_allgather_host(mbpe_comm *c, const void *mine, void *all, size_t bytes) {
    uint8_t *d = nullptr;
    MB_CUDA(cudaMalloc(&d, bytes * ((size_t)c->world + 1)));
    cudaError_t ce = cudaMemcpy(d, mine, bytes, cudaMemcpyHostToDevice);
    int nrc = 0;
    if (ce == cudaSuccess) nrc = NcclApi::get().AllGather(d, d + bytes, bytes, NCCL_UINT8, c->nccl, nullptr);
    if (ce == cudaSuccess && nrc == 0) ce = cudaDeviceSynchronize();
    if (ce == cudaSuccess && nrc == 0) ce = cudaMemcpy(all, d + bytes, bytes * c->world, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (nrc != 0) return set_error(MBPE_E_CUDA, "ncclAllGather failed during communicator setup");
    if (ce != cudaSuccess) return cuda_fail(ce, "communicator setup", __FILE__, __LINE__);
    return MBPE_OK;
}

// Map every peer's inbox into this process (cudaIpc) so that the resident CTA can store its records there directly.
// All ranks agree on the outcome (one more all-gather): either every rank has every mapping, or nobody uses them and
// the sharded trainer stays on the host-driven path.
static int comm_setup_peer_links(mbpe_comm *c) {
    c->resident_ok = false;
    if (c->world < 2 || c->world > (int)XCH_MAX_WORLD || getenv("MBPE_NO_PEER_EXCHANGE")) return MBPE_OK;
    uint32_t cap = 1u << 19; // records per (sender, parity): 8 MB; resident merges have a count <= cap / 4
    if (const char *v = getenv("MBPE_XCH_CAP")) cap = std::max<uint32_t>(1024, (uint32_t)strtoul(v, nullptr, 10));
    const size_t inbox_bytes = (size_t)c->world * 2 * cap * sizeof(XRec), flag_bytes = 4096;
    struct Handles {
        cudaIpcMemHandle_t inbox, flags;
        int ok;
    } mine;
    memset(&mine, 0, sizeof mine);
    mine.ok = cudaMalloc(&c->d_inbox, inbox_bytes) == cudaSuccess && cudaMalloc(&c->d_flags, flag_bytes) == cudaSuccess &&
              cudaMemset(c->d_flags, 0, flag_bytes) == cudaSuccess && cudaDeviceSynchronize() == cudaSuccess &&
              cudaIpcGetMemHandle(&mine.inbox, c->d_inbox) == cudaSuccess && cudaIpcGetMemHandle(&mine.flags, c->d_flags) == cudaSuccess;
    cudaGetLastError();
    std::vector<Handles> all(c->world);
    int rc = comm_allgather_host(c, &mine, all.data(), sizeof(Handles));
    if (rc) return rc;
    int ok = 1;
    for (int r = 0; r < c->world; r++) ok &= all[r].ok;
    for (int r = 0; r < c->world && ok; r++) {
        if (r == c->rank) continue;
        int can = 0;
        // (ranks of one box: the peer's device is visible here under SOME ordinal; lazy peer access enables the mapping)
        ok = cudaIpcOpenMemHandle(&c->peer_inbox[r], all[r].inbox, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess &&
             cudaIpcOpenMemHandle(&c->peer_flags[r], all[r].flags, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
        (void)can;
    }
    cudaGetLastError();
    std::vector<int> oks(c->world);
    rc = comm_allgather_host(c, &ok, oks.data(), sizeof(int));
    if (rc) return rc;
    for (int r = 0; r < c->world; r++) ok &= oks[r];
    if (!ok) return MBPE_OK; // host-driven path only
    PeerLinks &pl = c->links;
    memset(&pl, 0, sizeof pl);
    for (int r = 0; r < c->world; r++) {
        if (r == c->rank) continue;
        // inside peer r's inbox, the block [source = this rank][parity][cap]
        pl.send_to[r] = (XRec *)c->peer_inbox[r] + (size_t)c->rank * 2 * cap;
        pl.flag_to[r] = (unsigned long long *)c->peer_flags[r] + (size_t)c->rank * 2;
    }
    pl.inbox = c->d_inbox;
    pl.flags = c->d_flags;
    pl.cap = cap;
    pl.world = (uint32_t)c->world;
    pl.rank = (uint32_t)c->rank;
    c->resident_ok = true;
    return MBPE_OK;
}

extern "C" int mbpe_comm_create(const uint8_t *id_bytes, int rank, int world, int device, mbpe_comm **out) {
    if (!id_bytes || !out || world < 1 || rank < 0 || rank >= world) return set_error(MBPE_E_INVALID, "bad argument");
    *out = nullptr;
    int rc = use_device(device);
    if (rc) return rc;
    NcclApi &api = NcclApi::get();
    if (!api.ok) return set_error(MBPE_E_NO_DEVICE, "libnccl.so.2 could not be loaded");
    NcclApi::unique_id id;
    memcpy(id.internal, id_bytes, 128);
    mbpe_comm *c = new mbpe_comm();
    c->rank = rank;
    c->world = world;
    c->device = device;
    int nrc = api.CommInitRank(&c->nccl, world, id, rank);
    if (nrc != 0) {
        delete c;
        return set_error(MBPE_E_CUDA, std::string("ncclCommInitRank failed: ") + (api.GetErrorString ? api.GetErrorString(nrc) : "?"));
    }
    if ((rc = comm_setup_peer_links(c))) {
        mbpe_comm_destroy(c);
        return rc;
    }
    *out = c;
    return MBPE_OK;
}

extern "C" int mbpe_comm_resident(const mbpe_comm *c) { return c && c->resident_ok ? 1 : 0; }

extern "C" void mbpe_comm_destroy(mbpe_comm *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (uint32_t r = 0; r < XCH_MAX_WORLD; r++) {
        if (c->peer_inbox[r]) cudaIpcCloseMemHandle(c->peer_inbox[r]);
        if (c->peer_flags[r]) cudaIpcCloseMemHandle(c->peer_flags[r]);
    }
    if (c->nccl) NcclApi::get().CommDestroy(c->nccl); // (a collective teardown: the peers have closed their mappings too)
    cudaFree(c->d_inbox);
    cudaFree(c->d_flags);
    cudaGetLastError();
    delete c;
}

// this rank's share of the corpus, resident; run() can be repeated (the merge loop works on copies)
struct mbpe_sharded_trainer {
    mbpe_comm *comm = nullptr;
    uint32_t *d_tokens = nullptr, *d_weight = nullptr;
    uint64_t *d_off = nullptr;
    uint64_t n_local = 0, n_chunks_local = 0, pos_base = 0, n_global = 0;
    Ctl *pinned = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
};

extern "C" void mbpe_sharded_trainer_destroy(mbpe_sharded_trainer *t) {
    if (!t) return;
    if (t->comm) cudaSetDevice(t->comm->device);
    cudaFree(t->d_tokens);
    cudaFree(t->d_off);
    cudaFree(t->d_weight);
    if (t->pinned) cudaFreeHost(t->pinned);
    if (t->e0) cudaEventDestroy(t->e0);
    if (t->e1) cudaEventDestroy(t->e1);
    cudaGetLastError();
    delete t;
}

static int sharded_trainer_create_impl(mbpe_sharded_trainer *t, mbpe_comm *comm, const uint32_t *tokens, uint64_t n_tokens,
                                       const uint64_t *chunk_off, uint64_t n_chunks, const uint32_t *chunk_weight) {
    // this rank's contiguous, token-balanced share of the unique chunks (same cut on every rank)
    auto first_chunk = [&](int r) -> uint64_t {
        if (r <= 0) return 0;
        if (r >= comm->world) return n_chunks;
        const uint64_t target = n_tokens / comm->world * r;
        uint64_t lo = 0, hi = n_chunks;
        while (lo < hi) {
            uint64_t mid = (lo + hi) / 2;
            if (chunk_off[mid] < target)
                lo = mid + 1;
            else
                hi = mid;
        }
        return lo;
    };
    const uint64_t c0 = first_chunk(comm->rank), c1 = first_chunk(comm->rank + 1);
    const uint64_t t0 = chunk_off[c0], t1 = chunk_off[c1], nl = t1 - t0, ncl = c1 - c0;
    for (uint64_t c = c0; c < c1; c++)
        if (chunk_off[c + 1] - chunk_off[c] >= 2)
            for (uint64_t i = chunk_off[c]; i < chunk_off[c + 1]; i++)
                if (tokens[i] >= 256) return set_error(MBPE_E_INVALID, "token >= 256 inside a multi-token chunk");
    std::vector<uint64_t> loff(ncl + 1);
    for (uint64_t c = c0; c <= c1; c++) loff[c - c0] = chunk_off[c] - t0;
    t->comm = comm;
    t->n_local = nl;
    t->n_chunks_local = ncl;
    t->pos_base = t0;
    t->n_global = n_tokens;
    MB_CUDA(cudaMalloc(&t->d_tokens, std::max<uint64_t>(nl, 1) * 4));
    MB_CUDA(cudaMalloc(&t->d_off, (ncl + 1) * 8));
    MB_CUDA(cudaMemcpy(t->d_tokens, tokens + t0, nl * 4, cudaMemcpyHostToDevice));
    MB_CUDA(cudaMemcpy(t->d_off, loff.data(), (ncl + 1) * 8, cudaMemcpyHostToDevice));
    if (chunk_weight) {
        MB_CUDA(cudaMalloc(&t->d_weight, std::max<uint64_t>(ncl, 1) * 4));
        MB_CUDA(cudaMemcpy(t->d_weight, chunk_weight + c0, ncl * 4, cudaMemcpyHostToDevice));
    }
    MB_CUDA(cudaMallocHost(&t->pinned, sizeof(Ctl)));
    MB_CUDA(cudaEventCreate(&t->e0));
    MB_CUDA(cudaEventCreate(&t->e1));
    cudaMemPool_t pool; // keep freed blocks in the stream-ordered pool: repeated runs do not go back to the driver
    if (cudaDeviceGetDefaultMemPool(&pool, comm->device) == cudaSuccess) {
        uint64_t keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    return MBPE_OK;
}

extern "C" int mbpe_sharded_trainer_create(mbpe_comm *comm, const uint32_t *tokens, uint64_t n_tokens, const uint64_t *chunk_off,
                                           uint64_t n_chunks, const uint32_t *chunk_weight, mbpe_sharded_trainer **out) {
    if (!comm || !out || !chunk_off || (!tokens && n_tokens)) return set_error(MBPE_E_INVALID, "null argument");
    *out = nullptr;
    if (n_tokens >= (1ull << 30)) return set_error(MBPE_E_INVALID, "n_tokens must be < 2^30");
    if (chunk_off[0] != 0 || chunk_off[n_chunks] != n_tokens)
        return set_error(MBPE_E_INVALID, "chunk_off must start at 0 and end at n_tokens");
    int rc = use_device(comm->device);
    if (rc) return rc;
    mbpe_sharded_trainer *t = new mbpe_sharded_trainer();
    if ((rc = sharded_trainer_create_impl(t, comm, tokens, n_tokens, chunk_off, n_chunks, chunk_weight))) {
        mbpe_sharded_trainer_destroy(t);
        return rc;
    }
    *out = t;
    return MBPE_OK;
}

extern "C" int mbpe_sharded_trainer_run(mbpe_sharded_trainer *t, uint32_t vocab_size, int mode, int engine, void *stream,
                                        uint32_t *merges_out, int32_t *counts_out, uint32_t *n_merges_out, mbpe_train_stats *stats) {
    if (!t || !merges_out || !n_merges_out) return set_error(MBPE_E_INVALID, "null argument");
    if (vocab_size < 256) return set_error(MBPE_E_INVALID, "vocab_size must be >= 256 (Tokenizer.h:492)");
    if (mode != MBPE_MODE_FIRST && mode != MBPE_MODE_LEXICAL) return set_error(MBPE_E_INVALID, "bad mode");
    if (engine != MBPE_ENGINE_STEPWISE && engine != MBPE_ENGINE_PERSISTENT) return set_error(MBPE_E_INVALID, "bad engine");
    mbpe_comm *comm = t->comm;
    int rc = use_device(comm->device);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    CudaBE be;
    be.stream = st;
    be.sms = sm_count(comm->device);
    be.pinned = t->pinned;
    be.nccl = comm->nccl;
    be.comm_world = (uint32_t)comm->world;
    be.comm_rank = (uint32_t)comm->rank;
    if (comm->resident_ok && engine == MBPE_ENGINE_PERSISTENT) {
        be.links = &comm->links;
        be.xstep_ptr = &comm->xstep;
    }
    TrainConfig cfg{vocab_size, mode, engine, ~0u, env_u32("MBPE_CAND_WANT", 512), env_u32("MBPE_CAND_LIMIT", PS_SEL * PERSISTENT_THREADS), 0, env_u32("MBPE_TRAIN_PF", 0)};
    TrainOutcome o;
    *n_merges_out = 0;
    MB_CUDA(cudaEventRecord(t->e0, st));
    int drc;
    {
        TrainLoopSharded<CudaBE> loop(be);
        drc = loop.run(t->d_tokens, t->d_off, t->d_weight, t->n_local, t->n_chunks_local, t->pos_base, t->n_global, cfg, merges_out,
                       counts_out, &o);
    }
    MB_CUDA(cudaEventRecord(t->e1, st));
    MB_CUDA(cudaStreamSynchronize(st));
    float ms = 0;
    cudaEventElapsedTime(&ms, t->e0, t->e1);
    cudaFree(be.d_all);
    cudaFree(be.d_counts);
    cudaFree(be.d_my_count);
    if (be.err != cudaSuccess) return cuda_fail(be.err, be.err_what, __FILE__, __LINE__);
    if (drc == -2) return set_error(MBPE_E_CUDA, "sharded train: a peer did not answer inside the resident kernel (ranks out of step?)");
    if (drc) return set_error(MBPE_E_CUDA, "sharded train loop reached an unexpected state");
    *n_merges_out = finish_merges(o, vocab_size, mode, merges_out, counts_out);
    if (stats) {
        memset(stats, 0, sizeof *stats);
        stats->gpu_ms = ms;
        stats->n_positions = t->n_local;
        stats->n_pairs = o.n_pairs;
        stats->table_slots = o.table_slots;
        stats->n_launches = be.launches();
        stats->n_big_merges = o.n_big; // merges driven from the host (grid kernels + one all-gather each)
        stats->n_rebuilds = o.n_rebuilds;
        stats->rescan_bytes = o.rescan_bytes;
        for (int i = 0; i < 8; i++) stats->resident_cycles[i] = o.prof[i];
    }
    return MBPE_OK;
}

extern "C" int mbpe_train_sharded(mbpe_comm *comm, const uint32_t *tokens, uint64_t n_tokens, const uint64_t *chunk_off,
                                  uint64_t n_chunks, const uint32_t *chunk_weight, uint32_t vocab_size, int mode,
                                  void *stream, uint32_t *merges_out, int32_t *counts_out, uint32_t *n_merges_out,
                                  mbpe_train_stats *stats) {
    mbpe_sharded_trainer *t = nullptr;
    int rc = mbpe_sharded_trainer_create(comm, tokens, n_tokens, chunk_off, n_chunks, chunk_weight, &t);
    if (rc) return rc;
    const int engine = getenv("MBPE_SHARDED_STEPWISE") ? MBPE_ENGINE_STEPWISE : MBPE_ENGINE_PERSISTENT;
    rc = mbpe_sharded_trainer_run(t, vocab_size, mode, engine, stream, merges_out, counts_out, n_merges_out, stats);
    mbpe_sharded_trainer_destroy(t);
    return rc;
}

// ---------------------------------------------------------------------------------------------------------
// Sharded train front end: every rank pre-tokenises and deduplicates ITS part of the text, one all-gather moves the
// unique chunks, every rank merges them into the same corpus and runs the merge loop (Tokenizer.h:500-589 as a whole).
// ---------------------------------------------------------------------------------------------------------
namespace mbpe {
__global__ void k_pack_corpus(const uint32_t *tokens, uint64_t n_tokens, const uint64_t *off, uint64_t n_unique, uint8_t *bytes,
                              uint32_t *off32) {
    const uint64_t i0 = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = i0; i < n_tokens; i += stride) bytes[i] = (uint8_t)tokens[i]; // chunk tokens are bytes before training
    for (uint64_t i = i0; i <= n_unique; i += stride) off32[i] = (uint32_t)off[i];
}
} // namespace mbpe

extern "C" int mbpe_train_text_sharded(mbpe_comm *comm, mbpe_pretok *p, const uint8_t *text_part, uint64_t len, uint32_t vocab_size,
                                       int mode, uint32_t *merges_out, int32_t *counts_out, uint32_t *n_merges_out,
                                       mbpe_train_stats *stats, double *front_end_s) {
    if (!comm || !p || !merges_out || !n_merges_out || (len && !text_part)) return set_error(MBPE_E_INVALID, "null argument");
    int rc = use_device(comm->device);
    if (rc) return rc;
    const auto t_start = std::chrono::steady_clock::now();
    const int W = comm->world;
    mbpe_device_corpus mine{};
    rc = mbpe_pretok_corpus(p, text_part, len, &mine);
    // the outcome of the local front end is agreed on before any collective that depends on it
    struct Sizes {
        uint64_t n_tokens, n_unique;
        int64_t rc;
    } sz{mine.n_tokens, mine.n_unique, rc};
    std::vector<Sizes> all(W);
    int grc = comm_allgather_host(comm, &sz, all.data(), sizeof(Sizes));
    if (grc) {
        mbpe_device_corpus_free(&mine);
        return grc;
    }
    for (int r = 0; r < W; r++)
        if (all[r].rc) {
            mbpe_device_corpus_free(&mine);
            return rc ? rc : set_error((int)all[r].rc, "the device front end failed on another rank");
        }
    uint64_t max_t = 0, max_u = 0;
    for (int r = 0; r < W; r++) {
        if (all[r].n_tokens >= (1ull << 32) - 64) {
            mbpe_device_corpus_free(&mine);
            return set_error(MBPE_E_INVALID, "a rank's unique chunks exceed 4 GiB");
        }
        max_t = std::max(max_t, all[r].n_tokens);
        max_u = std::max(max_u, all[r].n_unique);
    }
    max_t = (max_t + 63) & ~63ull; // every rank's block starts 64-byte aligned
    const uint64_t off_stride = (max_u + 1 + 15) & ~15ull, w_stride = (max_u + 15) & ~15ull;
    uint8_t *d_bytes = nullptr;
    uint32_t *d_off = nullptr, *d_w = nullptr;
    auto release = [&]() {
        cudaFree(d_bytes);
        cudaFree(d_off);
        cudaFree(d_w);
        mbpe_device_corpus_free(&mine);
    };
    cudaError_t ce = cudaMalloc(&d_bytes, (size_t)(W + 1) * max_t + 64);
    if (ce == cudaSuccess) ce = cudaMalloc(&d_off, (size_t)(W + 1) * off_stride * 4);
    if (ce == cudaSuccess) ce = cudaMalloc(&d_w, (size_t)(W + 1) * std::max<uint64_t>(w_stride, 16) * 4);
    if (ce != cudaSuccess) {
        release();
        return cuda_fail(ce, "sharded front end buffers", __FILE__, __LINE__);
    }
    // block W = this rank's send buffers, blocks 0..W-1 = everybody's
    uint8_t *s_bytes = d_bytes + (size_t)W * max_t;
    uint32_t *s_off = d_off + (size_t)W * off_stride, *s_w = d_w + (size_t)W * std::max<uint64_t>(w_stride, 16);
    k_pack_corpus<<<sm_count(comm->device) * 8, 256>>>(mine.d_tokens, mine.n_tokens, mine.d_off, mine.n_unique, s_bytes, s_off);
    ce = cudaGetLastError();
    if (ce == cudaSuccess && mine.n_unique) ce = cudaMemcpyAsync(s_w, mine.d_weight, mine.n_unique * 4, cudaMemcpyDeviceToDevice, nullptr);
    NcclApi &api = NcclApi::get();
    int nrc = 0;
    if (ce == cudaSuccess && max_t) nrc = api.AllGather(s_bytes, d_bytes, max_t, NCCL_UINT8, comm->nccl, nullptr);
    if (ce == cudaSuccess && !nrc) nrc = api.AllGather(s_off, d_off, off_stride * 4, NCCL_UINT8, comm->nccl, nullptr);
    if (ce == cudaSuccess && !nrc && w_stride) nrc = api.AllGather(s_w, d_w, w_stride * 4, NCCL_UINT8, comm->nccl, nullptr);
    if (ce == cudaSuccess && !nrc) ce = cudaDeviceSynchronize();
    if (ce != cudaSuccess || nrc) {
        release();
        return nrc ? set_error(MBPE_E_CUDA, "ncclAllGather of the unique chunks failed") : cuda_fail(ce, "sharded front end", __FILE__, __LINE__);
    }
    std::vector<const uint8_t *> texts(W);
    std::vector<const uint32_t *> offs(W), ws(W);
    std::vector<uint64_t> nbytes(W), nchunks(W);
    for (int r = 0; r < W; r++) {
        texts[r] = d_bytes + (size_t)r * max_t;
        offs[r] = d_off + (size_t)r * off_stride;
        ws[r] = d_w + (size_t)r * w_stride;
        nbytes[r] = max_t + 64; // readable: the blocks are contiguous and the buffer has slack at its end
        nchunks[r] = all[r].n_unique;
    }
    mbpe_device_corpus merged{};
    rc = mbpe_pretok_merge_corpora(p, texts.data(), nbytes.data(), offs.data(), ws.data(), nchunks.data(), (uint32_t)W, &merged, nullptr);
    uint64_t n_chunks_total = 0;
    for (int r = 0; r < W; r++) n_chunks_total += all[r].n_unique;
    release();
    if (rc) return rc;
    if (front_end_s) *front_end_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
    mbpe_trainer *tr = nullptr;
    rc = mbpe_trainer_create_device(&merged, &tr);
    mbpe_device_corpus_free(&merged);
    if (rc) return rc;
    rc = mbpe_trainer_run(tr, vocab_size, mode, MBPE_ENGINE_PERSISTENT, nullptr, merges_out, counts_out, n_merges_out, stats);
    mbpe_trainer_destroy(tr);
    return rc;
}

// ---------------------------------------------------------------------------------------------------------
// C ABI: PairCount seam (PairCount.h:27-47). Slot.first doubles as insert_order (index of the first add that
// named the pair == next_insert++ order, PairCount.h:149).
// ---------------------------------------------------------------------------------------------------------
namespace mbpe {

struct PcBest {
    int32_t cnt;
    uint32_t valid;
    uint64_t tie;
    uint64_t key;
};
__device__ __forceinline__ bool pc_better(const PcBest &x, const PcBest &y) { // x ranks before y
    if (!x.valid) return false;
    if (!y.valid) return true;
    if (x.cnt != y.cnt) return x.cnt > y.cnt;
    return x.tie < y.tie;
}
__device__ __forceinline__ PcBest pc_shfl_down(const PcBest &v, int d) {
    PcBest r;
    r.cnt = __shfl_down_sync(0xffffffffu, v.cnt, d);
    r.valid = __shfl_down_sync(0xffffffffu, v.valid, d);
    r.tie = __shfl_down_sync(0xffffffffu, v.tie, d);
    r.key = __shfl_down_sync(0xffffffffu, v.key, d);
    return r;
}

__global__ void k_pc_add(Slot *slot, uint32_t cap_mask, const uint32_t *a, const uint32_t *b, const int32_t *delta,
                         uint64_t n, uint32_t order_base, uint32_t *n_pairs) {
    Ctx c{};
    c.slot = slot;
    c.cap_mask = cap_mask;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        bool created;
        uint32_t s = slot_upsert(c, pair_key(a[i], b[i]), &created);
        if (created) atomicAdd(n_pairs, 1u);
        atomicAdd(&slot[s].cnt, delta[i]);
        atomicMin(&slot[s].first, order_base + (uint32_t)i);
    }
}

// block-then-grid arg-max over the table: thread -> warp shuffle -> block shared memory -> per-block result;
// the last block to finish reduces the per-block results (threadfence + ticket).
__global__ void __launch_bounds__(256) k_pc_top(const Slot *slot, uint32_t cap, int mode, PcBest *block_best,
                                                unsigned *ticket, PcBest *result) {
    __shared__ PcBest sh[8];
    __shared__ bool last;
    PcBest best{0, 0, 0, 0};
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < cap; s += gridDim.x * blockDim.x) {
        uint64_t key = slot[s].key;
        if (key == EMPTY_KEY) continue;
        PcBest v{slot[s].cnt, 1u, mode == MBPE_MODE_LEXICAL ? key : (uint64_t)slot[s].first, key};
        if (pc_better(v, best)) best = v;
    }
    auto block_reduce = [&](PcBest v) {
        for (int d = 16; d > 0; d >>= 1) {
            PcBest o = pc_shfl_down(v, d);
            if (pc_better(o, v)) v = o;
        }
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x < 32) {
            v = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : PcBest{0, 0, 0, 0};
            for (int d = 4; d > 0; d >>= 1) {
                PcBest o = pc_shfl_down(v, d);
                if (pc_better(o, v)) v = o;
            }
        }
        return v;
    };
    best = block_reduce(best);
    if (threadIdx.x == 0) {
        block_best[blockIdx.x] = best;
        __threadfence();
        last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (last) {
        __threadfence();
        PcBest v{0, 0, 0, 0};
        for (uint32_t i = threadIdx.x; i < gridDim.x; i += blockDim.x) {
            PcBest o = block_best[i];
            if (pc_better(o, v)) v = o;
        }
        __syncthreads();
        v = block_reduce(v);
        if (threadIdx.x == 0) {
            *result = v;
            *ticket = 0;
        }
    }
}

__global__ void k_pc_get(const Slot *slot, uint32_t cap_mask, uint64_t key, int32_t *out /* cnt, found */) {
    Ctx c{};
    c.slot = const_cast<Slot *>(slot);
    c.cap_mask = cap_mask;
    uint32_t s = slot_find(c, key);
    out[0] = s == NIL ? 0 : slot[s].cnt;
    out[1] = s != NIL;
}

} // namespace mbpe

struct mbpe_paircount {
    int device = 0, mode = 0;
    Slot *slot = nullptr;
    uint32_t cap = 0;
    uint32_t *d_n_pairs = nullptr;
    uint64_t n_added = 0;
    PcBest *d_block_best = nullptr, *d_result = nullptr;
    unsigned *d_ticket = nullptr;
    int32_t *d_get = nullptr;
    int sms = 148;
};

static int pc_alloc_table(mbpe_paircount *pc, uint32_t cap) {
    MB_CUDA(cudaMalloc(&pc->slot, (uint64_t)cap * sizeof(Slot)));
    pc->cap = cap;
    k_par<PhClearSlots><<<std::min<uint32_t>((cap + 255) / 256, pc->sms * 8), 256>>>(PhClearSlots{pc->slot, cap});
    MB_CUDA(cudaGetLastError());
    return MBPE_OK;
}

extern "C" int mbpe_paircount_create(int mode, int device, mbpe_paircount **out) {
    if (!out) return set_error(MBPE_E_INVALID, "null argument");
    *out = nullptr;
    if (mode != MBPE_MODE_FIRST && mode != MBPE_MODE_LEXICAL) return set_error(MBPE_E_INVALID, "bad mode");
    int rc = use_device(device);
    if (rc) return rc;
    mbpe_paircount *pc = new mbpe_paircount();
    pc->device = device;
    pc->mode = mode;
    pc->sms = sm_count(device);
    if ((rc = pc_alloc_table(pc, 1024))) return rc;
    MB_CUDA(cudaMalloc(&pc->d_n_pairs, 4));
    MB_CUDA(cudaMemset(pc->d_n_pairs, 0, 4));
    MB_CUDA(cudaMalloc(&pc->d_block_best, sizeof(PcBest) * pc->sms * 8));
    MB_CUDA(cudaMalloc(&pc->d_result, sizeof(PcBest)));
    MB_CUDA(cudaMalloc(&pc->d_ticket, 4));
    MB_CUDA(cudaMemset(pc->d_ticket, 0, 4));
    MB_CUDA(cudaMalloc(&pc->d_get, 8));
    *out = pc;
    return MBPE_OK;
}

extern "C" void mbpe_paircount_destroy(mbpe_paircount *pc) {
    if (!pc) return;
    cudaSetDevice(pc->device);
    cudaFree(pc->slot);
    cudaFree(pc->d_n_pairs);
    cudaFree(pc->d_block_best);
    cudaFree(pc->d_result);
    cudaFree(pc->d_ticket);
    cudaFree(pc->d_get);
    delete pc;
}

extern "C" int mbpe_paircount_size(mbpe_paircount *pc, uint64_t *n_pairs) {
    if (!pc || !n_pairs) return set_error(MBPE_E_INVALID, "null argument");
    int rc = use_device(pc->device);
    if (rc) return rc;
    uint32_t v = 0;
    MB_CUDA(cudaMemcpy(&v, pc->d_n_pairs, 4, cudaMemcpyDeviceToHost));
    *n_pairs = v;
    return MBPE_OK;
}

extern "C" int mbpe_paircount_add(mbpe_paircount *pc, const uint32_t *a, const uint32_t *b, const int32_t *delta,
                                  uint64_t n) {
    if (!pc || (n && (!a || !b || !delta))) return set_error(MBPE_E_INVALID, "null argument");
    if (n == 0) return MBPE_OK;
    if (pc->n_added + n >= 0xFFFFFFFFull) return set_error(MBPE_E_INVALID, "insert_order space exhausted");
    int rc = use_device(pc->device);
    if (rc) return rc;
    uint64_t cur = 0;
    if ((rc = mbpe_paircount_size(pc, &cur))) return rc;
    if ((cur + n) * 2 > pc->cap) { // keep load <= 1/2 even if every record names a new pair
        Slot *old = pc->slot;
        uint32_t old_cap = pc->cap, cap = pc->cap;
        while ((cur + n) * 2 > cap) cap *= 2;
        if ((rc = pc_alloc_table(pc, cap))) return rc;
        Ctx c{};
        c.slot = pc->slot;
        c.cap_mask = cap - 1;
        k_par<PhRehash><<<std::min<uint32_t>((old_cap + 255) / 256, pc->sms * 8), 256>>>(PhRehash{c, old, old_cap});
        MB_CUDA(cudaGetLastError());
        MB_CUDA(cudaDeviceSynchronize());
        MB_CUDA(cudaFree(old));
    }
    uint32_t *da, *db;
    int32_t *dd;
    MB_CUDA(cudaMalloc(&da, n * 4));
    MB_CUDA(cudaMalloc(&db, n * 4));
    MB_CUDA(cudaMalloc(&dd, n * 4));
    MB_CUDA(cudaMemcpy(da, a, n * 4, cudaMemcpyHostToDevice));
    MB_CUDA(cudaMemcpy(db, b, n * 4, cudaMemcpyHostToDevice));
    MB_CUDA(cudaMemcpy(dd, delta, n * 4, cudaMemcpyHostToDevice));
    unsigned grid = (unsigned)std::min<uint64_t>((n + 255) / 256, (uint64_t)pc->sms * 8);
    k_pc_add<<<grid, 256>>>(pc->slot, pc->cap - 1, da, db, dd, n, (uint32_t)pc->n_added, pc->d_n_pairs);
    MB_CUDA(cudaGetLastError());
    MB_CUDA(cudaDeviceSynchronize());
    pc->n_added += n;
    cudaFree(da);
    cudaFree(db);
    cudaFree(dd);
    return MBPE_OK;
}

extern "C" int mbpe_paircount_top(mbpe_paircount *pc, uint32_t *a, uint32_t *b, int32_t *count, int *found) {
    if (!pc || !a || !b || !count || !found) return set_error(MBPE_E_INVALID, "null argument");
    int rc = use_device(pc->device);
    if (rc) return rc;
    unsigned grid = std::min<uint32_t>((pc->cap + 255) / 256, pc->sms * 8);
    k_pc_top<<<grid, 256>>>(pc->slot, pc->cap, pc->mode, pc->d_block_best, pc->d_ticket, pc->d_result);
    MB_CUDA(cudaGetLastError());
    PcBest r;
    MB_CUDA(cudaMemcpy(&r, pc->d_result, sizeof r, cudaMemcpyDeviceToHost));
    *found = r.valid != 0;
    *a = (uint32_t)(r.key >> 32);
    *b = (uint32_t)r.key;
    *count = r.cnt;
    return MBPE_OK;
}

extern "C" int mbpe_paircount_get(mbpe_paircount *pc, uint32_t a, uint32_t b, int32_t *count, int *found) {
    if (!pc || !count || !found) return set_error(MBPE_E_INVALID, "null argument");
    int rc = use_device(pc->device);
    if (rc) return rc;
    k_pc_get<<<1, 1>>>(pc->slot, pc->cap - 1, pair_key(a, b), pc->d_get);
    MB_CUDA(cudaGetLastError());
    int32_t r[2];
    MB_CUDA(cudaMemcpy(r, pc->d_get, 8, cudaMemcpyDeviceToHost));
    *count = r[0];
    *found = r[1];
    return MBPE_OK;
}
