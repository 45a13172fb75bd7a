// encode_hot.cuh -- k_encode_hot: the encode merge scan with the hottest chunks answered from shared memory
// (included by encode.cu only; same job and same arguments as k_encode_tiles, which stays the general-purpose path).
//
// Replaces internal_internal_encode / internal_encode + flatten (Tokenizer.h:325-377, :714-717) like k_encode_tiles.
//
// Why a second kernel. ncu on k_encode_tiles (profiles/r2): 9.4 warp instructions and ~3 L1 wavefronts per chunk, of which
// one whole wavefront-pair per chunk is the random 32-byte probe of the chunk cache -- 32 lanes, 32 different lines, ~66
// L1 cycles per warp probe. At 2 SM cycles per chunk (= 40 % of the HBM roofline) that alone is the whole L1 budget. Text
// is Zipfian, so this kernel answers most chunks from a direct-mapped table of the 8192 most frequent short chunks that
// every CTA keeps in shared memory (16 bytes per entry: 8 key bytes, up to three ids), one LDS.128 per chunk; only the
// remaining chunks go to the chunk cache in HBM, and only what that does not hold goes through the scan.
//
// Shape: ONE persistent CTA per SM; its warps never meet at a block barrier. A warp tile = 32 x CPT consecutive chunks, a
// CTA tile = the NW consecutive warp tiles of one CTA, handed out statically: CTA b works on CTA tiles b, b + G, b + 2G ...
// (every CTA is resident, so the look-back over CTA tiles cannot deadlock, and no ticket is needed).
//   1. boundaries and text words come straight from global memory (coalesced; the lines are shared by neighbouring lanes
//      through L1): key = the chunk's first 8 bytes, one multiplicative hash gives both the shared-memory slot and the home
//      slot in the chunk cache.
//   2. lanes the hot table did not answer probe the chunk cache's home slot (keys of up to 16 bytes, one 256-bit load);
//      everything else -- a home slot taken by another chunk, 17..64 bytes, not cached at all -- takes hot_slow_chunk: the
//      rest of the probe sequence, then the multi-pass scan itself (special tokens first), logged for k_cache_insert.
//   3. packed warp scans give every chunk its place in the warp tile; the ids are gathered in the warp's staging buffer.
//   4. the warp adds its count to the CTA tile (shared-memory counter). The LAST warp to arrive scans the NW counts,
//      announces the CTA tile, walks the predecessors (decoupled look-back, lookback.cuh) and publishes every warp's base.
//   5. the STORE of a tile is deferred by one tile: a warp probes tile i + 1 before it needs the base of tile i, so
//      nobody waits for the look-back (two staging buffers per warp).
#pragma once
#include "encode_tiles.cuh"

namespace mbpe {

#ifdef MBPE_HOT_STATS // build with MBPE_DEFS=-DMBPE_HOT_STATS for the per-step statistics of k_encode_hot (MBPE_DEBUG=1 prints them)
constexpr bool HOT_STATS = true;
#else
constexpr bool HOT_STATS = false;
#endif
constexpr int HOT_LOG2 = 13;
constexpr int HOT_N = 1 << HOT_LOG2; // entries of the shared-memory table (128 KB)
constexpr uint32_t HOT_COUNT_LOG2 = 20;
// Entry = {key lo, key hi, ids lo, ids hi}: key = the chunk's bytes (1..8, last byte non-zero, zero padded -- so the key
// alone says how long the chunk is); ids = up to three ids, each stored PLUS ONE in 21 bits, so the count is the number of
// non-zero fields. An unused entry has key 0 (no eligible chunk has it).
constexpr uint64_t HOT_ID_MASK = (1ull << CACHE_ID_BITS) - 1;

constexpr uint32_t HK_INLINE = 0, HK_HOT = 1, HK_SLOT = 2, HK_PARKED = 3; // where a chunk's ids are (bits 8..9 beside the count)

template <int THREADS, int CPT>
struct HotSmemT {
    static constexpr int NW = THREADS / 32;
    static constexpr int TW = 32 * CPT;  // chunks of a warp tile
    static constexpr int STG = TW * 3;   // ids staged per warp tile (average ~2.1 per chunk); more: stored directly
    static constexpr int TILE = NW * TW; // chunks of a CTA tile
    uint4 hot[HOT_N];
    alignas(16) uint32_t stage[2][NW][STG];
    unsigned long long wbase[2][NW]; // base of every warp tile of CTA tile (parity), published by the last warp to arrive
    uint32_t wsum[2][NW];
    uint32_t cnt[2];
    uint32_t flag[2]; // iteration + 1 whose wbase is valid
};

struct NoStage {
    uint32_t text[1];
};

// ---------------------------------------------------------------------------------------------------------
// The slow path of one chunk, by its own lane (the others of the warp wait): the whole probe sequence of the chunk cache
// for any key length, then the scan. Ids of a scanned chunk go to `cell` (64 words in HBM owned by this chunk of the warp
// tile). Returns count | kind << 8, and in `pay` the inline ids / the cache slot.
// ---------------------------------------------------------------------------------------------------------
__device__ __noinline__ uint32_t hot_slow_chunk(const EncArgs &a, uint32_t o, uint32_t len, uint32_t *cell, uint64_t chunk_index,
                                                uint64_t &pay) {
    pay = 0;
    if (len == 0) return 0;
    if (len > ENC_SHORT_MAX) { // optimistic launch: report it, the host runs the long path and repeats with k_encode_tiles
        const uint32_t q = atomicAdd(a.n_long, 1u);
        if (q < a.long_cap) a.long_list[q] = (uint32_t)chunk_index;
        return 0;
    }
    NoStage ns;
    if (a.cache.slots && len <= CACHE_MAX_LEN) {
        uint64_t key[4];
        big_key(a, ns, false, 0u, o, len, key);
        uint32_t h = cache_hash(key[0], key[1], key[2], key[3]) >> a.cache.shift;
        const bool short_key = len <= CACHE_SHORT_KEY;
        for (uint32_t probes = 0; probes < CACHE_MAX_PROBES; probes++) {
            uint64_t q0, q1, q2, q3;
            ld_sector256<0>(&a.cache.slots[h], q0, q1, q2, q3);
            if (q3 == 0) break;
            if (short_key) {
                if (q0 == key[0] && q1 == key[1] && (q3 >> 56) == len) {
                    const uint32_t n = (uint32_t)q3 & 0x7Fu;
                    if (!(q3 & CACHE_NOT_INLINE)) {
                        pay = q2;
                        return n | (HK_INLINE << 8);
                    }
                    pay = h;
                    return n | (HK_SLOT << 8);
                }
            } else if (q0 == key[0] && q1 == key[1] && q2 == key[2] && (q3 & ~CACHE_LONG_N_MASK) == key[3]) {
                pay = h;
                return ((uint32_t)(q3 >> CACHE_LONG_N_SHIFT) & 0xFFu) | (HK_SLOT << 8);
            }
            h = (h + 1) & a.cache.mask;
        }
    }
    atomicAdd(a.miss_count, 1u);
    const uint32_t sid = special_match(a.sp, len, [&](uint32_t i) { return __ldg(&a.bytes[o + i]); });
    if (sid != ENC_NONE) {
        __stcg(cell, sid);
        return 1u | (HK_PARKED << 8);
    }
    uint32_t t[ENC_SHORT_MAX];
    for (uint32_t i = 0; i < len; i++) t[i] = __ldg(&a.bytes[o + i]);
    uint32_t n = len;
    bool merged = true;
    while (merged && n >= 2) n = enc_pass(a.tab, t, n, merged);
    log_scanned(a, ns, false, 0u, o, len, t, n);
    for (uint32_t i = 0; i < n; i++) __stcg(cell + i, t[i]);
    return n | (HK_PARKED << 8);
}

// The ids of a lane's CPT chunks -> dst + loc[j] (`dst` = the warp's staging buffer in shared memory, or the stream itself).
// Ids that ride in registers are written first; ids in the value sector of a cache slot are fetched PIF slots at a time.
template <int CPT, int PIF>
__device__ __forceinline__ void hot_emit_all(const EncArgs &a, const uint32_t *nk, const uint64_t *pay, const uint32_t *loc, uint32_t *dst,
                                             const uint32_t *cells, uint32_t lane) {
    bool other = false;
#pragma unroll
    for (int j = 0; j < CPT; j++) {
        const uint32_t n = nk[j] & 0xFFu, kind = nk[j] >> 8;
        if (kind <= HK_HOT) { // three ids of 21 bits (HK_HOT: each plus one)
            uint32_t *d = dst + loc[j];
            if (n > 0) d[0] = ((uint32_t)pay[j] & (uint32_t)HOT_ID_MASK) - kind;
            if (n > 1) d[1] = ((uint32_t)(pay[j] >> CACHE_ID_BITS) & (uint32_t)HOT_ID_MASK) - kind;
            if (n > 2) d[2] = (uint32_t)(pay[j] >> (2 * CACHE_ID_BITS)) - kind;
        } else if (n) {
            other = true;
        }
    }
    if (!__any_sync(0xffffffffu, other)) return;
#pragma unroll
    for (int g = 0; g < CPT; g += PIF) {
        uint64_t q0[PIF], q1[PIF], q2[PIF], q3[PIF];
#pragma unroll
        for (int q = 0; q < PIF; q++) {
            q0[q] = q1[q] = q2[q] = q3[q] = 0;
            if ((nk[g + q] >> 8) == HK_SLOT && (nk[g + q] & 0xFFu))
                ld_sector256<0>(reinterpret_cast<const uint8_t *>(&a.cache.slots[(uint32_t)pay[g + q]]) + 32, q0[q], q1[q], q2[q], q3[q]);
        }
#pragma unroll
        for (int q = 0; q < PIF; q++) {
            const int j = g + q;
            const uint32_t n = nk[j] & 0xFFu, kind = nk[j] >> 8;
            uint32_t *d = dst + loc[j];
            if (kind == HK_SLOT && n) {
                if (n <= CACHE_INLINE_IDS) {
                    d[0] = (uint32_t)(q0[q] >> 32);
                    if (n > 1) d[1] = (uint32_t)q1[q];
                    if (n > 2) d[2] = (uint32_t)(q1[q] >> 32);
                    if (n > 3) d[3] = (uint32_t)q2[q];
                    if (n > 4) d[4] = (uint32_t)(q2[q] >> 32);
                    if (n > 5) d[5] = (uint32_t)q3[q];
                    if (n > 6) d[6] = (uint32_t)(q3[q] >> 32);
                } else {
                    const uint32_t *src = a.cache.arena + (uint32_t)(q0[q] >> 32);
#pragma unroll 1
                    for (uint32_t i = 0; i < n; i++) d[i] = __ldg(&src[i]);
                }
            } else if (kind == HK_PARKED && n) {
                const uint32_t *cell = cells + (size_t)(j * 32 + lane) * ENC_SHORT_MAX;
#pragma unroll 1
                for (uint32_t i = 0; i < n; i++) d[i] = __ldcg(cell + i);
            }
        }
    }
}

// The lookup of an open chunk goes on, inline (the lanes of a warp that need it run it together): the rest of the probe
// sequence for a short key whose home slot holds another chunk, the whole sequence for a key of 17..30 bytes. True when the
// chunk is cached (nk, pay set). Out of line: its registers are not charged to the fast path.
__device__ __noinline__ bool hot_probe_more(const EncArgs &a, const uint32_t *words, uint32_t wlast, uint32_t o0, uint32_t len, uint64_t k0,
                                               uint64_t k1, uint32_t h, uint32_t &nk, uint64_t &pay) {
    const bool short_key = len <= CACHE_SHORT_KEY;
    uint64_t k2 = 0, k3 = (uint64_t)len << 56;
    uint32_t probes = 0;
    if (short_key) {
        h = (h + 1) & a.cache.mask;
        probes = 1;
    } else {
        const uint32_t wi = o0 >> 2, sh = (o0 & 3) * 8;
        uint32_t t[7];
#pragma unroll
        for (int q = 0; q < 7; q++) t[q] = __ldg(&words[min(wi + 2 + q, wlast)]);
        k1 = ((uint64_t)__funnelshift_r(t[1], t[2], sh) << 32) | __funnelshift_r(t[0], t[1], sh);
        const uint64_t v2 = ((uint64_t)__funnelshift_r(t[3], t[4], sh) << 32) | __funnelshift_r(t[2], t[3], sh);
        const uint64_t v3 = ((uint64_t)__funnelshift_r(t[5], t[6], sh) << 32) | __funnelshift_r(t[4], t[5], sh);
        const uint32_t nb2 = len - 16; // 1..14
        k2 = nb2 >= 8 ? v2 : v2 & (~0ull >> ((8 - nb2) * 8));
        if (len > 24) k3 |= v3 & (~0ull >> ((32 - len) * 8));
        h = cache_hash(k0, k1, k2, k3) >> a.cache.shift;
    }
#pragma unroll 1
    for (; probes < CACHE_MAX_PROBES; probes++) {
        uint64_t q0, q1, q2, q3;
        ld_sector256<0>(&a.cache.slots[h], q0, q1, q2, q3);
        if (q3 == 0) return false;
        if (q0 == k0 && q1 == k1 && (short_key ? (q3 >> 56) == len : (q2 == k2 && (q3 & ~CACHE_LONG_N_MASK) == k3))) {
            if (short_key && !(q3 & CACHE_NOT_INLINE)) {
                pay = q2;
                nk = ((uint32_t)q3 & 0x7Fu) | (HK_INLINE << 8);
            } else {
                pay = h;
                nk = (short_key ? ((uint32_t)q3 & 0x7Fu) : ((uint32_t)(q3 >> CACHE_LONG_N_SHIFT) & 0xFFu)) | (HK_SLOT << 8);
                asm volatile("prefetch.global.L2 [%0];" ::"l"(&a.cache.slots[h].n));
            }
            return true;
        }
        h = (h + 1) & a.cache.mask;
    }
    return false;
}

// a warp waits for the base of CTA tile (iteration `it`) to be published
__device__ __forceinline__ void hot_wait_flag(const uint32_t *flag, uint32_t want, uint32_t lane) {
    if (lane == 0)
        while (*((volatile const uint32_t *)flag) != want) __nanosleep(100);
    __syncwarp();
    __threadfence_block();
}

// the ids of a finished warp tile leave: n words from the warp's staging buffer to out[base ..), 128 bytes per warp store
__device__ __forceinline__ void hot_store(const EncArgs &a, const uint32_t *stg, uint64_t base, uint32_t n, uint32_t lane) {
    if (base + n <= a.out_cap) {
        for (uint32_t i = lane; i < n; i += 32) __stcs(&a.out[base + i], stg[i]);
    } else {
        for (uint32_t i = lane; i < n; i += 32)
            if (base + i < a.out_cap) a.out[base + i] = stg[i];
        if (lane == 0) *a.overflow = 1;
    }
}

// PIF = chunk-cache probes a lane keeps in flight together (its CPT chunks are probed in groups of PIF)
template <int THREADS, int CPT, int PIF>
__global__ void __launch_bounds__(THREADS, 1) k_encode_hot(const __grid_constant__ EncArgs a) {
    using SM = HotSmemT<THREADS, CPT>;
    constexpr int NW = SM::NW, TW = SM::TW, STG = SM::STG;
    static_assert(NW <= 32 && CPT % 2 == 0 && CPT % PIF == 0, "the last warp scans the NW counts in one go; counts are scanned in packed pairs");
    extern __shared__ __align__(128) unsigned char enc_smem_raw[];
    SM &sm = *reinterpret_cast<SM *>(enc_smem_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t i = tid; i < (uint32_t)HOT_N; i += THREADS) sm.hot[i] = __ldg(&a.hot_img[i]);
    if (tid < 2) {
        sm.cnt[tid] = 0;
        sm.flag[tid] = 0;
    }
    __syncthreads();
    const uint32_t *const words = reinterpret_cast<const uint32_t *>(a.bytes);
    const uint32_t full_bytes = (uint32_t)(a.n_bytes_total & ~3ull); // chunks that end beyond it take the slow path (byte loads)
    const uint32_t wlast = full_bytes ? (full_bytes >> 2) - 1 : 0;   // word loads are clamped to the last whole word
    uint32_t *const cells = a.spill + ((size_t)blockIdx.x * NW + warp) * (size_t)TW * ENC_SHORT_MAX;
    const uint32_t G = gridDim.x;
    uint32_t pend_total = 0, pend_it = 0; // the previous tile of this warp: its ids wait in stage[pend_it & 1]
    bool pending = false;
    // MBPE_DEBUG statistics (EncArgs::prof): 0 chunks answered by the hot table, 1 by the home slot of the chunk cache, 2 by
    // the slow path, 3 by the inline rest of the probe sequence, 4..10 cycles of lane 0 per step (1, 2, 3+4, 5, wait for
    // the previous tile's base, its store, whole tile), 11 warp tiles
    uint32_t st_hot = 0, st_fast = 0, st_slow = 0, st_more = 0;
    unsigned long long st_cyc[7] = {0, 0, 0, 0, 0, 0, 0}, st_tiles = 0;
    long long t_lap = 0, t_tile = 0;
    auto lap = [&](int i) {
        if (HOT_STATS && a.prof) {
            const long long t = clock64();
            st_cyc[i] += (unsigned long long)(t - t_lap);
            t_lap = t;
        }
    };

    for (uint32_t it = 0, T = blockIdx.x;; it++, T += G) {
        const bool have = T < a.n_tiles;
        const uint32_t p = it & 1;
        uint32_t cur_total = 0;
        if (HOT_STATS && a.prof) t_lap = t_tile = clock64();
        if (have) {
            const uint64_t cw = min(a.chunk0 + ((uint64_t)T * NW + warp) * TW, a.chunk1);
            const uint32_t nc = (uint32_t)min((uint64_t)TW, a.chunk1 - cw);
            uint32_t nk[CPT];  // the chunk's length until it is resolved, then count | kind << 8
            uint64_t pay[CPT]; // inline ids / cache slot
            uint64_t key[CPT]; // first 8 bytes
            uint32_t f[CPT], ol[CPT]; // hash of the key, text offset
            uint32_t open = 0;        // bit j: chunk j still needs the chunk cache
            // the NEXT tile of this warp (the schedule is static): its boundaries are asked into L2 now, its text at step 3
            const bool have_next = T + G < a.n_tiles;
            const uint64_t cwn = min(a.chunk0 + ((uint64_t)(T + G) * NW + warp) * TW, a.chunk1);
            uint32_t nb0 = 0, nb1 = 0;
            if (have_next) {
                if (lane < (TW + 32) / 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(&a.off[min(cwn + lane * 32, a.chunk1)]));
                nb0 = __ldg(&a.off[cwn]);
                nb1 = __ldg(&a.off[min(cwn + TW, a.chunk1)]);
            }
            // ---- 1. boundaries, first 8 bytes, hot table -------------------------------------------------------------
#pragma unroll
            for (int j = 0; j < CPT; j++) {
                const uint32_t k = j * 32 + lane;
                const uint32_t o0 = __ldg(&a.off[cw + min(k, nc)]), o1 = __ldg(&a.off[cw + min(k + 1, nc)]);
                ol[j] = o0;
                nk[j] = o1 - o0;
            }
#pragma unroll
            for (int j = 0; j < CPT; j++) {
                const uint32_t o0 = ol[j];
                const uint32_t wi = o0 >> 2, sh = (o0 & 3) * 8;
                const uint32_t t0 = __ldg(&words[min(wi, wlast)]), t1 = __ldg(&words[min(wi + 1, wlast)]), t2 = __ldg(&words[min(wi + 2, wlast)]);
                const uint64_t v = ((uint64_t)__funnelshift_r(t1, t2, sh) << 32) | __funnelshift_r(t0, t1, sh);
                const uint32_t l8 = min(max(nk[j], 1u), 8u);
                key[j] = v & (~0ull >> ((8 - l8) * 8));
                f[j] = cache_hash_fin(key[j]);
            }
#pragma unroll
            for (int j = 0; j < CPT; j++) {
                const uint32_t len = nk[j];
                const uint32_t l8 = min(max(len, 1u), 8u);
                const uint64_t m = ~0ull >> ((8 - l8) * 8);
                const uint4 e = sm.hot[f[j] >> (32 - HOT_LOG2)];
                const uint64_t ek = ((uint64_t)e.y << 32) | e.x, ev = ((uint64_t)e.w << 32) | e.z;
                // (last byte non-zero <=> the key is larger than any key of len - 1 bytes)
                const bool hit = len - 1u < 8u && ol[j] + len <= full_bytes && key[j] > (m >> 8) && ek == key[j];
                pay[j] = 0;
                if (hit) {
                    pay[j] = ev;
                    nk[j] = (1u + ((ev >> CACHE_ID_BITS) != 0) + ((ev >> (2 * CACHE_ID_BITS)) != 0)) | (HK_HOT << 8);
                    if (HOT_STATS) st_hot++;
                } else if (len != 0) {
                    open |= 1u << j;
                }
            }
            lap(0);
            // ---- 2. chunk cache: home slot, keys of up to 16 bytes; everything else is the slow path --------------------
            if (__any_sync(0xffffffffu, open != 0)) {
#pragma unroll
                for (int g = 0; g < CPT; g += PIF) {
                    if (!__any_sync(0xffffffffu, (open >> g) & ((1u << PIF) - 1))) continue;
                    uint64_t q0[PIF], q1[PIF], q2[PIF], q3[PIF], k1[PIF];
                    uint32_t h[PIF];
                    bool fast[PIF];
#pragma unroll
                    for (int q = 0; q < PIF; q++) {
                        const int j = g + q;
                        const uint32_t len = nk[j], o0 = ol[j];
                        fast[q] = ((open >> j) & 1u) && a.cache.slots != nullptr && len <= CACHE_SHORT_KEY && o0 + len <= full_bytes;
                        k1[q] = 0;
                        h[q] = f[j];
                        q0[q] = q1[q] = q2[q] = q3[q] = 0;
                        if (fast[q]) {
                            if (len > 8) {
                                const uint32_t wi = o0 >> 2, sh = (o0 & 3) * 8;
                                const uint32_t t2 = __ldg(&words[min(wi + 2, wlast)]), t3 = __ldg(&words[min(wi + 3, wlast)]),
                                               t4 = __ldg(&words[min(wi + 4, wlast)]);
                                const uint64_t v = ((uint64_t)__funnelshift_r(t3, t4, sh) << 32) | __funnelshift_r(t2, t3, sh);
                                k1[q] = v & (~0ull >> ((16 - len) * 8));
                                h[q] = cache_hash_fin(cache_hash_lo(key[j], k1[q]));
                            }
                            h[q] >>= a.cache.shift;
                            ld_sector256<0>(&a.cache.slots[h[q]], q0[q], q1[q], q2[q], q3[q]);
                        }
                    }
#pragma unroll
                    for (int q = 0; q < PIF; q++) {
                        const int j = g + q;
                        const uint32_t len = nk[j];
                        if ((open >> j) & 1u) {
                            const bool hit = fast[q] && q0[q] == key[j] && q1[q] == k1[q] && (q3[q] >> 56) == len;
                            if (hit) {
                                if (HOT_STATS) st_fast++;
                                const uint32_t n = (uint32_t)q3[q] & 0x7Fu;
                                if (!(q3[q] & CACHE_NOT_INLINE)) {
                                    pay[j] = q2[q];
                                    nk[j] = n | (HK_INLINE << 8);
                                } else {
                                    pay[j] = h[q];
                                    nk[j] = n | (HK_SLOT << 8);
                                    asm volatile("prefetch.global.L2 [%0];" ::"l"(&a.cache.slots[h[q]].n)); // (step 5 reads the value sector)
                                }
                            } else {
                                // the home slot holds another chunk, or the key has 17..30 bytes: the lookup goes on inline
                                const bool in_words = ol[j] + len <= full_bytes && a.cache.slots != nullptr;
                                const bool more = in_words && (fast[q] ? q3[q] != 0 : len <= CACHE_MAX_LEN);
                                const bool cached_nowhere = fast[q] && q3[q] == 0;
                                bool found = false;
                                if (more) found = hot_probe_more(a, words, wlast, ol[j], len, key[j], k1[q], h[q], nk[j], pay[j]);
                                if (found) {
                                    if (HOT_STATS) st_more++;
                                } else {
                                    (void)cached_nowhere;
                                    if (HOT_STATS) st_slow++;
                                    nk[j] = hot_slow_chunk(a, ol[j], len, cells + (size_t)(j * 32 + lane) * ENC_SHORT_MAX, cw + j * 32 + lane, pay[j]);
                                }
                            }
                        }
                    }
                }
            }
            lap(1);
            if (have_next) {
                for (uint32_t g = nb0 + lane * 128; g < nb1 + 32 && g < full_bytes; g += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.bytes + g));
            }
            // ---- 3. every chunk's place in the warp tile (chunk order = j major): packed inclusive scans -----------------
            uint32_t loc[CPT], total = 0;
#pragma unroll
            for (int j = 0; j < CPT; j += 2) {
                const uint32_t n0 = nk[j] & 0xFFu, n1 = nk[j + 1] & 0xFFu;
                uint32_t incl = n0 | (n1 << 16);
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= (uint32_t)d) incl += v;
                }
                const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
                loc[j] = total + (incl & 0xFFFFu) - n0;
                total += tot & 0xFFFFu;
                loc[j + 1] = total + (incl >> 16) - n1;
                total += tot >> 16;
            }
            // ---- 4. the CTA tile: the last warp to arrive places all of its warp tiles --------------------------------------
            uint32_t last = 0;
            if (lane == 0) {
                ((volatile uint32_t *)sm.wsum[p])[warp] = total;
                __threadfence_block();
                last = atomicAdd(&sm.cnt[p], 1u) == (uint32_t)NW - 1;
            }
            last = __shfl_sync(0xffffffffu, last, 0);
            if (last) {
                __threadfence_block();
                const uint32_t v = lane < (uint32_t)NW ? ((volatile uint32_t *)sm.wsum[p])[lane] : 0u;
                uint32_t incl = v;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t u = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= (uint32_t)d) incl += u;
                }
                const uint32_t cta_total = __shfl_sync(0xffffffffu, incl, 31);
                if (lane == 0) sm.cnt[p] = 0;
                const uint64_t base0 = lookback_announce(a.status, T, cta_total, a.stream_base);
                const uint64_t b = lookback_resolve<4>(a.status, T, cta_total, base0);
                if (lane < (uint32_t)NW) ((volatile unsigned long long *)sm.wbase[p])[lane] = b + (incl - v);
                __threadfence_block();
                __syncwarp();
                if (lane == 0) {
                    ((volatile uint32_t *)sm.flag)[p] = it + 1;
                    if (T == a.n_tiles - 1) *a.d_n_out = b + cta_total;
                }
            }
            lap(2);
            // ---- 5. gather the ids: staging buffer, or (a tile with more ids than it holds) straight into the stream -------
            if (total <= (uint32_t)STG) {
                uint32_t *const stg = sm.stage[p][warp];
                hot_emit_all<CPT, PIF>(a, nk, pay, loc, stg, cells, lane);
                cur_total = total;
            } else {
                hot_wait_flag(&sm.flag[p], it + 1, lane);
                const uint64_t base = ((volatile unsigned long long *)sm.wbase[p])[warp];
                if (base + total <= a.out_cap) {
                    hot_emit_all<CPT, PIF>(a, nk, pay, loc, a.out + base, cells, lane);
                } else if (lane == 0) {
                    *a.overflow = 1;
                }
            }
            __syncwarp();
            lap(3);
        }
        // ---- 6. the previous tile leaves now: its base has long been published ---------------------------------------------
        if (pending && pend_total) {
            const uint32_t pp = pend_it & 1;
            hot_wait_flag(&sm.flag[pp], pend_it + 1, lane);
            lap(4);
            const uint64_t base = ((volatile unsigned long long *)sm.wbase[pp])[warp];
            hot_store(a, sm.stage[pp][warp], base, pend_total, lane);
            __syncwarp();
        }
        lap(5);
        if (HOT_STATS && a.prof) {
            st_cyc[6] += (unsigned long long)(clock64() - t_tile);
            st_tiles += have;
        }
        if (!have) break;
        pending = true;
        pend_total = cur_total;
        pend_it = it;
    }
    if (HOT_STATS && a.prof) {
        for (int d = 16; d > 0; d >>= 1) {
            st_hot += __shfl_xor_sync(0xffffffffu, st_hot, d);
            st_fast += __shfl_xor_sync(0xffffffffu, st_fast, d);
            st_slow += __shfl_xor_sync(0xffffffffu, st_slow, d);
            st_more += __shfl_xor_sync(0xffffffffu, st_more, d);
        }
        if (lane == 0) {
            atomicAdd(&a.prof[0], (unsigned long long)st_hot);
            atomicAdd(&a.prof[1], (unsigned long long)st_fast);
            atomicAdd(&a.prof[2], (unsigned long long)st_slow);
            atomicAdd(&a.prof[3], (unsigned long long)st_more);
            for (int i = 0; i < 7; i++) atomicAdd(&a.prof[4 + i], st_cyc[i]);
            atomicAdd(&a.prof[11], st_tiles);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// The image of the hot table, rebuilt from a strided sample of the chunks about to be encoded (four small launches):
// count every eligible sample chunk under 20 bits of its hash, elect per table slot the most frequent one, write its key,
// and fetch its ids from the chunk cache (an entry whose chunk the cache does not hold inline stays empty).
// ---------------------------------------------------------------------------------------------------------
struct HotBuild {
    const uint8_t *bytes;
    const uint32_t *off;
    uint64_t chunk0, stride;
    uint32_t n_sample;
    uint32_t *count;              // 1 << HOT_COUNT_LOG2, zeroed
    unsigned long long *best;     // HOT_N, zeroed
    uint4 *img;                   // HOT_N, zeroed
};
// key of sample chunk i if it is eligible (1..8 bytes, last byte non-zero)
__device__ __forceinline__ bool hot_sample_key(const HotBuild &b, uint32_t i, uint64_t &key) {
    const uint64_t c = b.chunk0 + (uint64_t)i * b.stride;
    const uint32_t o = __ldg(&b.off[c]), len = __ldg(&b.off[c + 1]) - o;
    if (len - 1u >= 8u) return false;
    key = 0;
    for (uint32_t q = 0; q < len; q++) key |= (uint64_t)__ldg(&b.bytes[o + q]) << (8 * q);
    return (key >> (8 * (len - 1))) != 0;
}
__global__ void k_hot_count(HotBuild b) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t key;
    if (i < b.n_sample && hot_sample_key(b, i, key)) atomicAdd(&b.count[cache_hash_fin(key) >> (32 - HOT_COUNT_LOG2)], 1u);
}
__global__ void k_hot_elect(HotBuild b) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t key;
    if (i < b.n_sample && hot_sample_key(b, i, key)) {
        const uint32_t f = cache_hash_fin(key);
        atomicMax(&b.best[f >> (32 - HOT_LOG2)], ((unsigned long long)b.count[f >> (32 - HOT_COUNT_LOG2)] << 32) | f);
    }
}
__global__ void k_hot_fill(HotBuild b) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t key;
    if (i < b.n_sample && hot_sample_key(b, i, key)) {
        const uint32_t f = cache_hash_fin(key);
        if (b.best[f >> (32 - HOT_LOG2)] == (((unsigned long long)b.count[f >> (32 - HOT_COUNT_LOG2)] << 32) | f))
            *reinterpret_cast<unsigned long long *>(&b.img[f >> (32 - HOT_LOG2)]) = key; // (chunks that share all 32 hash bits: any of them)
    }
}
__global__ void k_hot_resolve(HotBuild b, ChunkCache cc) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= (uint32_t)HOT_N) return;
    const uint4 e = b.img[s];
    const uint64_t key = ((uint64_t)e.y << 32) | e.x;
    uint4 out = make_uint4(0, 0, 0, 0);
    if (key != 0 && cc.slots) {
        const uint32_t len = 8 - (uint32_t)__clzll((long long)key) / 8;
        uint32_t h = cache_hash_fin(key) >> cc.shift;
        for (uint32_t probes = 0; probes < CACHE_MAX_PROBES; probes++) {
            const CacheSlot &sl = cc.slots[h];
            const uint64_t q3 = sl.k[3];
            if (q3 == 0) break;
            if (sl.k[0] == key && sl.k[1] == 0 && (q3 >> 56) == len) {
                const uint32_t n = (uint32_t)q3 & 0x7Fu;
                if (!(q3 & CACHE_NOT_INLINE) && n >= 1 && n <= CACHE_KEY_IDS) {
                    uint64_t ids = 0;
                    bool ok = true;
                    for (uint32_t q = 0; q < n; q++) {
                        const uint64_t id = (sl.k[2] >> (CACHE_ID_BITS * q)) & HOT_ID_MASK;
                        ok = ok && id + 1 <= HOT_ID_MASK;
                        ids |= (id + 1) << (CACHE_ID_BITS * q);
                    }
                    if (ok) out = make_uint4(e.x, e.y, (uint32_t)ids, (uint32_t)(ids >> 32));
                }
                break;
            }
            h = (h + 1) & cc.mask;
        }
    }
    b.img[s] = out;
}

} // namespace mbpe
