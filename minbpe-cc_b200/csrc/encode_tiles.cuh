// encode_tiles.cuh -- k_encode_tiles: the encode merge scan over tiles of chunks (included by encode.cu only).
//
// Replaces internal_internal_encode / internal_encode + flatten (Tokenizer.h:325-377, :714-717) for chunks of up to
// ENC_SHORT_MAX bytes; longer chunks are encoded by k_encode_long into a scratch stream and spliced in here.
//
// The flat id stream needs every tile's place = the number of ids before it. Round 1 (and the first versions of round 2)
// resolved that inside ONE pass by decoupled look-back; measured per tile (thread 0's clock, 1024 chunks): ~40 k cycles,
// of which ~13 k waiting in the look-back for the predecessors' counts (their arrival times jitter by the whole chain of
// ticket -> boundaries -> text -> probes) and ~7 k in that fetch chain, which cannot be prefetched because tiles must
// START in ticket order. So the stream is produced in TWO passes that have no dependency between tiles at all:
//   pass 1 (count)  every tile: probes, open chunks -> the tile's id count                      -> tile_total[t]
//   k_tile_scan     exclusive scan of tile_total (+ the ids of earlier launches)                -> tile_base[t]
//   pass 2 (write)  every tile: probes again (the slots are hot in L2), ids gathered in shared memory, stored at
//                   tile_base[t] as whole lines.
// With nothing to wait for, tiles are dealt statically (tile = CTA index + k x grid) and the NEXT tile's boundaries and
// text arrive by bulk asynchronous copy (cp.async.bulk -> the TMA engine, completion on mbarriers, double buffered) while
// the current one is processed. The text and the boundaries are read twice (2 x 1.8 B per text byte against 4.4 B of
// ids: +25 % traffic); chunks pass 1 had to scan are in the caches before pass 2 (k_cache_insert runs in between).
//
// One tile = THREADS x CPT consecutive chunks, one CTA:
//   1. fast path, every thread CPT chunks: chunks of <= 15 bytes (nine in ten) are looked up in the SMALL chunk cache,
//      "chunk bytes -> ids" in 32-byte slots = one DRAM sector, fetched by ONE 256-bit load (LDG.E.256) per probe, all
//      probes of a thread in flight together.
//   2. what that leaves open: chunks that may be cached elsewhere (16..31 bytes or more than 4 ids: BIG cache, 64-byte
//      slots; a taken home slot: the probe sequence goes on) are resolved by their own thread, no barrier; chunks nobody
//      has seen go to the tile's scan list: the multi-pass scan itself, one warp per chunk when few, one thread per chunk
//      when many; pass 1 appends them to a log that k_cache_insert folds into the caches.
//   3. pass 1: block sum. pass 2: block scan, gather, 16-byte stores.
// Results never depend on the caches: a miss is scanned, and special tokens are matched in the open-chunk path itself.
#pragma once
#include "lookback.cuh"

namespace mbpe {

// ---------------------------------------------------------------------------------------------------------
// mbarrier + bulk asynchronous copy (global -> shared through the TMA engine; SASS: UBLKCP + SYNCS)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MBPE_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MBPE_DONE_%=;\n"
        "bra MBPE_WAIT_%=;\n"
        "MBPE_DONE_%=:\n"
        "}\n" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}
// the one-pass streams (text, boundaries) should not push the randomly probed cache tables out of L2
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// bytes: multiple of 16; dst and src 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar)), "l"(policy)
                 : "memory");
}

// one 32-byte slot of the SMALL cache in ONE load (LDG.E.256, new with sm_100): one L1 wavefront per probe instead of two
__device__ __forceinline__ void ld_slot256(const void *slot, uint4 &k, uint4 &v) {
    unsigned long long q0, q1, q2, q3;
    asm("ld.global.nc.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(q0), "=l"(q1), "=l"(q2), "=l"(q3) : "l"(slot));
    k = make_uint4((uint32_t)q0, (uint32_t)(q0 >> 32), (uint32_t)q1, (uint32_t)(q1 >> 32));
    v = make_uint4((uint32_t)q2, (uint32_t)(q2 >> 32), (uint32_t)q3, (uint32_t)(q3 >> 32));
}

// ---------------------------------------------------------------------------------------------------------
// Chunk caches. The multi-pass scan of a chunk is a pure function of its bytes, and text repeats its chunks (Zipf):
// the encoder keeps "chunk bytes -> ids" in two open-addressed tables in HBM.
//   SMALL: chunks of <= 15 bytes with <= 4 ids, 32-byte slots {key 16 B, ids 16 B} = one sector per probe.
//   BIG:   chunks of <= 31 bytes (any id count, > 7 ids in an arena), 64-byte slots.
// The tile kernel only READS them; chunks it had to scan go to a log, and k_cache_insert adds the log to the tables
// between launches -- no kernel both reads and writes a table, so there is no publication protocol to get wrong.
// Results are bit-identical with or without the caches (MBPE_ENCODE_CACHE=0 disables them; tests run both).
// ---------------------------------------------------------------------------------------------------------
struct SmallSlot { // 32 bytes
    uint32_t k[4]; // bytes 0..14 little endian, zero padded; k[3] bits 24..27 = length (1..15), bits 28..30 = id count (7: stub);
                   // k[3] == 0: empty
    uint32_t v[4]; // the ids
};
static_assert(sizeof(SmallSlot) == 32, "one sector per entry");
constexpr uint32_t SMALL_MAX_LEN = 15, SMALL_MAX_IDS = 4, SMALL_KEY_MASK = 0x0FFFFFFFu;
constexpr uint32_t SMALL_STUB = 7; // id count of a stub: the chunk has more than 4 ids, they are in the BIG cache

struct CacheSlot { // 64 bytes = two sectors: key, value
    uint64_t k[4]; // chunk bytes, little endian, zero padded; top byte of k[3] = length (1..31); k[3] == 0: empty
    uint32_t n;    // number of ids (1..31)
    uint32_t v[7]; // n <= 7: the ids; otherwise v[0] = offset of the ids in the arena
};
static_assert(sizeof(CacheSlot) == 64, "two sectors per entry");
struct CacheLogEntry {
    uint64_t k[4]; // BIG-table key layout
    uint32_t n;
    uint32_t ids[31];
};
constexpr uint32_t CACHE_MAX_LEN = 31, CACHE_INLINE_IDS = 7;
// An entry lives at most this many slots from its home: inserts give up beyond it (the chunk is simply not cached), so
// lookups may stop there too -- no probe loop depends on the table having a free slot.
constexpr uint32_t CACHE_MAX_PROBES = 64;

struct ChunkCache {
    SmallSlot *small; // nullptr = caches disabled
    uint32_t small_shift; // slot = hash >> small_shift
    uint32_t small_mask;
    CacheSlot *slots;
    uint32_t mask;       // slots - 1
    CacheLogEntry *log;  // chunks the current launch had to scan
    uint32_t *log_count;
    uint32_t log_cap;
    uint32_t *used;      // [0] occupied BIG slots, [1] occupied SMALL slots (learning stops at half full)
    uint32_t *arena;     // ids of BIG entries with more than CACHE_INLINE_IDS ids
    uint32_t *arena_used;
    uint32_t arena_cap;
};

__host__ __device__ __forceinline__ uint32_t small_hash(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
    const uint64_t lo = ((uint64_t)w1 << 32) | w0, hi = ((uint64_t)w3 << 32) | w2;
    const uint64_t m = hi * 0x9E3779B97F4A7C15ull;
    const uint64_t x = lo ^ ((m >> 32) | (m << 32));
    return (uint32_t)((x * 0xD6E8FEB86659FD93ull) >> 32); // callers use the TOP bits: they depend on every key bit
}
__host__ __device__ __forceinline__ uint32_t cache_hash(uint64_t k0, uint64_t k1, uint64_t k2, uint64_t k3) {
    uint64_t h = (k0 ^ (k1 * 0x9E3779B97F4A7C15ull)) * 0xff51afd7ed558ccdULL;
    h ^= (k2 * 0xc2b2ae3d27d4eb4fULL) ^ (k3 * 0x165667b19e3779f9ULL);
    h ^= h >> 32;
    h *= 0xc4ceb9fe1a85ec53ULL;
    return (uint32_t)(h >> 32);
}

__global__ void k_cache_insert(ChunkCache cc) {
    const uint32_t n = min(*cc.log_count, cc.log_cap);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const CacheLogEntry &e = cc.log[i];
        const uint64_t k0 = e.k[0], k1 = e.k[1], k2 = e.k[2], k3 = e.k[3];
        const uint32_t en = e.n, len = (uint32_t)(k3 >> 56);
        // SMALL entry: the ids themselves (<= 4), or -- for a short chunk with more ids -- a stub (id count 7) that says
        // "look in BIG": the tile kernel's fast path then knows the chunk is cached without probing BIG for every miss
        if (len <= SMALL_MAX_LEN) {
            const bool stub = en > SMALL_MAX_IDS;
            if (*((volatile uint32_t *)(cc.used + 1)) * 2 <= cc.small_mask) { // (half full: stop learning)
                const uint32_t w0 = (uint32_t)k0, w1 = (uint32_t)(k0 >> 32), w2 = (uint32_t)k1;
                const uint32_t w3 = (uint32_t)(k1 >> 32) | (len << 24); // byte 15 is free: len <= 15
                uint32_t h = small_hash(w0, w1, w2, w3) >> cc.small_shift;
                for (uint32_t probes = 0; probes < CACHE_MAX_PROBES; probes++) {
                    SmallSlot *s = &cc.small[h];
                    uint32_t cur = *((volatile uint32_t *)&s->k[3]);
                    if (cur == 0) {
                        cur = atomicCAS(&s->k[3], 0u, w3 | ((stub ? SMALL_STUB : en) << 28));
                        if (cur == 0) { // claimed: k[3] is the claim word, the rest is written by the winner only
                            s->k[0] = w0;
                            s->k[1] = w1;
                            s->k[2] = w2;
                            for (uint32_t q = 0; q < SMALL_MAX_IDS; q++) s->v[q] = (!stub && q < en) ? e.ids[q] : 0u;
                            atomicAdd(cc.used + 1, 1u);
                            break;
                        }
                    }
                    // Same key: already there (the log holds duplicates of hot chunks). Words of a slot claimed in THIS launch
                    // may not be visible yet; then the duplicate takes a second slot -- harmless, both slots hold the same ids
                    // and a reader uses the first one it finds.
                    if ((cur & SMALL_KEY_MASK) == w3 && *((volatile uint32_t *)&s->k[0]) == w0 &&
                        *((volatile uint32_t *)&s->k[1]) == w1 && *((volatile uint32_t *)&s->k[2]) == w2)
                        break;
                    h = (h + 1) & cc.small_mask;
                }
            }
            if (!stub) continue;
        }
        if (*((volatile uint32_t *)cc.used) * 2 > cc.mask) continue; // half full: stop learning
        uint32_t h = cache_hash(k0, k1, k2, k3) & cc.mask;
        for (uint32_t probes = 0; probes < CACHE_MAX_PROBES; probes++) {
            unsigned long long *claim = reinterpret_cast<unsigned long long *>(&cc.slots[h].k[3]);
            unsigned long long cur = *((volatile unsigned long long *)claim);
            if (cur == 0) {
                uint32_t aoff = 0;
                if (en > CACHE_INLINE_IDS) { // reserve arena room BEFORE claiming, so a claimed slot is always completed
                    aoff = atomicAdd(cc.arena_used, en);
                    if (aoff + en > cc.arena_cap) break;
                }
                cur = atomicCAS(claim, 0ull, (unsigned long long)k3);
                if (cur == 0) {
                    cc.slots[h].k[0] = k0;
                    cc.slots[h].k[1] = k1;
                    cc.slots[h].k[2] = k2;
                    cc.slots[h].n = en;
                    if (en <= CACHE_INLINE_IDS) {
                        for (uint32_t q = 0; q < en; q++) cc.slots[h].v[q] = e.ids[q];
                    } else {
                        cc.slots[h].v[0] = aoff;
                        for (uint32_t q = 0; q < en; q++) cc.arena[aoff + q] = e.ids[q];
                    }
                    atomicAdd(cc.used, 1u);
                    break;
                }
            }
            if (cur == k3 && *((volatile uint64_t *)&cc.slots[h].k[0]) == k0 &&
                *((volatile uint64_t *)&cc.slots[h].k[1]) == k1 && *((volatile uint64_t *)&cc.slots[h].k[2]) == k2)
                break;
            h = (h + 1) & cc.mask;
        }
    }
}
__global__ void k_cache_reset_log(ChunkCache cc) { *cc.log_count = 0; }

// ---------------------------------------------------------------------------------------------------------
// special tokens on the encode side (Tokenizer.h:667-671): a chunk with exactly these bytes is this one id
// ---------------------------------------------------------------------------------------------------------
struct EncSpecials {
    const uint32_t *ids;
    const uint32_t *off; // n + 1 offsets into bytes
    const uint8_t *bytes;
    uint32_t n;
    unsigned long long len_mask; // bit min(length, 63) set for every token length present
};
template <class ByteAt>
__device__ __forceinline__ uint32_t special_match(const EncSpecials &sp, uint32_t len, ByteAt at) {
    if (sp.n == 0 || !((sp.len_mask >> (len < 63u ? len : 63u)) & 1ull)) return ENC_NONE;
    for (uint32_t s = 0; s < sp.n; s++) {
        const uint32_t b = __ldg(&sp.off[s]);
        if (__ldg(&sp.off[s + 1]) - b != len) continue;
        uint32_t i = 0;
        while (i < len && __ldg(&sp.bytes[b + i]) == at(i)) i++;
        if (i == len) return __ldg(&sp.ids[s]);
    }
    return ENC_NONE;
}

// ---------------------------------------------------------------------------------------------------------
// k_encode_tiles
// ---------------------------------------------------------------------------------------------------------
struct EncArgs {
    EncTable tab;
    ChunkCache cache;
    EncSpecials sp;
    uint64_t chunk0, chunk1;   // this launch covers chunks [chunk0, chunk1) in tiles
    const unsigned long long *stream_base; // ids produced by earlier launches of this call (device word) = base of tile 0
    const uint8_t *bytes;
    uint64_t n_bytes_total; // bytes readable from `bytes`
    const uint32_t *off;
    uint64_t n_chunks;
    uint32_t *out;
    uint64_t out_cap;
    unsigned long long *d_n_out;
    unsigned long long *out_off; // optional per-chunk token offsets (n_chunks + 1)
    uint32_t *tile_total;        // pass 1 out: ids of every tile of this launch
    const unsigned long long *tile_base; // pass 2 in: place of every tile in the stream (k_tile_scan)
    uint32_t n_tiles;
    const uint32_t *scratch_a;   // long-chunk tokens / counts; null = no long pre-pass was run (optimistic launch)
    const uint32_t *scratch_b;
    uint32_t *long_list;         // optimistic launch: chunks longer than ENC_SHORT_MAX are reported here
    uint32_t *n_long;
    uint32_t long_cap;
    uint32_t *overflow;          // set when out_cap is too small
    uint32_t *miss_count;        // chunks that went through the scan (statistics)
    uint32_t *spill;             // parking overflow: gridDim.x blocks of EncSmemT::SPILL words
    uint32_t bulk;               // bytes / off are 16-byte aligned: stage with cp.async.bulk
    uint32_t out_aligned;        // out is 16-byte aligned: ids leave as 16-byte stores
    unsigned long long *prof;    // optional: SM cycles per phase summed over CTAs (thread 0's clock), see ENC_PROF_*
    uint32_t ablate;             // MBPE_ENC_ABLATE (profiling only, WRONG results): 2 no cache probe (every short chunk
                                 // "hits" with two fake ids), 4 no id stores, 8 open chunks are not resolved (one fake id)
};

// EncArgs::prof[pass * 8 + i]: cycles thread 0 spent ... 0 waiting for the tile's data, 1 cache probes + cached open chunks
// (+ barrier), 2 scan list, 3 block sum / scan (+ barrier), 4 gather (+ barrier), 5 issuing the prefetch, 6 storing the
// ids, 7 tiles processed
constexpr int ENC_PROF_N = 16;
constexpr uint32_t TILE_NONE = 0xFFFFFFFFu;
constexpr uint64_t ENC_MAX_SUBBATCH = 1ull << 26; // chunks per launch at most (tile ids and look-back words are 32-bit safe)
// Parking area of a tile = ids of its open chunks until they are written: the first PARK words live in shared memory, the
// rest in the CTA's spill block in HBM (one index space; every id covers at least one byte of text, so a tile of
// ENC_SHORT_MAX-byte chunks at most needs TILE * ENC_SHORT_MAX words: it always fits, nothing is ever encoded twice).
constexpr uint32_t ET_WARP_SCAN_MAX = 48;     // up to this many scans per tile run one warp per chunk

template <int THREADS, int CPT>
struct EncSmemT {
    static constexpr int TILE = THREADS * CPT;
    static constexpr int TEXT_CAP = TILE * 10;  // staged text bytes per tile (average chunk ~5 bytes); wider tiles read HBM
    static constexpr int STAGE = TILE * 5 / 2;  // ids gathered per tile (average ~2.1 per chunk); more: direct stores
    static constexpr int PARK = TILE;           // parking words in shared memory (warm caches park a few dozen ids per tile)
    static constexpr int SPILL = TILE * 64;     // parking words per CTA in HBM (EncArgs::spill): the worst case
    alignas(128) uint32_t off_buf[2][TILE + 8];          // double buffered: the next tile's data arrives during this tile
    alignas(128) uint32_t text_buf[2][TEXT_CAP / 4 + 16]; // + halo: key assembly reads whole words past the chunk's end
    alignas(16) uint32_t stage[STAGE + 4];
    uint32_t park[PARK];
    uint32_t meta[TILE];       // per open chunk (index = its place on the open list): start in park (20 bits) | id count << 20
    uint16_t open_k[TILE];     // open list (chunks nobody has seen before): chunk index within the tile
    uint32_t warp_scratch[THREADS / 32][32];
    uint32_t warp_sum[THREADS / 32];
    alignas(8) uint64_t bar_off[2], bar_txt[2]; // mbarriers per buffer: boundaries landed (thread 0 waits), text landed (all wait)
    const uint32_t *off;       // the current tile's buffers (set by thread 0 before the tile's first barrier; helpers that run
    const uint32_t *text;      // after a barrier read them)
    uint32_t a0[2], staged[2], n_open, park_used;
    unsigned long long prof[8];
};

template <class SM>
__device__ __forceinline__ uint8_t tile_byte(const EncArgs &a, const SM &sm, bool staged, uint32_t a0, uint32_t g) {
    return staged ? reinterpret_cast<const uint8_t *>(sm.text)[g - a0] : __ldg(&a.bytes[g]);
}

// One warp scans one chunk of <= 32 bytes: lane i holds token i. Per pass every lane looks its pair up at once (one
// lookup latency per pass instead of one per position); the left-to-right non-overlapping rule of
// Tokenizer.h:336-359 is applied to the ballot mask: inside each maximal run of mergeable positions the 1st, 3rd,
// 5th... merge. Runs are separated by parity of their start bit with the carry trick
//   runs_even = F & ~(F + even_starts),   runs_odd = F & ~(F + odd_starts)
// (bit 31 of F is always clear: lane 31 has no right neighbour, so the additions cannot overflow).
// Returns the final length; on return lane i < length holds id i in `tok`. `scratch` = 32 words of shared memory
// owned by the warp.
__device__ __forceinline__ uint32_t scan_chunk_warp(const EncTable &tab, uint32_t &tok, uint32_t len, uint32_t *scratch) {
    const uint32_t lane = threadIdx.x & 31;
    while (len >= 2) {
        const uint32_t nxt = __shfl_down_sync(0xffffffffu, tok, 1);
        uint32_t id = ENC_NONE;
        if (lane + 1 < len) id = enc_lookup_id(tab, tok, nxt);
        const uint32_t F = __ballot_sync(0xffffffffu, id != ENC_NONE);
        if (F == 0) break;
        const uint32_t starts = F & ~(F << 1);
        const uint32_t runs_even = F & ~(F + (starts & 0x55555555u));
        const uint32_t runs_odd = F & ~(F + (starts & 0xAAAAAAAAu));
        const uint32_t M = (runs_even & 0x55555555u) | (runs_odd & 0xAAAAAAAAu); // heads of merged pairs
        const uint32_t valid = len >= 32 ? 0xffffffffu : ((1u << len) - 1);
        const uint32_t keep = valid & ~(M << 1);                                   // tails disappear
        if ((keep >> lane) & 1u) scratch[__popc(keep & ((1u << lane) - 1))] = ((M >> lane) & 1u) ? id : tok;
        __syncwarp();
        len = __popc(keep);
        tok = lane < len ? scratch[lane] : 0u;
        __syncwarp();
    }
    return len;
}

// BIG-table key of chunk [o, o + len), len <= 31: bytes little endian, zero padded, length in the top byte
template <class SM>
__device__ __forceinline__ void big_key(const EncArgs &a, const SM &sm, bool staged, uint32_t a0, uint32_t o, uint32_t len,
                                        uint64_t *key) {
    if (staged) { // whole words of the staged text (the halo makes the over-read safe), bytes past the chunk cleared
        const uint32_t r = o - a0, wi = r >> 2, sh = (r & 3) * 8;
        uint32_t w[9];
#pragma unroll
        for (int q = 0; q < 9; q++) w[q] = sm.text[wi + q];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t lo = __funnelshift_r(w[2 * q], w[2 * q + 1], sh), hi = __funnelshift_r(w[2 * q + 1], w[2 * q + 2], sh);
            const int nb = (int)len - 8 * q; // bytes of this word that belong to the chunk
            const uint64_t v = ((uint64_t)hi << 32) | lo;
            key[q] = nb >= 8 ? v : nb <= 0 ? 0ull : (v & ((1ull << (8 * nb)) - 1));
        }
    } else {
        key[0] = key[1] = key[2] = key[3] = 0;
        for (uint32_t i = 0; i < len; i++) key[i >> 3] |= (uint64_t)tile_byte(a, sm, staged, a0, o + i) << ((i & 7) * 8);
    }
    key[3] |= (uint64_t)len << 56;
}

// a chunk the scan had to encode goes to the log that k_cache_insert folds into the caches after the launch
template <class SM>
__device__ __forceinline__ void log_scanned(const EncArgs &a, const SM &sm, bool staged, uint32_t a0, uint32_t o,
                                            uint32_t len, const uint32_t *ids, uint32_t n) {
    if (!a.cache.small || len > CACHE_MAX_LEN) return;
    const uint32_t li = atomicAdd(a.cache.log_count, 1u);
    if (li >= a.cache.log_cap) return;
    CacheLogEntry &e = a.cache.log[li];
    uint64_t key[4];
    big_key(a, sm, staged, a0, o, len, key);
    e.k[0] = key[0];
    e.k[1] = key[1];
    e.k[2] = key[2];
    e.k[3] = key[3];
    e.n = n;
    for (uint32_t i = 0; i < n; i++) e.ids[i] = ids[i];
}

template <class SM>
__device__ __forceinline__ void park_store(const EncArgs &a, SM &sm, uint32_t at, uint32_t v) {
    if (at < (uint32_t)SM::PARK)
        sm.park[at] = v;
    else
        __stcg(a.spill + (size_t)blockIdx.x * SM::SPILL + (at - SM::PARK), v);
}
template <class SM>
__device__ __forceinline__ uint32_t park_load(const EncArgs &a, const SM &sm, uint32_t at) {
    return at < (uint32_t)SM::PARK ? sm.park[at] : __ldcg(a.spill + (size_t)blockIdx.x * SM::SPILL + (at - SM::PARK));
}
// result of an open chunk: its ids go to the parking area, meta = start | n << 20
template <class SM>
__device__ __forceinline__ uint32_t park_ids(const EncArgs &a, SM &sm, const uint32_t *ids, uint32_t n) {
    const uint32_t at = atomicAdd(&sm.park_used, n);
    for (uint32_t i = 0; i < n; i++) park_store(a, sm, at + i, ids[i]);
    return at | (n << 20);
}

// the multi-pass scan of one chunk (<= ENC_SHORT_MAX bytes) by ONE thread (special tokens first); returns meta
template <class SM>
__device__ __forceinline__ uint32_t scan_serial(const EncArgs &a, SM &sm, uint32_t a0, bool staged, uint32_t o, uint32_t len, bool log) {
    uint32_t t[ENC_SHORT_MAX];
    const uint32_t sid = special_match(a.sp, len, [&](uint32_t i) { return tile_byte(a, sm, staged, a0, o + i); });
    if (sid != ENC_NONE) {
        t[0] = sid;
        return park_ids(a, sm, t, 1);
    }
    for (uint32_t i = 0; i < len; i++) t[i] = tile_byte(a, sm, staged, a0, o + i);
    uint32_t n = len;
    bool merged = true;
    while (merged && n >= 2) n = enc_pass(a.tab, t, n, merged);
    if (log) log_scanned(a, sm, staged, a0, o, len, t, n);
    return park_ids(a, sm, t, n);
}

// one warp encodes the open chunk at open-list place q (<= 32 bytes) and parks its ids
template <class SM>
__device__ __forceinline__ void scan_by_warp(const EncArgs &a, SM &sm, uint32_t a0, bool staged, uint32_t q, uint32_t *scratch, bool log) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t k = sm.open_k[q], so = sm.off[k], mlen = sm.off[k + 1] - so;
    uint32_t sid = ENC_NONE;
    if (a.sp.n) sid = special_match(a.sp, mlen, [&](uint32_t i) { return tile_byte(a, sm, staged, a0, so + i); });
    uint32_t tok = lane < mlen ? tile_byte(a, sm, staged, a0, so + lane) : 0u;
    uint32_t mn = 1;
    if (sid != ENC_NONE)
        tok = sid; // (every lane found the same token)
    else
        mn = scan_chunk_warp(a.tab, tok, mlen, scratch); // lane i < mn: id i in tok
    uint32_t start = 0;
    if (lane == 0) start = atomicAdd(&sm.park_used, mn);
    start = __shfl_sync(0xffffffffu, start, 0);
    if (lane < mn) park_store(a, sm, start + lane, tok);
    if (lane == 0) sm.meta[q] = start | (mn << 20);
    if (log && a.cache.small && mlen <= CACHE_MAX_LEN && sid == ENC_NONE) { // teach the caches
        uint32_t li = 0;
        if (lane == 0) li = atomicAdd(a.cache.log_count, 1u);
        li = __shfl_sync(0xffffffffu, li, 0);
        if (li < a.cache.log_cap) {
            CacheLogEntry &e = a.cache.log[li];
            if (lane < mn) e.ids[lane] = tok;
            if (lane == 0) {
                uint64_t key[4];
                big_key(a, sm, staged, a0, so, mlen, key);
                e.k[0] = key[0];
                e.k[1] = key[1];
                e.k[2] = key[2];
                e.k[3] = key[3];
                e.n = mn;
            }
        }
    }
    __syncwarp();
}

// A chunk the fast path left open but that may well be cached -- 16..31 bytes (BIG cache), a short chunk whose home slot
// in the SMALL cache holds somebody else (its probe sequence goes on) or a stub (more than 4 ids: BIG), a special token.
// By the chunk's own thread, no barrier: the lanes that need it diverge for ONE round trip in the common cases (the
// caller says what the home slot showed, so a stub or a long chunk goes straight to BIG, whose key and value sectors are
// fetched together). Out of line so that its registers are not charged to the fast path.
// hint: 0 = whole SMALL probe sequence (home slot not seen), 1 = home slot taken by another chunk (continue behind it),
//       2 = stub / long chunk: BIG only.
// Returns the id count (0: in no cache -- the scan has to encode it); <= 4 ids come back in v, more are parked (v.x = start).
template <class SM>
__device__ __noinline__ uint32_t resolve_cached(const EncArgs &a, SM &sm, uint32_t a0, bool staged, uint32_t o, uint32_t len, uint32_t hint,
                                                uint4 &v) {
    if (a.sp.n) {
        const uint32_t sid = special_match(a.sp, len, [&](uint32_t i) { return tile_byte(a, sm, staged, a0, o + i); });
        if (sid != ENC_NONE) {
            v.x = sid;
            return 1;
        }
    }
    if (!a.cache.small || len > CACHE_MAX_LEN) return 0;
    uint64_t key[4];
    big_key(a, sm, staged, a0, o, len, key);
    if (len <= SMALL_MAX_LEN && hint != 2) {
        const uint32_t w0 = (uint32_t)key[0], w1 = (uint32_t)(key[0] >> 32), w2 = (uint32_t)key[1];
        const uint32_t w3 = (uint32_t)(key[1] >> 32) | (len << 24);
        uint32_t h = ((small_hash(w0, w1, w2, w3) >> a.cache.small_shift) + (hint == 1 ? 1u : 0u)) & a.cache.small_mask;
        bool stub = false;
        for (uint32_t probes = hint == 1 ? 1u : 0u; probes < CACHE_MAX_PROBES; probes++) {
            uint4 kq, vv;
            ld_slot256(&a.cache.small[h], kq, vv);
            if (kq.w == 0) return 0; // not cached at all
            if ((kq.w & SMALL_KEY_MASK) == w3 && kq.x == w0 && kq.y == w1 && kq.z == w2) {
                if ((kq.w >> 28) <= SMALL_MAX_IDS) {
                    v = vv;
                    return kq.w >> 28;
                }
                stub = true;
                break;
            }
            h = (h + 1) & a.cache.small_mask;
        }
        if (!stub) return 0;
    }
    uint32_t h = cache_hash(key[0], key[1], key[2], key[3]) & a.cache.mask;
    for (uint32_t probes = 0; probes < CACHE_MAX_PROBES; probes++) {
        uint4 k0, k1, v0, v1; // the whole 64-byte slot: two 32-byte loads in flight together
        ld_slot256(&a.cache.slots[h], k0, k1);
        ld_slot256(reinterpret_cast<const uint8_t *>(&a.cache.slots[h]) + 32, v0, v1);
        if ((k1.z | k1.w) == 0) return 0; // k[3] == 0: empty
        if (k0.x == (uint32_t)key[0] && k0.y == (uint32_t)(key[0] >> 32) && k0.z == (uint32_t)key[1] &&
            k0.w == (uint32_t)(key[1] >> 32) && k1.x == (uint32_t)key[2] && k1.y == (uint32_t)(key[2] >> 32) &&
            k1.z == (uint32_t)key[3] && k1.w == (uint32_t)(key[3] >> 32)) {
            const uint32_t n = v0.x;
            if (n <= 4) {
                v = make_uint4(v0.y, v0.z, v0.w, v1.x);
                return n;
            }
            const uint32_t at = atomicAdd(&sm.park_used, n);
            v.x = at;
            if (n <= CACHE_INLINE_IDS) {
                const uint32_t ids[CACHE_INLINE_IDS] = {v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                for (uint32_t i = 0; i < n; i++) park_store(a, sm, at + i, ids[i]);
            } else {
                const uint32_t *src = a.cache.arena + v0.y;
                for (uint32_t i = 0; i < n; i++) park_store(a, sm, at + i, __ldg(&src[i]));
            }
            return n;
        }
        h = (h + 1) & a.cache.mask;
    }
    return 0;
}

// The chunks nobody has seen before (the tile's scan list), by all threads of the CTA; the caller's barrier follows.
// Few (warm caches): latency matters -- one WARP per chunk, lanes = positions, one lookup latency per pass. Many (cold
// caches): throughput matters -- one THREAD per chunk; also for chunks of 33..64 bytes.
template <int THREADS, class SM>
__device__ __noinline__ void scan_open_chunks(const EncArgs &a, SM &sm, uint32_t a0, bool staged, uint32_t n_scan, bool log) {
    constexpr int NW = THREADS / 32;
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0 && log) atomicAdd(a.miss_count, n_scan);
    if (n_scan > ET_WARP_SCAN_MAX) {
        for (uint32_t s = tid; s < n_scan; s += THREADS) {
            const uint32_t q = s, k = sm.open_k[q], so = sm.off[k];
            sm.meta[q] = scan_serial(a, sm, a0, staged, so, sm.off[k + 1] - so, log);
        }
        return;
    }
    for (uint32_t s = warp; s < n_scan; s += NW) {
        const uint32_t q = s, k = sm.open_k[q];
        if (sm.off[k + 1] - sm.off[k] > 32) { // 33..64 bytes do not fit the lanes
            if ((tid & 31) == 0) sm.meta[q] = scan_serial(a, sm, a0, staged, sm.off[k], sm.off[k + 1] - sm.off[k], log);
            __syncwarp();
            continue;
        }
        scan_by_warp(a, sm, a0, staged, q, sm.warp_scratch[warp], log);
    }
}

// one chunk's ids -> dst[0 .. n) (shared-memory gather buffer, or the stream itself for oversized tiles).
// openq: the chunk's place on the open list, or TILE_NONE: its ids are in v (<= 4) or parked at v.x (more)
template <class SM>
__device__ __forceinline__ void emit_chunk(const EncArgs &a, SM &sm, uint32_t n, uint32_t openq, uint32_t o0, uint32_t o1, uint4 v,
                                           uint32_t *dst, uint64_t room) {
    if (n == 0) return;
    if (n > room) {
        *a.overflow = 1;
        return;
    }
    if (openq == TILE_NONE && o1 - o0 > ENC_SHORT_MAX) {
        for (uint32_t i = 0; i < n; i++) dst[i] = a.scratch_a[o0 + i];
    } else if (openq == TILE_NONE && n <= 4) {
        dst[0] = v.x;
        if (n > 1) dst[1] = v.y;
        if (n > 2) dst[2] = v.z;
        if (n > 3) dst[3] = v.w;
    } else {
        const uint32_t start = openq == TILE_NONE ? v.x : (sm.meta[openq] & 0xFFFFF);
        for (uint32_t i = 0; i < n; i++) dst[i] = park_load(a, sm, start + i);
    }
}

// thread 0, first half of a tile fetch: the tile's boundaries on their way into off_buf[buf]
template <class SM>
__device__ __forceinline__ void fetch_off(const EncArgs &a, SM &sm, uint32_t tile, uint32_t buf, uint64_t policy) {
    const uint64_t c0 = a.chunk0 + (uint64_t)tile * SM::TILE;
    const uint32_t nc = (uint32_t)min((uint64_t)SM::TILE, a.chunk1 - c0);
    const uint32_t nw = nc + 1, nb = nw & ~3u; // whole 16-byte vectors by bulk copy, the last <= 3 words by hand
    if (nb) {
        mbar_arrive_expect_tx(&sm.bar_off[buf], nb * 4);
        bulk_g2s(sm.off_buf[buf], a.off + c0, nb * 4, &sm.bar_off[buf], policy);
    } else {
        mbar_arrive(&sm.bar_off[buf]);
    }
    for (uint32_t i = nb; i < nw; i++) sm.off_buf[buf][i] = __ldg(&a.off[c0 + i]);
}
// second half, once the boundaries have landed: the text window into text_buf[buf]; completes bar_txt[buf]
template <class SM>
__device__ __forceinline__ void fetch_text(const EncArgs &a, SM &sm, uint32_t tile, uint32_t buf, uint32_t off_parity, uint64_t policy) {
    mbar_wait(&sm.bar_off[buf], off_parity);
    const uint64_t c0 = a.chunk0 + (uint64_t)tile * SM::TILE;
    const uint32_t nc = (uint32_t)min((uint64_t)SM::TILE, a.chunk1 - c0);
    const uint32_t b0 = sm.off_buf[buf][0], b1 = sm.off_buf[buf][nc];
    const uint32_t a0 = b0 & ~15u; // 16-byte aligned window start
    const uint32_t span = b1 - a0;
    const bool staged = span <= (uint32_t)SM::TEXT_CAP;
    sm.a0[buf] = a0;
    sm.staged[buf] = staged;
    if (staged) {
        // whole vectors that lie inside the buffer by bulk copy, the last (< 16) bytes of the buffer by hand
        const uint64_t avail = (a.n_bytes_total - a0) & ~15ull;
        const uint32_t full = (uint32_t)min((uint64_t)((span + 15) & ~15u), avail);
        for (uint32_t g = a0 + full; g < b1; g++) reinterpret_cast<uint8_t *>(sm.text_buf[buf])[g - a0] = __ldg(&a.bytes[g]);
        if (full) {
            mbar_arrive_expect_tx(&sm.bar_txt[buf], full);
            bulk_g2s(sm.text_buf[buf], a.bytes + a0, full, &sm.bar_txt[buf], policy);
            return;
        }
    }
    mbar_arrive(&sm.bar_txt[buf]);
}

// PASS 1: count (tile_total). PASS 2: write (ids at tile_base).
template <int THREADS, int CPT, int MIN_CTAS, int PASS>
__global__ void __launch_bounds__(THREADS, MIN_CTAS) k_encode_tiles(const EncArgs a) {
    using SM = EncSmemT<THREADS, CPT>;
    constexpr int TILE = SM::TILE, NW = THREADS / 32;
    extern __shared__ __align__(128) unsigned char enc_smem_raw[];
    SM &sm = *reinterpret_cast<SM *>(enc_smem_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool bulk = a.bulk != 0;
    if (tid == 0) {
        for (int b = 0; b < 2; b++) {
            mbar_init(&sm.bar_off[b], 1);
            mbar_init(&sm.bar_txt[b], 1);
        }
        mbar_init_fence();
        sm.n_open = sm.park_used = 0;
        for (int i = 0; i < 8; i++) sm.prof[i] = 0;
    }
    __syncthreads();
    long long t_lap = clock64();
    auto lap = [&](int i) { // thread 0: cycles since the previous lap go to counter i
        if (a.prof && tid == 0) {
            const long long t = clock64();
            sm.prof[i] += (unsigned long long)(t - t_lap);
            t_lap = t;
        }
    };
    const uint64_t policy = l2_evict_first_policy();
    // tiles of this CTA: blockIdx.x, + gridDim.x, ... (no tile depends on another one)
    if (bulk && tid == 0 && blockIdx.x < a.n_tiles) {
        fetch_off(a, sm, blockIdx.x, 0, policy);
        fetch_text(a, sm, blockIdx.x, 0, 0, policy);
        if (blockIdx.x + gridDim.x < a.n_tiles) fetch_off(a, sm, blockIdx.x + gridDim.x, 1, policy);
    }
    uint32_t it = 0;
    for (uint32_t tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, it++) {
        const uint32_t buf = it & 1, par = (it >> 1) & 1; // buffer of this tile, and how often it has been used before (parity)
        const uint64_t c0 = a.chunk0 + (uint64_t)tile * TILE;
        const uint32_t nc = (uint32_t)min((uint64_t)TILE, a.chunk1 - c0);
        // ---- 0. the tile's boundaries and text in shared memory ----------------------------------------------------
        if (bulk) {
            mbar_wait(&sm.bar_txt[buf], par);
        } else { // unaligned caller buffers: cooperative loads, no prefetch
            __syncthreads();
            uint32_t *ob = sm.off_buf[buf];
            for (uint32_t i = tid; i <= nc; i += THREADS) ob[i] = __ldg(&a.off[c0 + i]);
            __syncthreads();
            const uint32_t b0 = ob[0], b1 = ob[nc], w0 = b0 & ~3u;
            const bool st = (b1 - w0) <= (uint32_t)SM::TEXT_CAP;
            if (st) // byte loads: nothing is known about the alignment of the buffer
                for (uint32_t g = b0 + tid; g < b1; g += THREADS) reinterpret_cast<uint8_t *>(sm.text_buf[buf])[g - w0] = __ldg(&a.bytes[g]);
            if (tid == 0) {
                sm.a0[buf] = w0;
                sm.staged[buf] = st;
            }
            __syncthreads();
        }
        sm.off = sm.off_buf[buf]; // (every thread stores the same two pointers: a reader has written them itself or sees the
        sm.text = sm.text_buf[buf]; //  identical value; every tile ends with a barrier, so nobody still reads the previous ones)
        const uint32_t *const soff = sm.off_buf[buf];
        const uint32_t *const stext = sm.text_buf[buf];
        const uint32_t a0 = sm.a0[buf];
        const bool staged = sm.staged[buf] != 0;
        lap(0);
        // ---- 1. fast path: the home slot of every chunk in the SMALL cache, all of a thread's probes in flight ------
        uint32_t o[CPT + 1];
#pragma unroll
        for (int j = 0; j <= CPT; j++) o[j] = soff[min(tid * CPT + j, nc)];
        uint32_t cnt[CPT];
        uint32_t openq[CPT]; // place on the open list, or TILE_NONE: cnt / vq are final
        uint4 vq[CPT];       // hit: the chunk's ids
        uint32_t hints = 0;  // two bits per chunk: what its home slot in the SMALL cache showed (see resolve_cached)
        const bool keyed = a.cache.small != nullptr && staged && !(a.ablate & 2); // (an unstaged tile -- very long chunks -- is all "open")
        if (keyed) {
            uint4 kw[CPT], kq[CPT];
#pragma unroll
            for (int j = 0; j < CPT; j++) {
                const uint32_t len = o[j + 1] - o[j];
                // the 16 bytes at the chunk's start, bytes at and after `len` cleared, length in the top byte
                const uint32_t r = o[j] - a0, wi = r >> 2, sh = (r & 3) * 8;
                const uint32_t nb = len < 16 ? len : 0; // ones over the key's first `len` bytes, word by word
                uint4 lm;
                lm.x = (uint32_t)((1ull << (8 * min(nb, 4u))) - 1);
                lm.y = (uint32_t)((1ull << (8 * (min(max(nb, 4u), 8u) - 4))) - 1);
                lm.z = (uint32_t)((1ull << (8 * (min(max(nb, 8u), 12u) - 8))) - 1);
                lm.w = (uint32_t)((1ull << (8 * (min(max(nb, 12u), 16u) - 12))) - 1);
                const uint32_t t0 = stext[wi], t1 = stext[wi + 1], t2 = stext[wi + 2], t3 = stext[wi + 3], t4 = stext[wi + 4];
                kw[j].x = __funnelshift_r(t0, t1, sh) & lm.x;
                kw[j].y = __funnelshift_r(t1, t2, sh) & lm.y;
                kw[j].z = __funnelshift_r(t2, t3, sh) & lm.z;
                kw[j].w = (__funnelshift_r(t3, t4, sh) & lm.w) | (len << 24);
                const uint32_t h = small_hash(kw[j].x, kw[j].y, kw[j].z, kw[j].w) >> a.cache.small_shift;
                if (PASS == 1) // (counting needs the key half only; empty / long chunks probe too: the answer is ignored)
                    kq[j] = __ldg(reinterpret_cast<const uint4 *>(&a.cache.small[h]));
                else
                    ld_slot256(&a.cache.small[h], kq[j], vq[j]);
            }
#pragma unroll
            for (int j = 0; j < CPT; j++) {
                const uint32_t len = o[j + 1] - o[j];
                const bool small = len - 1u < SMALL_MAX_LEN;
                const bool hit = small && (kq[j].w & SMALL_KEY_MASK) == kw[j].w && kq[j].x == kw[j].x && kq[j].y == kw[j].y && kq[j].z == kw[j].z;
                const bool final = hit && (kq[j].w >> 28) <= SMALL_MAX_IDS; // (a stub is a hit that only says "BIG has it")
                cnt[j] = final ? kq[j].w >> 28 : 0u;
                openq[j] = final ? TILE_NONE : 0u; // 0: undecided, see below
                // what the home slot showed, for resolve_cached: 1 = another chunk, 2 = this chunk's stub (or a long chunk)
                hints |= (hit ? 2u : (small && kq[j].w != 0) ? 1u : (small ? 3u : 2u)) << (2 * j); // 3: empty home slot = not cached
            }
        } else {
#pragma unroll
            for (int j = 0; j < CPT; j++) {
                cnt[j] = 0;
                openq[j] = 0;
                vq[j] = make_uint4(0, 0, 0, 0);
            }
            hints = 0; // (no home slot was looked at: the whole probe sequence is open)
        }
        // the next tile's boundaries have had a whole fast path to land: start its text on its way, and the boundaries
        // of the tile after it (their buffers belong to tiles that are completely done)
        if (bulk && tid == 0) {
            const uint32_t t1 = tile + gridDim.x;
            if (t1 < a.n_tiles) fetch_text(a, sm, t1, buf ^ 1, ((it + 1) >> 1) & 1, policy);
        }
        // open chunks: bit j of `open`. First the ones that need no look-up at all ...
        uint32_t open = 0;
#pragma unroll
        for (int j = 0; j < CPT; j++) {
            if (openq[j] == TILE_NONE) continue; // hit
            openq[j] = TILE_NONE;
            const uint32_t k = tid * CPT + j, len = o[j + 1] - o[j];
            if (k >= nc || len == 0) continue;
            if (len > ENC_SHORT_MAX) {
                if (a.scratch_b) {
                    cnt[j] = a.scratch_b[o[j]]; // encoded by k_encode_long
                } else if (PASS == 1) {         // optimistic launch: report it, the host runs the long path and repeats
                    const uint32_t q = atomicAdd(a.n_long, 1u);
                    if (q < a.long_cap) a.long_list[q] = (uint32_t)(c0 + k);
                }
                continue;
            }
            if (a.ablate & 10) { // (profiling only: 2 = every chunk "hits", 8 = open chunks are not resolved)
                cnt[j] = (a.ablate & 2) ? 2 : 1;
                vq[j] = make_uint4(o[j], len, 0, 0);
                continue;
            }
            open |= 1u << j;
        }
        // ... then, one open chunk per lane and round (a lane rarely has two), the ones that may be cached elsewhere: a special
        // token; a chunk of 16..31 bytes or a stub (BIG); a short one whose home slot holds somebody else. No barrier: the lanes
        // that have one diverge into resolve_cached for about one memory round trip.
        for (uint32_t todo = open; todo;) {
            const int j = __ffs(todo) - 1;
            todo &= todo - 1;
            uint32_t oj = 0, ej = 0;
#pragma unroll
            for (int q = 0; q < CPT; q++)
                if (q == j) {
                    oj = o[q];
                    ej = o[q + 1];
                }
            const uint32_t len = ej - oj, hint = (hints >> (2 * j)) & 3u;
            const bool maybe_cached = a.cache.small != nullptr && len <= CACHE_MAX_LEN && hint != 3u;
            if (!(maybe_cached || (a.sp.n && ((a.sp.len_mask >> (len < 63u ? len : 63u)) & 1ull)))) continue;
            uint4 r = make_uint4(0, 0, 0, 0);
            const uint32_t n = resolve_cached(a, sm, a0, staged, oj, len, hint == 3u ? 0u : hint, r);
            if (n == 0) continue;
            open &= ~(1u << j);
#pragma unroll
            for (int q = 0; q < CPT; q++)
                if (q == j) {
                    cnt[q] = n;
                    vq[q] = r;
                }
        }
        // ... and what nobody has seen before goes on the tile's scan list
#pragma unroll
        for (int j = 0; j < CPT; j++) {
            if (!((open >> j) & 1u)) continue;
            const uint32_t q = atomicAdd(&sm.n_open, 1u);
            sm.open_k[q] = (uint16_t)(tid * CPT + j);
            openq[j] = q;
        }
        __syncthreads();
        lap(1);
        // ---- 2. chunks nobody has seen before ----------------------------------------------------------------------------
        {
            const uint32_t n_scan = sm.n_open;
            if (n_scan) {
                scan_open_chunks<THREADS>(a, sm, a0, staged, n_scan, PASS == 1);
                __syncthreads();
            }
        }
        lap(2);
        uint32_t sum = 0;
#pragma unroll
        for (int j = 0; j < CPT; j++) {
            if (openq[j] != TILE_NONE) cnt[j] = sm.meta[openq[j]] >> 20;
            sum += cnt[j];
        }
        // ---- 3. the tile's id count (pass 1) / every thread's place in the tile (pass 2) -----------------------------
        uint32_t incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        if (lane == 31) sm.warp_sum[warp] = incl;
        __syncthreads();
        uint32_t warp_base = 0, total = 0;
#pragma unroll
        for (int w = 0; w < NW; w++) {
            const uint32_t v = sm.warp_sum[w];
            if (w < (int)warp) warp_base += v;
            total += v;
        }
        lap(3);
        if (tid == 0) sm.n_open = sm.park_used = 0; // (everybody is past the open list; the next appends come after a barrier below)
        if (PASS == 1) {
            if (tid == 0) a.tile_total[tile] = total;
            __syncthreads();
        } else {
            const uint64_t base = a.tile_base[tile];
            const bool via_smem = total <= (uint32_t)SM::STAGE;
            const uint32_t loc0 = warp_base + (incl - sum);
            if (!(a.ablate & 4)) {
                uint32_t loc = loc0;
#pragma unroll
                for (int j = 0; j < CPT; j++) {
                    const uint64_t at = base + loc;
                    if (via_smem)
                        emit_chunk(a, sm, cnt[j], openq[j], o[j], o[j + 1], vq[j], sm.stage + loc, ~0ull);
                    else // a tile with more ids than the gather buffer holds: every thread stores its own
                        emit_chunk(a, sm, cnt[j], openq[j], o[j], o[j + 1], vq[j], a.out + at, at < a.out_cap ? a.out_cap - at : 0);
                    loc += cnt[j];
                }
            }
            if (a.out_off) {
                uint32_t loc = loc0;
#pragma unroll
                for (int j = 0; j < CPT; j++) {
                    const uint32_t k = tid * CPT + j;
                    if (k < nc) a.out_off[c0 + k] = base + loc;
                    loc += cnt[j];
                }
            }
            __syncthreads();
            lap(4);
            if (via_smem && !(a.ablate & 4)) {
                if (base + total <= a.out_cap && a.out_aligned) {
                    // 16-byte stores: vector v = stream words [w0 + 4v, w0 + 4v + 4), w0 = base rounded down to 4 words
                    const uint32_t pad = (uint32_t)(base & 3);
                    uint32_t *const gout = a.out + (base - pad);
                    const uint32_t n_vec = (pad + total + 3) >> 2;
                    for (uint32_t v = tid; v < n_vec; v += THREADS) {
                        const int lo = (int)(4 * v) - (int)pad; // index of the vector's first word in the gather buffer
                        if (lo >= 0 && (uint32_t)lo + 4 <= total) {
                            const uint4 q = make_uint4(sm.stage[lo], sm.stage[lo + 1], sm.stage[lo + 2], sm.stage[lo + 3]);
                            __stcs(reinterpret_cast<uint4 *>(gout) + v, q);
                        } else {
                            for (int i = lo < 0 ? 0 : lo; i < lo + 4 && (uint32_t)i < total; i++) __stcs(&a.out[base + i], sm.stage[i]);
                        }
                    }
                } else {
                    for (uint32_t i = tid; i < total; i += THREADS)
                        if (base + i < a.out_cap) a.out[base + i] = sm.stage[i];
                    if (tid == 0 && base + total > a.out_cap) *a.overflow = 1;
                }
            }
            if (tile == a.n_tiles - 1 && tid == 0 && a.out_off && a.chunk1 == a.n_chunks) a.out_off[a.n_chunks] = base + total;
        }
        if (PASS == 2) __syncthreads(); // every tile ends with a barrier: nothing of it is still read when the next one starts
        // the boundaries of the tile after the next one go into this tile's buffer
        if (bulk && tid == 0) {
            const uint32_t t2 = tile + 2 * gridDim.x;
            if (t2 < a.n_tiles) fetch_off(a, sm, t2, buf, policy);
        }
        lap(6);
        if (a.prof && tid == 0) sm.prof[7] += 1;
    }
    if (a.prof && tid == 0)
        for (int i = 0; i < 8; i++) atomicAdd(&a.prof[(PASS - 1) * 8 + i], sm.prof[i]);
}

// tile_total[0 .. n) -> tile_base[0 .. n) (exclusive scan, starting at the ids of earlier launches); *d_n_out = the new total.
// One CTA: a launch has at most 2^26 / 256 tiles.
__global__ void __launch_bounds__(1024) k_tile_scan(const uint32_t *tile_total, uint32_t n, unsigned long long *tile_base,
                                                   unsigned long long *d_n_out) {
    __shared__ unsigned long long warp_sum[32];
    __shared__ unsigned long long carry;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = *d_n_out;
    __syncthreads();
    for (uint32_t i0 = 0; i0 < n; i0 += 1024 * 4) {
        uint32_t v[4];
        unsigned long long mine = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t i = i0 + tid * 4 + q;
            v[q] = i < n ? tile_total[i] : 0u;
            mine += v[q];
        }
        unsigned long long incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long x = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += x;
        }
        if (lane == 31) warp_sum[warp] = incl;
        __syncthreads();
        unsigned long long wb = 0, all = 0;
        for (int w = 0; w < 32; w++) {
            const unsigned long long x = warp_sum[w];
            if (w < (int)warp) wb += x;
            all += x;
        }
        unsigned long long at = carry + wb + (incl - mine);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t i = i0 + tid * 4 + q;
            if (i < n) tile_base[i] = at;
            at += v[q];
        }
        __syncthreads();
        if (tid == 0) carry += all;
        __syncthreads();
    }
    if (tid == 0) *d_n_out = carry;
}

} // namespace mbpe
