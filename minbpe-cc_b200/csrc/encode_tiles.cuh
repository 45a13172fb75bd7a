// encode_tiles.cuh -- k_encode_tiles: the encode merge scan over tiles of chunks (included by encode.cu only).
//
// Replaces internal_internal_encode / internal_encode + flatten (Tokenizer.h:325-377, :714-717) for chunks of up to
// ENC_SHORT_MAX bytes; longer chunks are encoded by k_encode_long into a scratch stream and spliced in here.
//
// One tile = THREADS x CPT consecutive chunks, one CTA, one pass over the data:
//   0. thread 0 takes the next tile (ticket: tiles START in stream order, which is what makes the look-back of step 3
//      deadlock free) and brings its boundaries and its text window into shared memory with two bulk asynchronous copies
//      (cp.async.bulk -> the TMA engine, completion on mbarriers), while the other threads still store the previous
//      tile's ids. The ticket is taken as late as possible: see fetch_tile_bulk.
//   1. every thread, CPT chunks: the chunk's bytes (<= 30) are the key of the chunk cache, "chunk bytes -> ids" in
//      64-byte slots. The key sector also carries the id count and, for keys of <= 16 bytes, up to three ids, so ONE
//      256-bit load (LDG.E.256) answers nine chunks in ten, and every lane runs the same instructions. What a thread
//      learns about a chunk goes into the chunk's 8-byte record in shared memory, not into registers: the probe loop
//      needs 40 registers, neighbouring lanes work on neighbouring chunks (conflict-free shared memory traffic) and
//      their ids land side by side in step 3.
//   2. chunks the cache does not hold (cold cache, chunks of 31..64 bytes, a full table) go to the tile's scan list: the
//      multi-pass scan itself, one warp per chunk when few, one thread per chunk when many; their ids wait in a parking
//      area (shared memory, overflow in HBM), and a log hands them to k_cache_insert, which runs between launches.
//   3. block scan of the id counts; warp 0 announces the tile's count, everybody gathers the tile's ids in shared memory
//      (ids that live in the value sector of a cache slot: all of a thread's loads in flight together), and only then
//      warp 0 walks the predecessors' counts (decoupled look-back, 128 per round trip) -- by then most of them have
//      announced themselves. The ids leave as 16-byte stores.
// Results never depend on the cache: a miss is scanned, and special tokens are matched in the scan path itself.
// What bounds it (profiles/README.md, round 2): not bytes -- DRAM traffic is the algorithmic 13 bytes per chunk -- but the
// chain of latencies of one tile (ticket -> boundaries -> text -> four dependent probe rounds -> look-back) at the
// occupancy 64 registers x 256 threads x 4 CTAs allow; DESIGN.md "encode: what was tried" lists the measured dead ends.
#pragma once
#include <type_traits>
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

// one 32-byte sector in ONE load (LDG.E.256, new with sm_100): one L1 wavefront instead of two.
// LD = how the load treats L1: 0 read-only path, allocates (LDG.CONSTANT); 1 L2 only (ld.cg); 2 read-only path, no L1
// allocation (LDG.NA); 3 read-only path, evict-last in L1 (LDG.EL)
template <int LD = 0>
__device__ __forceinline__ void ld_sector256(const void *p, uint64_t &q0, uint64_t &q1, uint64_t &q2, uint64_t &q3) {
    unsigned long long a0, a1, a2, a3;
    if (LD == 0) asm("ld.global.nc.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(a0), "=l"(a1), "=l"(a2), "=l"(a3) : "l"(p));
    if (LD == 1) asm("ld.global.cg.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(a0), "=l"(a1), "=l"(a2), "=l"(a3) : "l"(p));
    if (LD == 2) asm("ld.global.nc.L1::no_allocate.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(a0), "=l"(a1), "=l"(a2), "=l"(a3) : "l"(p));
    if (LD == 3) asm("ld.global.nc.L1::evict_last.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(a0), "=l"(a1), "=l"(a2), "=l"(a3) : "l"(p));
    q0 = a0;
    q1 = a1;
    q2 = a2;
    q3 = a3;
}

// ---------------------------------------------------------------------------------------------------------
// Chunk cache. The multi-pass scan of a chunk is a pure function of its bytes, and text repeats its chunks (Zipf):
// the encoder keeps "chunk bytes (<= 30) -> ids" in an open-addressed table in HBM, 64-byte slots = key sector + value
// sector (more than 7 ids: in an arena); nine chunks in ten are answered by the key sector alone (see CacheSlot). The tile kernel only READS it; chunks it had to scan go to a log, and
// k_cache_insert adds the log to the table between launches -- no kernel both reads and writes the table, so there is no
// publication protocol to get wrong. Results are bit-identical with or without the cache (MBPE_ENCODE_CACHE=0 disables
// it; tests run both).
// ---------------------------------------------------------------------------------------------------------
struct CacheSlot { // 64 bytes = two sectors
    // Key sector. Chunk bytes little endian, zero padded; top byte of k[3] = length (1..30); k[3] == 0: empty slot.
    // The id count rides in the key sector, so that a probe is ONE 32-byte load:
    //   SHORT key (<= 16 bytes): k[0], k[1] = the bytes; k[3] bits 0..6 = id count, bit 7 = "ids not inline";
    //                            k[2] = up to three ids of 21 bits (nine chunks in ten are answered by this sector alone);
    //   LONG key (17..30 bytes): k[0..2] and k[3] bits 0..47 = the bytes; k[3] bits 48..55 = id count.
    uint64_t k[4];
    // Value sector (read when the ids are not inline: long keys, more than three ids, ids of more than 21 bits).
    uint32_t n;    // number of ids (1..30)
    uint32_t v[7]; // n <= 7: the ids; otherwise v[0] = offset of the ids in the arena
};
static_assert(sizeof(CacheSlot) == 64, "two sectors per entry");
struct CacheLogEntry {
    uint64_t k[4];
    uint32_t n;
    uint32_t ids[31];
};
constexpr uint32_t CACHE_MAX_LEN = 30, CACHE_SHORT_KEY = 16, CACHE_INLINE_IDS = 7, CACHE_KEY_IDS = 3, CACHE_ID_BITS = 21;
constexpr uint64_t CACHE_NOT_INLINE = 0x80;
constexpr uint32_t CACHE_LONG_N_SHIFT = 48;
constexpr uint64_t CACHE_LONG_N_MASK = 0xFFull << CACHE_LONG_N_SHIFT;
// An entry lives at most this many slots from its home: inserts give up beyond it (the chunk is simply not cached), so
// lookups may stop there too -- no probe loop depends on the table having a free slot.
constexpr uint32_t CACHE_MAX_PROBES = 64;

struct ChunkCache {
    CacheSlot *slots;    // nullptr = cache disabled
    uint32_t shift;      // home slot = hash >> shift (the top bits of a multiplicative hash are the good ones)
    uint32_t mask;       // slots - 1
    CacheLogEntry *log;  // chunks the current launch had to scan
    uint32_t *log_count;
    uint32_t log_cap;
    uint32_t *used;      // occupied slots (learning stops at half full)
    uint32_t *arena;     // ids of entries with more than CACHE_INLINE_IDS ids
    uint32_t *arena_used;
    uint32_t arena_cap;
};

// Two multiplies for the chunks of <= 16 bytes (k2 == 0, k3 == length only), two more for the longer ones. The length
// is not hashed ("a" and "a\0" share a home slot; the key compare tells them apart).
__host__ __device__ __forceinline__ uint64_t cache_hash_lo(uint64_t k0, uint64_t k1) {
    const uint64_t m = k1 * 0x9E3779B97F4A7C15ull;
    return k0 ^ ((m >> 32) | (m << 32));
}
__host__ __device__ __forceinline__ uint64_t cache_hash_hi(uint64_t k2, uint64_t k3_bytes) { // k3 without its length byte
    const uint64_t m = k2 * 0xC2B2AE3D27D4EB4Full;
    return ((m >> 32) | (m << 32)) + k3_bytes * 0x165667B19E3779F9ull;
}
__host__ __device__ __forceinline__ uint32_t cache_hash_fin(uint64_t x) { return (uint32_t)((x * 0xD6E8FEB86659FD93ull) >> 32); }
__host__ __device__ __forceinline__ uint32_t cache_hash(uint64_t k0, uint64_t k1, uint64_t k2, uint64_t k3) {
    uint64_t x = cache_hash_lo(k0, k1);
    const uint64_t b3 = k3 & 0x00FFFFFFFFFFFFFFull;
    if (k2 | b3) x ^= cache_hash_hi(k2, b3);
    return cache_hash_fin(x);
}

__global__ void k_cache_insert(ChunkCache cc) {
    const uint32_t n = min(*cc.log_count, cc.log_cap);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const CacheLogEntry &e = cc.log[i];
        const uint64_t k0 = e.k[0], k1 = e.k[1];
        uint64_t k2 = e.k[2], k3 = e.k[3];
        const uint32_t en = e.n;
        if (*((volatile uint32_t *)cc.used) * 2 > cc.mask) return; // half full: stop learning
        uint32_t h = cache_hash(k0, k1, k2, k3) >> cc.shift;
        const bool short_key = (k3 >> 56) <= CACHE_SHORT_KEY;
        if (short_key) { // the answer rides in the key sector (a pure function of the key, like the key words themselves)
            bool inline_ids = en <= CACHE_KEY_IDS;
            for (uint32_t q = 0; q < en && inline_ids; q++) inline_ids = e.ids[q] < (1u << CACHE_ID_BITS);
            k3 |= en | (inline_ids ? 0 : CACHE_NOT_INLINE);
            if (inline_ids)
                for (uint32_t q = 0; q < en; q++) k2 |= (uint64_t)e.ids[q] << (CACHE_ID_BITS * q);
        } else {
            k3 |= (uint64_t)en << CACHE_LONG_N_SHIFT;
        }
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
                if (cur == 0) { // claimed: k[3] is the claim word, the rest is written by the winner only
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
            // Same key: already there (the log holds duplicates of hot chunks). Words of a slot claimed in THIS launch
            // may not be visible yet; then the duplicate takes a second slot -- harmless, both slots hold the same ids
            // and a reader uses the first one it finds.
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
    unsigned long long *status;  // look-back words, one per tile, zeroed
    uint32_t *ticket;            // zeroed
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
    const uint4 *hot_img;        // k_encode_hot: image of the shared-memory table of the hottest chunks (encode_hot.cuh)
    uint32_t pf_ahead;           // MBPE_ENC_PF: the boundaries (exact) and the text (estimated) of the tile this many tickets ahead are asked into L2
    uint32_t ablate;             // MBPE_ENC_ABLATE (profiling only, WRONG results): 1 no look-back wait, 2 no cache probe
                                 // (every short chunk "hits" with two fake ids), 4 no id stores
};

// EncArgs::prof[i]: cycles thread 0 spent ... 0 waiting for the tile's data, 1 cache probes (+ barrier), 2 scan list,
// 3 count scan (+ barrier), 4 gather + look-back (+ barrier), 5 fetching the next tile (ticket + two bulk copies),
// 6 storing the ids, 7 tiles processed; inside 4: 8 look-back walk, 9 gather, 10 barrier (thread 0), 11 gather, 12 barrier
// (thread 32: what the other warps wait for warp 0)
constexpr int ENC_PROF_N = 16;
constexpr uint32_t TILE_NONE = 0xFFFFFFFFu;
constexpr uint64_t ENC_MAX_SUBBATCH = 1ull << 26; // chunks per launch at most (tile ids and look-back words are 32-bit safe)
constexpr uint32_t ET_WARP_SCAN_MAX = 48;     // up to this many scans per tile run one warp per chunk
// what the tile knows about a chunk after the probes (EncSmemT::rec)
constexpr uint32_t CK_REGS = 0;   // ids (<= CACHE_KEY_IDS) in the record itself, or no ids at all
constexpr uint32_t CK_SLOT = 1;   // cached with more ids: the value sector of its slot is fetched again when writing
constexpr uint32_t CK_PARKED = 2; // scanned by the tile: ids in the parking area
constexpr uint32_t CK_LONG = 3;   // longer than ENC_SHORT_MAX: ids in the long-chunk scratch stream
constexpr uint32_t REC_REF = 0x80000000u; // EncSmemT::rec[].y: not CK_REGS

// Parking area of a tile = ids of its scanned chunks until they are written: the first PARK words live in shared memory,
// the rest in the CTA's spill block in HBM (one index space; every id covers at least one byte of text, so a tile of
// ENC_SHORT_MAX-byte chunks at most needs TILE * ENC_SHORT_MAX words: it always fits, nothing is ever encoded twice).
template <int THREADS, int CPT>
struct EncSmemT {
    static constexpr int TILE = THREADS * CPT;
    static constexpr int TEXT_CAP = TILE * 8;   // staged text bytes per tile (average chunk ~5 bytes); wider tiles read HBM
    static constexpr int STAGE = TILE * 5 / 2;  // ids gathered per tile (average ~2.1 per chunk); more: direct stores
    static constexpr int PARK = TILE / 4;       // parking words in shared memory (a warm cache parks a few dozen ids per tile)
    static constexpr int SPILL = TILE * 64;     // parking words per CTA in HBM (EncArgs::spill): the worst case
    union alignas(128) {
        uint32_t off[TILE + 8]; // steps 0 .. 2: the tile's chunk boundaries
        uint32_t loc[TILE];     // step 3: every chunk's offset within its warp's segment of the tile
    };
    alignas(128) uint32_t text[TEXT_CAP / 4 + 16]; // + halo: key assembly reads whole words past the chunk's end
    alignas(16) uint32_t stage[STAGE + 4];
    // per chunk: its id count, and where the ids are: rec.y bit 31 clear = CK_REGS, {x, y} = the key sector's id word (three
    // ids of 21 bits); bit 31 set: y bits 29..30 = CK_* kind, x = cache slot (CK_SLOT) / start in the parking area
    // (CK_PARKED) / text offset (CK_LONG)
    alignas(16) uint2 rec[TILE];
    alignas(16) uint32_t cnt[TILE];
    uint32_t park[PARK];
    uint16_t open_k[TILE];     // scan list: chunk index within the tile
    uint4 len_mask[17];        // len_mask[l] keeps the first l bytes of 16
    uint32_t warp_scratch[THREADS / 32][32];
    uint32_t warp_sum[THREADS / 32];
    alignas(8) uint64_t bar_off, bar_tile; // mbarriers: boundaries landed (thread 0 only waits), tile ready (all wait)
    unsigned long long base;
    uint32_t tile, a0, staged, n_open, park_used, off_span;
    unsigned long long prof[ENC_PROF_N];
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

// cache key of chunk [o, o + len), len <= 31: bytes little endian, zero padded, length in the top byte
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

// a chunk the scan had to encode goes to the log that k_cache_insert folds into the cache after the launch
template <class SM>
__device__ __forceinline__ void log_scanned(const EncArgs &a, const SM &sm, bool staged, uint32_t a0, uint32_t o,
                                            uint32_t len, const uint32_t *ids, uint32_t n) {
    if (!a.cache.slots || len > CACHE_MAX_LEN) return;
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
// result of a scanned chunk: its ids go to the parking area; returns start | n << 20
template <class SM>
__device__ __forceinline__ uint32_t park_ids(const EncArgs &a, SM &sm, const uint32_t *ids, uint32_t n) {
    const uint32_t at = atomicAdd(&sm.park_used, n);
    for (uint32_t i = 0; i < n; i++) park_store(a, sm, at + i, ids[i]);
    return at | (n << 20);
}

// a scanned chunk's record: id count and where its ids are parked (meta = start | n << 20, see park_ids)
template <class SM>
__device__ __forceinline__ void set_parked(SM &sm, uint32_t k, uint32_t meta) {
    sm.rec[k] = make_uint2(meta & 0xFFFFF, REC_REF | (CK_PARKED << 29));
    sm.cnt[k] = meta >> 20;
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
    if (lane == 0) set_parked(sm, k, start | (mn << 20));
    if (log && a.cache.slots && mlen <= CACHE_MAX_LEN && sid == ENC_NONE) { // teach the cache
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

// The chunks the cache does not hold (the tile's scan list), by all threads of the CTA; the caller's barrier follows.
// Few (warm cache): latency matters -- one WARP per chunk, lanes = positions, one lookup latency per pass. Many (cold
// cache): throughput matters -- one THREAD per chunk; also for chunks of 33..64 bytes.
template <int THREADS, class SM>
__device__ __noinline__ void scan_open_chunks(const EncArgs &a, SM &sm, uint32_t a0, bool staged, uint32_t n_scan, bool log) {
    constexpr int NW = THREADS / 32;
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0 && log) atomicAdd(a.miss_count, n_scan);
    if (n_scan > ET_WARP_SCAN_MAX) {
        for (uint32_t s = tid; s < n_scan; s += THREADS) {
            const uint32_t q = s, k = sm.open_k[q], so = sm.off[k];
            set_parked(sm, k, scan_serial(a, sm, a0, staged, so, sm.off[k + 1] - so, log));
        }
        return;
    }
    for (uint32_t s = warp; s < n_scan; s += NW) {
        const uint32_t q = s, k = sm.open_k[q];
        if (sm.off[k + 1] - sm.off[k] > 32) { // 33..64 bytes do not fit the lanes
            if ((tid & 31) == 0) set_parked(sm, k, scan_serial(a, sm, a0, staged, sm.off[k], sm.off[k + 1] - sm.off[k], log));
            __syncwarp();
            continue;
        }
        scan_by_warp(a, sm, a0, staged, q, sm.warp_scratch[warp], log);
    }
}


__device__ __forceinline__ bool rec_is_slot(const uint2 r) { return (r.y >> 29) == (4u | CK_SLOT); }

// one chunk's ids -> dst[0 .. n), n > 0 (shared-memory gather buffer, or the stream itself for oversized tiles); r = its
// record, q0 .. q3 = the value sector of its cache slot if rec_is_slot(r) (the caller has all of a thread's in flight together)
template <class SM>
__device__ __forceinline__ void emit_chunk(const EncArgs &a, SM &sm, const uint2 r, uint32_t n, uint64_t q0, uint64_t q1, uint64_t q2,
                                           uint64_t q3, uint32_t *dst, uint64_t room) {
    const uint32_t kind = (r.y & REC_REF) ? (r.y >> 29) & 3u : CK_REGS;
    if (n > room) {
        *a.overflow = 1;
        return;
    }
    if (kind == CK_REGS) { // {x, y} = the key sector's id word: three ids of 21 bits
        dst[0] = r.x & 0x1FFFFFu;
        if (n > 1) dst[1] = __funnelshift_r(r.x, r.y, CACHE_ID_BITS) & 0x1FFFFFu;
        if (n > 2) dst[2] = r.y >> (2 * CACHE_ID_BITS - 32);
    } else if (kind == CK_SLOT) { // q = n, v[0 .. 6]
        if (n <= CACHE_INLINE_IDS) {
            dst[0] = (uint32_t)(q0 >> 32);
            if (n > 1) dst[1] = (uint32_t)q1;
            if (n > 2) dst[2] = (uint32_t)(q1 >> 32);
            if (n > 3) dst[3] = (uint32_t)q2;
            if (n > 4) dst[4] = (uint32_t)(q2 >> 32);
            if (n > 5) dst[5] = (uint32_t)q3;
            if (n > 6) dst[6] = (uint32_t)(q3 >> 32);
        } else {
            const uint32_t *src = a.cache.arena + (uint32_t)(q0 >> 32);
#pragma unroll 1
            for (uint32_t i = 0; i < n; i++) dst[i] = __ldg(&src[i]);
        }
    } else if (kind == CK_PARKED) {
#pragma unroll 1
        for (uint32_t i = 0; i < n; i++) dst[i] = park_load(a, sm, r.x + i);
    } else {
#pragma unroll 1
        for (uint32_t i = 0; i < n; i++) dst[i] = a.scratch_a[r.x + i];
    }
}

// The ids of this thread's CPT chunks (chunk j * THREADS + tid at place[j] of the tile) -> dst + place[j]. Chunks whose
// ids sit in the value sector of their cache slot fetch it first, all of them before the first use: one round trip per
// thread, not one per chunk.
template <int THREADS, int CPT, class SM>
__device__ __forceinline__ void gather_chunks(const EncArgs &a, SM &sm, const uint32_t *place, uint32_t *dst, uint64_t room) {
    uint64_t q[CPT][4];
#pragma unroll
    for (int j = 0; j < CPT; j++) {
        const uint32_t k = j * THREADS + threadIdx.x;
        const uint2 r = sm.rec[k];
        q[j][0] = q[j][1] = q[j][2] = q[j][3] = 0;
        if (rec_is_slot(r) && sm.cnt[k])
            ld_sector256(reinterpret_cast<const uint8_t *>(&a.cache.slots[r.x]) + 32, q[j][0], q[j][1], q[j][2], q[j][3]);
    }
#pragma unroll
    for (int j = 0; j < CPT; j++) { // (record and count are read again: cheaper than keeping them across the loads)
        const uint32_t k = j * THREADS + threadIdx.x, n = sm.cnt[k];
        if (n) emit_chunk(a, sm, sm.rec[k], n, q[j][0], q[j][1], q[j][2], q[j][3], dst + place[j], room > place[j] ? room - place[j] : 0);
    }
}

// thread 0: ticket, then the tile's boundaries and text window into shared memory by bulk copies. Publishes sm.tile /
// sm.a0 / sm.staged and completes sm.bar_tile when everything has landed.
// The ticket is taken as LATE as possible, when the CTA has nothing else left to do but store its ids. Taking it earlier
// (to prefetch the next tile during the gather or the look-back) was measured twice and is a disaster: a CTA that holds
// a ticket while it waits in a look-back blocks every later tile, whose CTAs then hold THEIR next tickets longer -- the
// look-back went from 4.5 k to 45 k cycles per tile. Tiles must start in stream order, and a ticket must not be held idle.
template <class SM>
__device__ __forceinline__ void fetch_tile_bulk(const EncArgs &a, SM &sm, uint32_t &off_parity) {
    const uint32_t t = atomicAdd(a.ticket, 1u);
    if (t >= a.n_tiles) {
        sm.tile = TILE_NONE;
        mbar_arrive(&sm.bar_tile);
        return;
    }
    const uint64_t policy = l2_evict_first_policy();
    const uint64_t c0 = a.chunk0 + (uint64_t)t * SM::TILE;
    const uint32_t nc = (uint32_t)min((uint64_t)SM::TILE, a.chunk1 - c0);
    const uint32_t nw = nc + 1, nb = nw & ~3u; // whole 16-byte vectors by bulk copy, the last <= 3 words by hand
    if (nb) {
        mbar_arrive_expect_tx(&sm.bar_off, nb * 4);
        bulk_g2s(sm.off, a.off + c0, nb * 4, &sm.bar_off, policy);
    } else {
        mbar_arrive(&sm.bar_off);
    }
    for (uint32_t i = nb; i < nw; i++) sm.off[i] = __ldg(&a.off[c0 + i]);
    mbar_wait(&sm.bar_off, off_parity);
    off_parity ^= 1;
    const uint32_t b0 = sm.off[0], b1 = sm.off[nc];
    const uint32_t a0 = b0 & ~15u; // 16-byte aligned window start
    const uint32_t span = b1 - a0;
    const bool staged = span <= (uint32_t)SM::TEXT_CAP;
    sm.tile = t;
    sm.a0 = a0;
    sm.staged = staged;
    sm.off_span = span;
    if (staged) {
        // whole vectors that lie inside the buffer by bulk copy, the last (< 16) bytes of the buffer by hand
        const uint64_t avail = (a.n_bytes_total - a0) & ~15ull;
        const uint32_t full = (uint32_t)min((uint64_t)((span + 15) & ~15u), avail);
        for (uint32_t g = a0 + full; g < b1; g++) reinterpret_cast<uint8_t *>(sm.text)[g - a0] = __ldg(&a.bytes[g]);
        if (full) {
            mbar_arrive_expect_tx(&sm.bar_tile, full);
            bulk_g2s(sm.text, a.bytes + a0, full, &sm.bar_tile, policy);
            return;
        }
    }
    mbar_arrive(&sm.bar_tile);
}

// PIF = cache probes a thread keeps in flight together (its CPT chunks are probed in groups of PIF).
// Chunk <-> thread: steps 1 and 3b (probe, gather) take chunk j * THREADS + tid for j < CPT -- neighbouring lanes work on
// neighbouring chunks, so their shared-memory traffic is conflict free and their ids land side by side; step 3a (the count
// scan) takes CPT consecutive chunks per thread. Nothing about a chunk is kept in registers between the steps: it is all
// in the chunk's record in shared memory, which is what lets the probe loop run at 64 registers without spilling.
template <int THREADS, int CPT, int MIN_CTAS, int PIF, int LD, int PRE>
__global__ void __launch_bounds__(THREADS, MIN_CTAS) k_encode_tiles(const __grid_constant__ EncArgs a) {
    using SM = EncSmemT<THREADS, CPT>;
    constexpr int TILE = SM::TILE, NW = THREADS / 32;
    static_assert(CPT % PIF == 0 && NW % CPT == 0 && (CPT == 2 || CPT % 4 == 0), "see the chunk <-> thread mapping");
    constexpr int SEG_PER_J = NW / CPT; // warp segments of the count scan covered by one round of THREADS chunks
    extern __shared__ __align__(128) unsigned char enc_smem_raw[];
    SM &sm = *reinterpret_cast<SM *>(enc_smem_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool bulk = a.bulk != 0;
    if (tid < 17) { // len_mask[l]: ones over the first l bytes
        uint32_t m[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int nb = (int)tid - 4 * q;
            m[q] = nb >= 4 ? 0xFFFFFFFFu : nb <= 0 ? 0u : ((1u << (8 * nb)) - 1);
        }
        sm.len_mask[tid] = make_uint4(m[0], m[1], m[2], m[3]);
    }
    if (tid == 0) {
        mbar_init(&sm.bar_off, 1);
        mbar_init(&sm.bar_tile, 1);
        mbar_init_fence();
        sm.n_open = 0;
        sm.park_used = 0;
        for (int i = 0; i < ENC_PROF_N; i++) sm.prof[i] = 0;
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
    uint32_t tile_parity = 0, off_parity = 0;
    if (bulk && tid == 0) fetch_tile_bulk(a, sm, off_parity);
    for (;;) {
        // ---- 0. the tile's boundaries and text in shared memory ----------------------------------------------------
        if (bulk) {
            mbar_wait(&sm.bar_tile, tile_parity);
            tile_parity ^= 1;
        } else { // unaligned caller buffers: ticket and cooperative loads
            __syncthreads();
            if (tid == 0) {
                const uint32_t t = atomicAdd(a.ticket, 1u);
                sm.tile = t < a.n_tiles ? t : TILE_NONE;
                sm.n_open = 0;
                sm.park_used = 0;
            }
            __syncthreads();
        }
        const uint32_t tile = sm.tile;
        if (tile == TILE_NONE) break;
        const uint64_t c0 = a.chunk0 + (uint64_t)tile * TILE;
        const uint32_t nc = (uint32_t)min((uint64_t)TILE, a.chunk1 - c0);
        if (!bulk) {
            for (uint32_t i = tid; i <= nc; i += THREADS) sm.off[i] = __ldg(&a.off[c0 + i]);
            __syncthreads();
            const uint32_t b0 = sm.off[0], b1 = sm.off[nc], w0 = b0 & ~3u;
            const bool st = (b1 - w0) <= (uint32_t)SM::TEXT_CAP;
            if (st) // byte loads: nothing is known about the alignment of the buffer
                for (uint32_t g = b0 + tid; g < b1; g += THREADS) reinterpret_cast<uint8_t *>(sm.text)[g - w0] = __ldg(&a.bytes[g]);
            if (tid == 0) {
                sm.a0 = w0;
                sm.staged = st;
            }
            __syncthreads();
        }
        const uint32_t a0 = sm.a0;
        const bool staged = sm.staged != 0;
        lap(0);
        // ---- 1. cache probes: every chunk's home slot, PIF of them in flight per thread ----------------------------
        const bool keyed = a.cache.slots != nullptr && staged && !(a.ablate & 2); // (an unstaged tile -- very long chunks -- is all scanned)
        // key of chunk [o0, o0 + len) and, if WITH_HASH, its home slot. Chunks of 0 or more than CACHE_MAX_LEN bytes build a key
        // and probe like everybody else -- the text window has a halo, the slot address is always valid -- and ignore the answer.
        auto make_key = [&](auto with_hash, uint32_t o0, uint32_t len, uint64_t &k0, uint64_t &k1, uint64_t &k2, uint64_t &k3) -> uint32_t {
            const uint32_t r = o0 - a0, wi = r >> 2, sh = (r & 3) * 8;
            const uint4 lm = sm.len_mask[min(len, 16u)];
            const uint32_t t0 = sm.text[wi], t1 = sm.text[wi + 1], t2 = sm.text[wi + 2], t3 = sm.text[wi + 3], t4 = sm.text[wi + 4];
            const uint32_t w0 = __funnelshift_r(t0, t1, sh) & lm.x, w1 = __funnelshift_r(t1, t2, sh) & lm.y;
            const uint32_t w2 = __funnelshift_r(t2, t3, sh) & lm.z, w3 = __funnelshift_r(t3, t4, sh) & lm.w;
            k0 = ((uint64_t)w1 << 32) | w0;
            k1 = ((uint64_t)w3 << 32) | w2;
            k2 = 0;
            k3 = (uint64_t)len << 56;
            uint64_t x = 0;
            if (decltype(with_hash)::value) x = cache_hash_lo(k0, k1);
            if (len > CACHE_SHORT_KEY) { // one chunk in a hundred
                const uint4 hm = sm.len_mask[min(len, 32u) - 16];
                const uint32_t t5 = sm.text[wi + 5], t6 = sm.text[wi + 6], t7 = sm.text[wi + 7], t8 = sm.text[wi + 8];
                const uint32_t w4 = __funnelshift_r(t4, t5, sh) & hm.x, w5 = __funnelshift_r(t5, t6, sh) & hm.y;
                const uint32_t w6 = __funnelshift_r(t6, t7, sh) & hm.z, w7 = __funnelshift_r(t7, t8, sh) & hm.w & 0x00FFFFFFu;
                k2 = ((uint64_t)w5 << 32) | w4;
                const uint64_t b3 = ((uint64_t)w7 << 32) | w6;
                k3 |= b3;
                if (decltype(with_hash)::value && (k2 | b3)) x ^= cache_hash_hi(k2, b3);
            }
            return decltype(with_hash)::value ? cache_hash_fin(x) >> a.cache.shift : 0u;
        };
        // PRE: a first pass over the thread's chunks only computes the home slots and asks L2 for them, so that the probes
        // below -- one round trip after the other -- find them there instead of each waiting for DRAM in its turn
        uint32_t hpre[PRE ? CPT : 1];
        if (PRE && keyed) {
#pragma unroll
            for (int j = 0; j < CPT; j++) {
                const uint32_t k = j * THREADS + tid, o0 = sm.off[min(k, nc)], len = sm.off[min(k + 1, nc)] - o0;
                uint64_t k0, k1, k2, k3;
                hpre[j] = make_key(std::true_type{}, o0, len, k0, k1, k2, k3);
                asm volatile("prefetch.global.L2 [%0];" ::"l"(&a.cache.slots[hpre[j]]) : "memory");
            }
        }
#pragma unroll
        for (int g = 0; g < CPT; g += PIF) {
            uint64_t k0[PIF], k1[PIF], k2[PIF], k3[PIF], q0[PIF], q1[PIF], q2[PIF], q3[PIF];
            uint32_t h[PIF], o0[PIF], len[PIF];
#pragma unroll
            for (int p = 0; p < PIF; p++) {
                const uint32_t k = (g + p) * THREADS + tid;
                o0[p] = sm.off[min(k, nc)];
                len[p] = sm.off[min(k + 1, nc)] - o0[p]; // (0 past the tile's end)
                if (keyed) {
                    if (PRE) {
                        make_key(std::false_type{}, o0[p], len[p], k0[p], k1[p], k2[p], k3[p]);
                        h[p] = hpre[g + p];
                    } else {
                        h[p] = make_key(std::true_type{}, o0[p], len[p], k0[p], k1[p], k2[p], k3[p]);
                    }
                    ld_sector256<LD>(&a.cache.slots[h[p]], q0[p], q1[p], q2[p], q3[p]);
                }
            }
#pragma unroll
            for (int p = 0; p < PIF; p++) {
                const uint32_t k = (g + p) * THREADS + tid;
                const bool short_key = len[p] <= CACHE_SHORT_KEY;
                // a short key owns k[0], k[1] and the length byte of its sector (the rest is the answer), a long one all but the count byte
                auto same_key = [&]() {
                    return q0[p] == k0[p] && q1[p] == k1[p] &&
                           (short_key ? (q3[p] >> 56) == len[p] : (q2[p] == k2[p] && (q3[p] & ~CACHE_LONG_N_MASK) == k3[p]));
                };
                bool hit = false;
                if (keyed && len[p] - 1u < CACHE_MAX_LEN) {
                    hit = same_key();
                    if (!hit && q3[p] != 0) { // the home slot holds another chunk: the probe sequence goes on
#pragma unroll 1
                        for (uint32_t probes = 1; probes < CACHE_MAX_PROBES; probes++) {
                            h[p] = (h[p] + 1) & a.cache.mask;
                            ld_sector256<LD>(&a.cache.slots[h[p]], q0[p], q1[p], q2[p], q3[p]);
                            if (q3[p] == 0) break;
                            if (same_key()) {
                                hit = true;
                                break;
                            }
                        }
                    }
                }
                uint2 rec = make_uint2(0, 0);
                uint32_t n = 0;
                if (hit) {
                    if (short_key && !(q3[p] & CACHE_NOT_INLINE)) { // nine chunks in ten end here
                        n = (uint32_t)q3[p] & 0x7Fu;
                        rec = make_uint2((uint32_t)q2[p], (uint32_t)(q2[p] >> 32));
                    } else {
                        n = short_key ? ((uint32_t)q3[p] & 0x7Fu) : (uint32_t)(q3[p] >> CACHE_LONG_N_SHIFT) & 0xFFu;
                        rec = make_uint2(h[p], REC_REF | (CK_SLOT << 29));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(&a.cache.slots[h[p]].n)); // (the gather reads the value sector)
                    }
                } else if (len[p] != 0) {
                    if (len[p] > ENC_SHORT_MAX) {
                        rec = make_uint2(o0[p], REC_REF | (CK_LONG << 29));
                        if (a.scratch_b) {
                            n = a.scratch_b[o0[p]]; // encoded by k_encode_long
                        } else {                    // optimistic launch: report it, the host runs the long path and repeats
                            const uint32_t q = atomicAdd(a.n_long, 1u);
                            if (q < a.long_cap) a.long_list[q] = (uint32_t)(c0 + k);
                        }
                    } else if (a.ablate & 2) {
                        n = 2;
                        rec = make_uint2(o0[p] | (len[p] << CACHE_ID_BITS), 0);
                    } else {
                        const uint32_t q = atomicAdd(&sm.n_open, 1u);
                        sm.open_k[q] = (uint16_t)k;
                        rec.y = REC_REF | (CK_PARKED << 29); // (count and place: set_parked)
                    }
                }
                sm.rec[k] = rec;
                sm.cnt[k] = n;
            }
        }
        __syncthreads();
        lap(1);
        // ---- 2. chunks the cache does not hold ------------------------------------------------------------------------
        {
            const uint32_t n_scan = sm.n_open;
            if (n_scan) {
                scan_open_chunks<THREADS>(a, sm, a0, staged, n_scan, true);
                __syncthreads();
            }
        }
        lap(2);
        // ---- 3a. count scan: CPT consecutive chunks per thread -> loc[k], the chunk's offset within the warp's segment -----
        {
            uint32_t c[CPT];
            if constexpr (CPT == 2) {
                const uint2 x = reinterpret_cast<const uint2 *>(sm.cnt)[tid];
                c[0] = x.x;
                c[1] = x.y;
            } else {
#pragma unroll
                for (int q = 0; q < CPT / 4; q++) {
                    const uint4 x = reinterpret_cast<const uint4 *>(sm.cnt)[tid * (CPT / 4) + q];
                    c[4 * q] = x.x;
                    c[4 * q + 1] = x.y;
                    c[4 * q + 2] = x.z;
                    c[4 * q + 3] = x.w;
                }
            }
            uint32_t sum = 0;
#pragma unroll
            for (int i = 0; i < CPT; i++) sum += c[i];
            uint32_t incl = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += v;
            }
            uint32_t run = incl - sum;
#pragma unroll
            for (int i = 0; i < CPT; i++) {
                const uint32_t n = c[i];
                c[i] = run;
                run += n;
            }
            // (loc lies over the boundaries: the scan list was their last reader, and a barrier has passed since)
            if constexpr (CPT == 2) {
                reinterpret_cast<uint2 *>(sm.loc)[tid] = make_uint2(c[0], c[1]);
            } else {
#pragma unroll
                for (int q = 0; q < CPT / 4; q++)
                    reinterpret_cast<uint4 *>(sm.loc)[tid * (CPT / 4) + q] = make_uint4(c[4 * q], c[4 * q + 1], c[4 * q + 2], c[4 * q + 3]);
            }
            if (lane == 31) sm.warp_sum[warp] = incl;
        }
        __syncthreads();
        lap(3);
        // ---- 3b. place in the stream: look-back by warp 0; the ids are gathered in shared memory meanwhile ------------
        uint32_t total = 0, pre[CPT]; // pre[j]: ids of the tile before the segment of this thread's j-th chunk
#pragma unroll
        for (int j = 0; j < CPT; j++) pre[j] = 0;
#pragma unroll
        for (int w = 0; w < NW; w++) {
            const uint32_t v = sm.warp_sum[w];
#pragma unroll
            for (int j = 0; j < CPT; j++)
                if (w < j * SEG_PER_J + (int)(warp / CPT)) pre[j] += v;
            total += v;
        }
        long long t_dbg = a.prof ? clock64() : 0;
        // warp 0 announces the tile's id count, gathers its share like everybody else, and only then walks the predecessors:
        // most of them have announced themselves by then, so the walk rarely has to wait
        uint64_t base0 = 0;
        if (warp == 0 && !(a.ablate & 1)) base0 = lookback_announce(a.status, tile, total, a.stream_base);
        const bool via_smem = total <= (uint32_t)SM::STAGE;
#pragma unroll
        for (int j = 0; j < CPT; j++) pre[j] += sm.loc[j * THREADS + tid]; // = the chunk's place in the tile
        if (via_smem && !(a.ablate & 4)) gather_chunks<THREADS, CPT>(a, sm, pre, sm.stage, ~0ull);
        if (a.prof && (tid == 0 || tid == 32)) { const long long t = clock64(); sm.prof[tid ? 11 : 9] += t - t_dbg; t_dbg = t; }
        if (warp == 0) {
            const uint64_t b = (a.ablate & 1) ? (uint64_t)tile * (TILE * 9 / 4) : lookback_resolve<4>(a.status, tile, total, base0);
            if (lane == 0) sm.base = b;
        }
        if (a.prof && tid == 0) { const long long t = clock64(); sm.prof[8] += t - t_dbg; t_dbg = t; }
        __syncthreads();
        if (a.prof && (tid == 0 || tid == 32)) { const long long t = clock64(); sm.prof[tid ? 12 : 10] += t - t_dbg; t_dbg = t; }
        lap(4);
        const uint64_t base = sm.base;
        if (!via_smem && !(a.ablate & 4)) { // a tile with more ids than the gather buffer holds: every thread stores its own
            gather_chunks<THREADS, CPT>(a, sm, pre, a.out + base, base < a.out_cap ? a.out_cap - base : 0);
            __syncthreads(); // (the gather reads the records and the parking area: they are replaced below)
        }
        // ---- 4. the next tile's loads start now; this tile's ids leave as whole lines --------------------------------
        if (bulk && tid == 0) {
            sm.n_open = 0; // (everybody is past the scan list; the next tile's appends come after the next mbarrier wait)
            sm.park_used = 0;
            fetch_tile_bulk(a, sm, off_parity);
        }
        lap(5);
        if (a.pf_ahead && bulk) {
            // One round of tickets from now some CTA fetches tile `tile + pf_ahead`: its boundaries are at a known place, its
            // text about pf_ahead tiles of this tile's size further on (a window of 16 KiB around the estimate). Asking L2 for
            // them now turns the two dependent HBM round trips of fetch_tile_bulk into L2 round trips.
            const uint64_t ca = a.chunk0 + ((uint64_t)tile + a.pf_ahead) * TILE;
            if (ca < a.chunk1) {
                if (tid < TILE / 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.off + min(ca + tid * 32, a.chunk1)));
                if (tid >= 64 && tid < 64 + 128) {
                    const uint64_t span = (uint64_t)sm.off_span, est = (uint64_t)a0 + span * a.pf_ahead + span / 2;
                    const uint64_t g = (est > 8192 ? est - 8192 : 0) + (uint64_t)(tid - 64) * 128;
                    if (g < a.n_bytes_total) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.bytes + g));
                }
            }
        }
        if (a.out_off) {
#pragma unroll
            for (int j = 0; j < CPT; j++) {
                const uint32_t k = j * THREADS + tid;
                if (k < nc) a.out_off[c0 + k] = base + pre[j];
            }
        }
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
        if (tile == a.n_tiles - 1 && tid == 0) {
            *a.d_n_out = base + total;
            if (a.out_off && a.chunk1 == a.n_chunks) a.out_off[a.n_chunks] = base + total;
        }
        lap(6);
        if (a.prof && tid == 0) sm.prof[7] += 1;
    }
    if (a.prof && tid == 0)
        for (int i = 0; i < ENC_PROF_N; i++) atomicAdd(&a.prof[i], sm.prof[i]);
}

} // namespace mbpe
