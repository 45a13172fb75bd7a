// encode.cu -- encode merge scan (K5) and decode gather (K6) for sm_100a, plus their C ABI.
//
// Replaces (file:line under /root/reference/code/include/Tokenizer.h):
//   internal_internal_encode :325-367   per chunk: left-to-right scan replacing ANY known pair, repeated until
//                                       a pass merges nothing (not rank-ordered BPE)
//   internal_encode + flatten :370-377, :714-717
//   merges_lookup             :74, :833-837 (later duplicate pairs overwrite the id)
//   decode                    :725-751
//
// Data layout in HBM: text bytes (u8), chunk boundaries (u32 offsets, n_chunks+1, batch < 4 GiB), output ids (u32,
// one flat stream in chunk order). The pair lookup is an open-addressed table of 8-byte slots
// {a:21 | b:21 | id:21} (vocab <= 2^21), 2-4x over-provisioned, read through the read-only path: 32k merges =
// 512 KB, resident in L2 and mostly in L1.
//
// k_encode_tiles: tiles of 1024 chunks staged in shared memory; a chunk's ids come from the chunk cache (chunk bytes
// -> ids, learned between sub-batches) or, on a miss, from the multi-pass scan itself (one warp per chunk); the flat
// output position comes from a block scan + decoupled look-back over tiles, so the stream is produced in ONE
// pass: every text byte and boundary is read once, every id written once (SURVEY 8(d) B_enc).
// Chunks longer than ENC_SHORT_MAX (encoder "basic": the whole text is one chunk) are encoded first by
// k_encode_long into a scratch stream and spliced in by k_encode_tiles.
// k_decode_tiles: ids -> bytes gather through a packed (length + bytes) vocabulary word, same look-back.
// mbpe_decode_file: .enc file -> text file in blocks (reader thread, device, writer thread).
#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"
#include "lookback.cuh"

namespace mbpe {

constexpr uint32_t ENC_SHORT_MAX = 64;   // bytes; longer chunks take the scratch path
constexpr int ENC_THREADS = 256;
constexpr uint64_t ENC_EMPTY = ~0ull;
constexpr uint32_t ENC_ID_BITS = 21;
constexpr uint32_t ENC_ID_MASK = (1u << ENC_ID_BITS) - 1;

__host__ __device__ __forceinline__ uint64_t enc_key(uint32_t a, uint32_t b) { return ((uint64_t)a << ENC_ID_BITS) | b; }
__host__ __device__ __forceinline__ uint32_t enc_hash(uint64_t k) {
    k *= 0x9E3779B97F4A7C15ull;
    return (uint32_t)(k >> 32);
}

struct EncTable {
    const uint64_t *slots;
    uint32_t mask;
};

__device__ __forceinline__ bool enc_lookup(const EncTable &t, uint32_t a, uint32_t b, uint32_t &id) {
    const uint64_t key = enc_key(a, b);
    uint32_t h = enc_hash(key) & t.mask;
    for (;;) {
        uint64_t s = __ldg(&t.slots[h]);
        if (s == ENC_EMPTY) return false;
        if ((s >> ENC_ID_BITS) == key) {
            id = (uint32_t)s & ENC_ID_MASK;
            return true;
        }
        h = (h + 1) & t.mask;
    }
}

constexpr uint32_t ENC_NONE = 0xFFFFFFFFu;
__device__ __forceinline__ uint32_t enc_lookup_id(const EncTable &t, uint32_t a, uint32_t b) {
    uint32_t id;
    return enc_lookup(t, a, b, id) ? id : ENC_NONE;
}

// one pass of Tokenizer.h:336-359 over t[0..len): returns new length, sets merged
__device__ __forceinline__ uint32_t enc_pass(const EncTable &tab, uint32_t *t, uint32_t len, bool &merged) {
    uint32_t w = 0, i = 0;
    merged = false;
    while (i < len) {
        uint32_t id;
        if (i + 1 < len && enc_lookup(tab, t[i], t[i + 1], id)) {
            t[w++] = id;
            i += 2;
            merged = true;
        } else {
            t[w++] = t[i++];
        }
    }
    return w;
}

// ---------------------------------------------------------------------------------------------------------
// k_encode_long: chunks longer than ENC_SHORT_MAX, one thread each, ping-pong in two scratch streams.
// Result: tokens at scratch_a[off[c] ..], count at scratch_b[off[c]].
// ---------------------------------------------------------------------------------------------------------
__global__ void k_encode_find_long(const uint32_t *off, uint64_t n_chunks, uint32_t *long_list, uint32_t *n_long,
                                   uint32_t cap) {
    for (uint64_t c = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; c < n_chunks; c += (uint64_t)gridDim.x * blockDim.x)
        if (off[c + 1] - off[c] > ENC_SHORT_MAX) {
            uint32_t k = atomicAdd(n_long, 1u);
            if (k < cap) long_list[k] = (uint32_t)c;
        }
}

__global__ void k_encode_long(EncTable tab, const uint8_t *bytes, const uint32_t *off, const uint32_t *long_list,
                              uint32_t n_long, uint32_t *scratch_a, uint32_t *scratch_b) {
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n_long; k += gridDim.x * blockDim.x) {
        uint32_t c = long_list[k], o = off[c], len = off[c + 1] - o;
        uint32_t *t = scratch_a + o;
        for (uint32_t i = 0; i < len; i++) t[i] = bytes[o + i];
        bool merged = true;
        while (merged && len >= 2) len = enc_pass(tab, t, len, merged); // in place: write index never passes read index
        scratch_b[o] = len;
    }
}

// ---------------------------------------------------------------------------------------------------------
// k_encode_tiles
// ---------------------------------------------------------------------------------------------------------
struct CacheSlot;
struct CacheLogEntry;
struct ChunkCache {
    CacheSlot *slots;    // nullptr = cache disabled
    uint32_t mask;       // slots - 1
    CacheLogEntry *log;  // chunks the current sub-batch had to scan
    uint32_t *log_count;
    uint32_t log_cap;
    uint32_t *used;      // occupied slots (learning stops at half full)
    uint32_t *arena;     // ids of entries with more than CACHE_INLINE_IDS ids
    uint32_t *arena_used;
    uint32_t arena_cap;
};

struct EncArgs {
    EncTable tab;
    ChunkCache cache;
    uint64_t chunk0, chunk1;   // this launch covers chunks [chunk0, chunk1) in tiles of ET_CHUNKS
    const unsigned long long *stream_base; // ids produced by earlier sub-batches (device word) = base of tile 0
    const uint8_t *bytes;
    uint64_t n_bytes_total; // bytes readable from `bytes` (vector loads never cross it)
    const uint32_t *off;
    uint64_t n_chunks;
    uint32_t *out;
    uint64_t out_cap;
    unsigned long long *d_n_out;
    unsigned long long *out_off; // optional per-chunk token offsets (n_chunks + 1)
    unsigned long long *status;  // look-back words, one per tile, zeroed
    uint32_t *ticket;            // zeroed
    uint32_t n_tiles;
    const uint32_t *scratch_a;   // long-chunk tokens / counts (may be null)
    const uint32_t *scratch_b;
    uint32_t *overflow;          // set when out_cap is too small
    uint32_t *miss_count;        // chunks that went through the scan (statistics)
    uint32_t ablate;             // MBPE_ENC_ABLATE (profiling only; 1, 2, 4 give WRONG results): 1 no look-back wait, 2 no
                                 // cache probe (every chunk "hits" with two fake ids), 4 no id stores, 16 ids stored by their threads instead of through shared memory
};

// ---------------------------------------------------------------------------------------------------------
// Chunk cache. The multi-pass scan of a chunk is a pure function of its bytes, and text repeats its chunks
// (Zipf): the encoder keeps an open-addressed table  chunk bytes (<= 15) -> ids (<= 3)  in HBM (32-byte slots, one
// sector per probe, L2/L1 resident for the hot words). k_encode_tiles only READS it (read-only path); chunks it
// does not find are encoded by the scan itself and appended to a log, and k_cache_insert adds the log to the table
// between sub-batches. No kernel both reads and writes the table, so there is no publication protocol to get
// wrong. Results are bit-identical with or without the cache (MBPE_ENCODE_CACHE=0 disables it; tests run both).
// ---------------------------------------------------------------------------------------------------------
struct CacheSlot { // 64 bytes = two sectors: key, value
    uint64_t k[4]; // chunk bytes, little endian, zero padded; top byte of k[3] = length (1..31); k[3] == 0: empty
    uint32_t n;    // number of ids (1..31)
    uint32_t v[7]; // n <= 7: the ids; otherwise v[0] = offset of the ids in the arena
};
static_assert(sizeof(CacheSlot) == 64, "two sectors per entry");
struct CacheLogEntry {
    uint64_t k[4];
    uint32_t n;
    uint32_t ids[31];
};
constexpr uint32_t CACHE_MAX_LEN = 31, CACHE_INLINE_IDS = 7;

__device__ __forceinline__ uint32_t cache_hash(uint64_t k0, uint64_t k1, uint64_t k2, uint64_t k3) {
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
        if (*((volatile uint32_t *)cc.used) * 2 > cc.mask) return; // half full: stop learning
        const uint64_t k0 = e.k[0], k1 = e.k[1], k2 = e.k[2], k3 = e.k[3];
        const uint32_t en = e.n;
        uint32_t h = cache_hash(k0, k1, k2, k3) & cc.mask;
        for (;;) {
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
// k_encode_tiles. A tile = ET_CHUNKS consecutive chunks; its boundaries and its text are staged into shared
// memory with coalesced loads. Each thread owns ET_CPT consecutive chunks:
//   1. chunks <= 15 bytes: assemble the 16-byte key from shared memory, probe the cache (read-only path);
//   2. misses go to a tile work list and are encoded by the scan, one chunk per thread, all lanes busy;
//   3. block scan + decoupled look-back give the tile its place in the flat stream; ids are written in order.
// ---------------------------------------------------------------------------------------------------------
constexpr int ET_CAP = 8192;         // staged text bytes per tile (avg chunk 5 B -> 5 KB); bigger tiles read HBM directly
constexpr uint32_t ET_MISS_OUT = 2048; // ids of scanned chunks parked in shared memory until the tile offset is known
constexpr uint32_t META_NONE = 0xFFFFF;
constexpr uint64_t ENC_MAX_SUBBATCH = 1ull << 26; // chunks per launch at most (tile ids and look-back words are 32-bit safe)
constexpr uint32_t ET_WARP_SCAN_MAX = 48; // up to this many misses per tile are scanned one warp per chunk

template <int THREADS, int ET_CPT, int STG>
struct EncSmemT {
    static constexpr int ET_CHUNKS = THREADS * ET_CPT;
    uint32_t off[ET_CHUNKS + 1];
    alignas(16) uint32_t text[ET_CAP / 4 + 8]; // raw bytes, 16-byte aligned window
    uint16_t miss[ET_CHUNKS];      // work list: chunk index within the tile
    uint32_t meta[ET_CHUNKS];      // scanned chunks: start in miss_out (20 bits, META_NONE = not parked) | count << 20
    uint32_t miss_out[ET_MISS_OUT];
    uint32_t stage[STG ? STG : 1];
    uint32_t warp_scratch[THREADS / 32][32];
    uint32_t tile, n_miss, miss_used;
    uint32_t warp_sum[THREADS / 32];
    unsigned long long base;
};

template <class SM>
__device__ __forceinline__ uint8_t tile_byte(const EncArgs &a, const SM &sm, bool staged, uint32_t a0, uint32_t g) {
    return staged ? reinterpret_cast<const uint8_t *>(sm.text)[g - a0] : __ldg(&a.bytes[g]);
}

// scan one chunk (<= ENC_SHORT_MAX bytes) into t[]; returns the id count
template <class SM>
__device__ __forceinline__ uint32_t scan_chunk(const EncArgs &a, const SM &sm, bool staged, uint32_t a0, uint32_t o,
                                               uint32_t len, uint32_t *t) {
    for (uint32_t i = 0; i < len; i++) t[i] = tile_byte(a, sm, staged, a0, o + i);
    bool merged = true;
    while (merged && len >= 2) len = enc_pass(a.tab, t, len, merged);
    return len;
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

template <int THREADS, int ET_CPT, int MIN_CTAS, bool STREAM, int STG>
__global__ void __launch_bounds__(THREADS, MIN_CTAS) k_encode_tiles(const EncArgs a) {
    using EncSmem = EncSmemT<THREADS, ET_CPT, STG>;
    constexpr int ET_CHUNKS = EncSmem::ET_CHUNKS;
    constexpr int ENC_THREADS = THREADS; // shadows the file-level default inside this kernel
    extern __shared__ __align__(16) unsigned char enc_smem_raw[];
    EncSmem &sm = *reinterpret_cast<EncSmem *>(enc_smem_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool use_cache = a.cache.slots != nullptr;
    for (;;) {
        __syncthreads();
        if (tid == 0) {
            sm.tile = atomicAdd(a.ticket, 1u); // tiles start in order: look-back cannot deadlock
            sm.n_miss = 0;
            sm.miss_used = 0;
        }
        __syncthreads();
        const uint32_t tile = sm.tile;
        if (tile >= a.n_tiles) return;
        const uint64_t c0 = a.chunk0 + (uint64_t)tile * ET_CHUNKS;
        const uint32_t nc = (uint32_t)min((uint64_t)ET_CHUNKS, a.chunk1 - c0);
        for (uint32_t i = tid; i <= nc; i += ENC_THREADS) sm.off[i] = STREAM ? __ldcs(&a.off[c0 + i]) : __ldg(&a.off[c0 + i]);
        __syncthreads();
        const uint32_t b0 = sm.off[0], b1 = sm.off[nc];
        const uint32_t a0 = b0 & ~15u; // 16-byte aligned window start (device buffers are 256-byte aligned)
        const bool staged = (b1 - a0) <= ET_CAP;
        if (staged) {
            uint4 *dst = reinterpret_cast<uint4 *>(sm.text);
            for (uint32_t v = tid; a0 + v * 16 < b1; v += ENC_THREADS) {
                const uint32_t g0 = a0 + v * 16;
                uint4 q;
                if (g0 + 16 <= a.n_bytes_total) {
                    q = STREAM ? __ldcs(reinterpret_cast<const uint4 *>(a.bytes + g0)) : __ldg(reinterpret_cast<const uint4 *>(a.bytes + g0));
                } else { // last vector of the buffer
                    uint32_t w[4] = {0, 0, 0, 0};
                    for (uint32_t g = g0; g < a.n_bytes_total; g++)
                        w[(g - g0) >> 2] |= (uint32_t)__ldg(&a.bytes[g]) << (((g - g0) & 3) * 8);
                    q = make_uint4(w[0], w[1], w[2], w[3]);
                }
                dst[v] = q;
            }
            __syncthreads();
        }
        // ---- 1. cache probes -------------------------------------------------------------------------------
        uint32_t cnt[ET_CPT], ids[ET_CPT][3];
        uint32_t state[ET_CPT]; // 0 = ids[] valid (hit with <= 3 ids / empty chunk), 1 = scanned, 2 = long chunk,
                                // 3 = hit with more ids: ids[j][0] = cache slot, ids are fetched again when writing
        // key of chunk [o, o+len), len <= 31: bytes little endian, zero padded, length in the top byte
        auto make_key = [&](uint32_t o, uint32_t len, uint64_t *key) {
            const uint32_t r = o - a0, wi = r >> 2, sh = (r & 3) * 8;
            uint32_t w[9];
#pragma unroll
            for (int q = 0; q < 5; q++) w[q] = sm.text[wi + q];
            uint32_t v[8];
#pragma unroll
            for (int q = 0; q < 4; q++) v[q] = __funnelshift_r(w[q], w[q + 1], sh);
#pragma unroll
            for (int q = 4; q < 8; q++) v[q] = 0;
            if (len > 16) {
#pragma unroll
                for (int q = 5; q < 9; q++) w[q] = sm.text[wi + q];
#pragma unroll
                for (int q = 4; q < 8; q++) v[q] = __funnelshift_r(w[q], w[q + 1], sh);
            }
            key[0] = ((uint64_t)v[1] << 32) | v[0];
            key[1] = ((uint64_t)v[3] << 32) | v[2];
            key[2] = ((uint64_t)v[5] << 32) | v[4];
            key[3] = ((uint64_t)v[7] << 32) | v[6];
#pragma unroll
            for (int q = 0; q < 4; q++) { // zero everything at and after byte `len`
                const int lo = q * 8;
                if ((int)len <= lo)
                    key[q] = 0;
                else if ((int)len < lo + 8)
                    key[q] &= (1ull << ((len - lo) * 8)) - 1;
            }
            key[3] |= (uint64_t)len << 56;
        };
        const bool cacheable_tile = use_cache && staged;
#pragma unroll
        for (int j = 0; j < ET_CPT; j++) {
            const uint32_t k = tid * ET_CPT + j;
            cnt[j] = 0;
            state[j] = 0;
            if (k >= nc) continue;
            const uint32_t o = sm.off[k], len = sm.off[k + 1] - o;
            if (len == 0) continue;
            if (len > ENC_SHORT_MAX) {
                state[j] = 2;
                cnt[j] = a.scratch_b[o]; // encoded by k_encode_long
                continue;
            }
            bool hit = false;
            if (cacheable_tile && len <= CACHE_MAX_LEN) {
                uint64_t key[4];
                make_key(o, len, key);
                uint32_t h = cache_hash(key[0], key[1], key[2], key[3]) & a.cache.mask;
                if (a.ablate & 2) {
                    cnt[j] = 2;
                    ids[j][0] = (uint32_t)key[0];
                    ids[j][1] = h;
                    continue;
                }
                for (;;) {
                    // the whole 64-byte entry at once: three independent 16-byte loads, one round trip
                    const ulonglong2 *sp = reinterpret_cast<const ulonglong2 *>(&a.cache.slots[h]);
                    const ulonglong2 lo = __ldg(sp);
                    const ulonglong2 hi = __ldg(sp + 1);
                    const uint4 tv = __ldg(reinterpret_cast<const uint4 *>(sp + 2)); // n, v[0..2]
                    if (hi.y == 0) break; // empty: not cached
                    if (hi.y == key[3] && hi.x == key[2] && lo.x == key[0] && lo.y == key[1]) {
                        cnt[j] = tv.x;
                        ids[j][0] = tv.y;
                        ids[j][1] = tv.z;
                        ids[j][2] = tv.w;
                        if (tv.x > 3) {
                            state[j] = 3;
                            ids[j][0] = h;
                        }
                        hit = true;
                        break;
                    }
                    h = (h + 1) & a.cache.mask;
                }
            }
            if (!hit) {
                state[j] = 1;
                sm.miss[atomicAdd(&sm.n_miss, 1u)] = (uint16_t)k;
            }
        }
        __syncthreads();
        // ---- 2. scan the misses, ids parked in shared memory until the tile knows its place ------------------
        const uint32_t n_miss = sm.n_miss;
        if (tid == 0 && n_miss) atomicAdd(a.miss_count, n_miss);
        if (n_miss <= ET_WARP_SCAN_MAX) {
            // few misses (warm cache): latency matters -- one WARP per chunk, lanes = positions
            for (uint32_t q = warp; q < n_miss; q += ENC_THREADS / 32) {
                const uint32_t mk = sm.miss[q], o = sm.off[mk], mlen = sm.off[mk + 1] - o;
                if (mlen > 32) continue; // 33..64 bytes: serial scan below
                uint32_t tok = lane < mlen ? tile_byte(a, sm, staged, a0, o + lane) : 0u;
                const uint32_t mn = scan_chunk_warp(a.tab, tok, mlen, sm.warp_scratch[warp]); // lane i < mn: id i in tok
                uint32_t start = 0;
                if (lane == 0) start = atomicAdd(&sm.miss_used, mn);
                start = __shfl_sync(0xffffffffu, start, 0);
                if (start + mn <= ET_MISS_OUT) {
                    if (lane < mn) sm.miss_out[start + lane] = tok;
                } else {
                    start = META_NONE;
                }
                if (lane == 0) sm.meta[mk] = start | (mn << 20);
                if (use_cache && mlen <= CACHE_MAX_LEN) { // teach the cache
                    uint32_t li = 0;
                    if (lane == 0) li = atomicAdd(a.cache.log_count, 1u);
                    li = __shfl_sync(0xffffffffu, li, 0);
                    if (li < a.cache.log_cap) {
                        CacheLogEntry &e = a.cache.log[li];
                        if (lane < mn) e.ids[lane] = tok;
                        if (lane == 0) {
                            uint64_t key[4] = {0, 0, 0, 0};
                            for (uint32_t i = 0; i < mlen; i++)
                                key[i >> 3] |= (uint64_t)tile_byte(a, sm, staged, a0, o + i) << ((i & 7) * 8);
                            key[3] |= (uint64_t)mlen << 56;
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
        }
        for (uint32_t q = tid; q < n_miss; q += ENC_THREADS) {
            // many misses (cold cache): throughput matters -- one THREAD per chunk; also chunks of 33..64 bytes
            const uint32_t mk = sm.miss[q], o = sm.off[mk], mlen = sm.off[mk + 1] - o;
            if (n_miss <= ET_WARP_SCAN_MAX && mlen <= 32) continue;
            uint32_t t[ENC_SHORT_MAX];
            const uint32_t mn = scan_chunk(a, sm, staged, a0, o, mlen, t);
            uint32_t start = atomicAdd(&sm.miss_used, mn);
            if (start + mn <= ET_MISS_OUT) {
                for (uint32_t i = 0; i < mn; i++) sm.miss_out[start + i] = t[i];
            } else {
                start = META_NONE; // no room: the owner scans it again when writing
            }
            sm.meta[mk] = start | (mn << 20);
            if (use_cache && mlen <= CACHE_MAX_LEN) { // teach the cache
                const uint32_t li = atomicAdd(a.cache.log_count, 1u);
                if (li < a.cache.log_cap) {
                    uint64_t key[4] = {0, 0, 0, 0};
                    for (uint32_t i = 0; i < mlen; i++)
                        key[i >> 3] |= (uint64_t)tile_byte(a, sm, staged, a0, o + i) << ((i & 7) * 8);
                    key[3] |= (uint64_t)mlen << 56;
                    CacheLogEntry &e = a.cache.log[li];
                    e.k[0] = key[0];
                    e.k[1] = key[1];
                    e.k[2] = key[2];
                    e.k[3] = key[3];
                    e.n = mn;
                    for (uint32_t i = 0; i < mn; i++) e.ids[i] = t[i];
                }
            }
        }
        __syncthreads();
        uint32_t sum = 0;
#pragma unroll
        for (int j = 0; j < ET_CPT; j++) {
            if (state[j] == 1) cnt[j] = sm.meta[tid * ET_CPT + j] >> 20;
            sum += cnt[j];
        }
        // ---- 3. place in the stream: block exclusive scan + look-back -----------------------------------------
        uint32_t incl = sum;
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        if (lane == 31) sm.warp_sum[warp] = incl;
        __syncthreads();
        uint32_t warp_base = 0, total = 0;
#pragma unroll
        for (int w = 0; w < ENC_THREADS / 32; w++) {
            uint32_t v = sm.warp_sum[w];
            if (w < (int)warp) warp_base += v;
            total += v;
        }
        if (warp == 0) {
            const uint64_t b = (a.ablate & 1) ? (uint64_t)tile * (ET_CHUNKS * 9 / 4)
                                              : lookback_base(a.status, tile, total, a.stream_base);
            if (lane == 0) sm.base = b;
        }
        __syncthreads();
        const uint64_t base = sm.base;
        // ids go to shared memory first and leave for HBM as whole 128-byte lines (a thread's own stores would touch
        // one sector per lane); tiles whose ids do not fit the buffer, or the end of a full output buffer, store directly
        const bool via_smem = STG > 0 && total <= (uint32_t)STG && base + total <= a.out_cap && !(a.ablate & 16);
        uint32_t loc = warp_base + (incl - sum);
        uint32_t *const gout = a.out + base; // only dereferenced below out_cap
#pragma unroll
        for (int j = 0; j < ET_CPT; j++) {
            const uint32_t k = tid * ET_CPT + j;
            if (k >= nc) continue;
            const uint32_t n = cnt[j];
            if (a.out_off) a.out_off[c0 + k] = base + loc;
            if (a.ablate & 4) {
            } else if (via_smem || base + loc + n <= a.out_cap) {
                uint32_t *const dstp = (via_smem ? sm.stage : gout) + loc;
                if (state[j] == 0) {
                    if (n > 0) dstp[0] = ids[j][0];
                    if (n > 1) dstp[1] = ids[j][1];
                    if (n > 2) dstp[2] = ids[j][2];
                } else if (state[j] == 3) {
                    const CacheSlot &cs = a.cache.slots[ids[j][0]];
                    if (n <= CACHE_INLINE_IDS) { // the value sector again: two vector loads, predicated stores, no loop
                        const uint4 lo4 = __ldg(reinterpret_cast<const uint4 *>(&cs.n));      // n, v0, v1, v2
                        const uint4 hi4 = __ldg(reinterpret_cast<const uint4 *>(&cs.v[3]));   // v3 .. v6
                        dstp[0] = lo4.y;
                        dstp[1] = lo4.z;
                        dstp[2] = lo4.w;
                        dstp[3] = hi4.x;
                        if (n > 4) dstp[4] = hi4.y;
                        if (n > 5) dstp[5] = hi4.z;
                        if (n > 6) dstp[6] = hi4.w;
                    } else {
                        const uint32_t *src = a.cache.arena + __ldg(&cs.v[0]);
                        for (uint32_t i = 0; i < n; i++) dstp[i] = __ldg(&src[i]);
                    }
                } else if (state[j] == 1) {
                    const uint32_t start = sm.meta[k] & 0xFFFFF;
                    if (start != META_NONE) {
                        for (uint32_t i = 0; i < n; i++) dstp[i] = sm.miss_out[start + i];
                    } else {
                        uint32_t t[ENC_SHORT_MAX];
                        const uint32_t o = sm.off[k];
                        scan_chunk(a, sm, staged, a0, o, sm.off[k + 1] - o, t);
                        for (uint32_t i = 0; i < n; i++) dstp[i] = t[i];
                    }
                } else {
                    const uint32_t o = sm.off[k];
                    for (uint32_t i = 0; i < n; i++) dstp[i] = a.scratch_a[o + i];
                }
            } else if (n) {
                *a.overflow = 1;
            }
            loc += n;
        }
        if (via_smem && !(a.ablate & 4)) {
            __syncthreads();
            for (uint32_t i = tid; i < total; i += ENC_THREADS) {
                if (STREAM) __stcs(&gout[i], sm.stage[i]);
                else gout[i] = sm.stage[i];
            }
        }
        if (tile == a.n_tiles - 1 && tid == 0) {
            *a.d_n_out = base + total;
            if (a.out_off && a.chunk1 == a.n_chunks) a.out_off[a.n_chunks] = base + total;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// k_decode_tiles: ids -> bytes gather (Tokenizer.h:725-751). vocab entry = (offset, length) into a byte arena;
// special ids are looked up first in a sorted side table; unknown ids contribute nothing.
// ---------------------------------------------------------------------------------------------------------
struct DecArgs {
    const uint32_t *ids;
    uint64_t n_ids;
    const uint32_t *v_off; // vocab_size + 1
    const uint8_t *v_bytes;
    const unsigned long long *v_pack; // per id: length in the low byte (0xFF = longer than 7 bytes), then the bytes
    uint32_t vocab_size;
    const uint32_t *sp_ids; // sorted
    const uint32_t *sp_off; // n_sp + 1 into sp_bytes
    const uint8_t *sp_bytes;
    uint32_t n_sp;
    uint8_t *out;
    uint64_t out_cap; // 0 with out == null: size only
    unsigned long long *d_n_out;
    unsigned long long *status;
    uint32_t *ticket;
    uint32_t n_tiles;
};

// A tile = DEC_THREADS x DEC_IPT consecutive ids. Each thread owns DEC_IPT consecutive ids: coalesced 16-byte id
// loads, (offset, length) from the vocabulary index (128 KB for 32k ids: L1/L2 resident), block scan + look-back for
// the tile's place in the byte stream, then every token's bytes are gathered into shared memory and the tile leaves
// as aligned 4-byte words, 128 bytes per warp store (tokens average 2-3 bytes: per-thread stores would touch one
// sector each). Tiles whose bytes do not fit the staging buffer store directly.
constexpr int DEC_THREADS = 256, DEC_IPT = 8, DEC_IDS = DEC_THREADS * DEC_IPT;
constexpr uint32_t DEC_STAGE = 24576; // bytes staged per tile (avg ~5 KB)

struct DecSmem {
    alignas(16) uint8_t stage[DEC_STAGE + 16];
    uint32_t s_warp[DEC_THREADS / 32];
    uint32_t s_tile;
    unsigned long long s_base;
};

__global__ void __launch_bounds__(DEC_THREADS) k_decode_tiles(const DecArgs a) {
    __shared__ DecSmem sm;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (;;) {
        __syncthreads();
        if (tid == 0) sm.s_tile = atomicAdd(a.ticket, 1u);
        __syncthreads();
        const uint32_t tile = sm.s_tile;
        if (tile >= a.n_tiles) return;
        const uint64_t k0 = (uint64_t)tile * DEC_IDS + (uint64_t)tid * DEC_IPT;
        uint32_t id[DEC_IPT];
        if (k0 + DEC_IPT <= a.n_ids) { // ids is 16-byte aligned (device allocation), k0 a multiple of 8
            const uint4 q0 = __ldcs(reinterpret_cast<const uint4 *>(a.ids + k0));
            const uint4 q1 = __ldcs(reinterpret_cast<const uint4 *>(a.ids + k0 + 4));
            id[0] = q0.x, id[1] = q0.y, id[2] = q0.z, id[3] = q0.w, id[4] = q1.x, id[5] = q1.y, id[6] = q1.z, id[7] = q1.w;
        } else {
#pragma unroll
            for (int j = 0; j < DEC_IPT; j++) id[j] = k0 + j < a.n_ids ? __ldcs(a.ids + k0 + j) : 0xFFFFFFFFu;
        }
        // one 8-byte load per id answers "how long, which bytes" for tokens of up to 7 bytes (nearly all of them);
        // longer tokens and special tokens are resolved again through the index when their bytes are copied
        unsigned long long pk[DEC_IPT]; // low byte: length (0xFF: long / special -> resolve()), then the bytes
        uint32_t len[DEC_IPT], sum = 0;
        auto find_special = [&](uint32_t idv) -> int { // special tokens override the vocabulary (Tokenizer.h:733)
            int lo = 0, hi = (int)a.n_sp - 1;
            while (lo <= hi) {
                const int mid = (lo + hi) >> 1;
                const uint32_t v = __ldg(&a.sp_ids[mid]);
                if (v == idv) return mid;
                if (v < idv)
                    lo = mid + 1;
                else
                    hi = mid - 1;
            }
            return -1;
        };
        constexpr unsigned long long PK_SPECIAL = 1ull << 63; // with low byte 0xFF: bits 8..39 = index of the special token
#pragma unroll
        for (int j = 0; j < DEC_IPT; j++) pk[j] = (a.v_pack && k0 + j < a.n_ids && id[j] < a.vocab_size) ? __ldg(&a.v_pack[id[j]]) : 0xFFull;
#pragma unroll
        for (int j = 0; j < DEC_IPT; j++) {
            len[j] = 0;
            if (k0 + j >= a.n_ids) continue;
            const int sp = a.n_sp ? find_special(id[j]) : -1;
            if (sp >= 0) {
                len[j] = __ldg(&a.sp_off[sp + 1]) - __ldg(&a.sp_off[sp]);
                pk[j] = PK_SPECIAL | ((unsigned long long)(uint32_t)sp << 8) | 0xFF;
            } else if (id[j] >= a.vocab_size) {
                pk[j] = 0; // unknown ids contribute nothing (Tokenizer.h:739-742)
            } else if ((pk[j] & 0xFF) == 0xFF) {
                len[j] = __ldg(&a.v_off[id[j] + 1]) - __ldg(&a.v_off[id[j]]);
            } else {
                len[j] = (uint32_t)(pk[j] & 0xFF);
            }
            sum += len[j];
        }
        uint32_t incl = sum;
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        if (lane == 31) sm.s_warp[warp] = incl;
        __syncthreads();
        uint32_t warp_base = 0, total = 0;
#pragma unroll
        for (int w = 0; w < DEC_THREADS / 32; w++) {
            const uint32_t v = sm.s_warp[w];
            if (w < (int)warp) warp_base += v;
            total += v;
        }
        if (warp == 0) {
            const uint64_t b = lookback_base(a.status, tile, total);
            if (lane == 0) sm.s_base = b;
        }
        __syncthreads();
        const uint64_t base = sm.s_base;
        if (a.out) {
            // the staged image starts at the 4-byte boundary below `base`: stage[pad + i] = byte i of the tile
            const uint32_t pad = (uint32_t)(base & 3);
            const bool via_smem = pad + total <= DEC_STAGE && base + total <= a.out_cap;
            uint32_t loc = warp_base + (incl - sum);
#pragma unroll
            for (int j = 0; j < DEC_IPT; j++) {
                if (len[j] == 0) continue;
                if ((pk[j] & 0xFF) != 0xFF) { // bytes are in the register
                    unsigned long long v = pk[j] >> 8;
                    if (via_smem) {
                        for (uint32_t i = 0; i < len[j]; i++, v >>= 8) sm.stage[pad + loc + i] = (uint8_t)v;
                    } else if (base + loc + len[j] <= a.out_cap) {
                        for (uint32_t i = 0; i < len[j]; i++, v >>= 8) a.out[base + loc + i] = (uint8_t)v;
                    }
                } else {
                    const uint8_t *src = (pk[j] & PK_SPECIAL) ? a.sp_bytes + __ldg(&a.sp_off[(uint32_t)(pk[j] >> 8)])
                                                                : a.v_bytes + __ldg(&a.v_off[id[j]]);
                    if (via_smem) {
                        for (uint32_t i = 0; i < len[j]; i++) sm.stage[pad + loc + i] = __ldg(src + i);
                    } else if (base + loc + len[j] <= a.out_cap) {
                        for (uint32_t i = 0; i < len[j]; i++) a.out[base + loc + i] = __ldg(src + i);
                    }
                }
                loc += len[j];
            }
            if (via_smem) {
                __syncthreads();
                const uint64_t w0 = base - pad;                       // 4-byte aligned (out is a device allocation)
                const uint32_t n_words = (pad + total + 3) >> 2;
                const uint32_t *sw = reinterpret_cast<const uint32_t *>(sm.stage);
                for (uint32_t w = tid; w < n_words; w += DEC_THREADS) {
                    const uint32_t lo = w * 4, hi = lo + 4;
                    if (lo >= pad && hi <= pad + total) {
                        __stcs(reinterpret_cast<uint32_t *>(a.out + w0) + w, sw[w]);
                    } else { // first / last word of the tile: shared with the neighbouring tiles, byte stores
                        for (uint32_t i = max(lo, pad); i < min(hi, pad + total); i++) a.out[w0 + i] = sm.stage[i];
                    }
                }
            }
        }
        if (tile == a.n_tiles - 1 && tid == 0) *a.d_n_out = base + total;
    }
}
} // namespace mbpe

using namespace mbpe;

// ---------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------
struct mbpe_encoder {
    int device = 0, sms = 148;
    uint32_t n_merges = 0, vocab_size = 256;
    uint64_t *d_slots = nullptr;
    uint32_t mask = 0;
    // decode tables
    uint32_t *d_voff = nullptr;
    uint8_t *d_vbytes = nullptr;
    unsigned long long *d_vpack = nullptr; // decode: length + up to 7 bytes per id in one word
    uint32_t *d_sp_ids = nullptr, *d_sp_off = nullptr;
    uint8_t *d_sp_bytes = nullptr;
    uint32_t n_sp = 0;
    // scratch (grown on demand)
    unsigned long long *d_status = nullptr;
    uint64_t status_cap = 0;
    uint32_t *d_small = nullptr; // [0] ticket, [1] n_long, [2] overflow
    unsigned long long *d_n_out = nullptr;
    uint32_t *d_long_list = nullptr;
    uint64_t long_cap = 0;
    uint32_t *d_scratch_a = nullptr, *d_scratch_b = nullptr;
    uint64_t scratch_cap = 0;
    uint64_t launches = 0;
    // chunk cache (learned across calls)
    CacheSlot *d_cache = nullptr;
    CacheLogEntry *d_cache_log = nullptr;
    uint32_t *d_cache_arena = nullptr;
    uint32_t cache_slots = 0, cache_log_cap = 0, cache_arena_cap = 0;
    uint32_t *d_cache_ctr = nullptr; // [0] log count, [1] used slots, [2] arena cursor
    uint64_t sub_batch_chunks = 0; // fixed sub-batch size (MBPE_ENCODE_SUBBATCH), 0 = geometric schedule
    uint64_t chunks_seen = 0;      // chunks encoded so far with this handle: how warm the cache is
    int cfg = 0; // kernel shape, see enc_configs
    bool specials_seeded = false; // mbpe_encoder_seed_special_chunks succeeded
    size_t l2_window_max = 0, l2_persist_bytes = 0;
};

namespace mbpe {
struct EncConfig {
    int threads, cpt, ctas;
    void (*kernel)(const EncArgs);
    size_t smem;
};
#define ENC_CFG(T, C, M, P, G) EncConfig{T, C, M, k_encode_tiles<T, C, M, P, G>, sizeof(EncSmemT<T, C, G>)}
// (threads, chunks per thread, CTAs per SM, streaming hints on the one-pass loads/stores, staging words); 0 = default
static const EncConfig enc_configs[] = {ENC_CFG(256, 4, 4, true, 3072),  ENC_CFG(256, 4, 4, true, 4096), ENC_CFG(256, 4, 4, true, 0),
                                        ENC_CFG(256, 4, 4, false, 0),    ENC_CFG(256, 4, 4, false, 4096), ENC_CFG(128, 4, 8, true, 2048),
                                        ENC_CFG(128, 4, 8, true, 0),     ENC_CFG(256, 4, 5, true, 3072)};
constexpr int N_ENC_CONFIGS = sizeof(enc_configs) / sizeof(enc_configs[0]);
} // namespace mbpe

extern "C" int mbpe_encoder_create(const uint32_t *merges, uint32_t n_merges, int device, mbpe_encoder **out) {
    if (!out || (n_merges && !merges)) return set_error(MBPE_E_INVALID, "null argument");
    *out = nullptr;
    if (256ull + n_merges > (1ull << ENC_ID_BITS)) return set_error(MBPE_E_INVALID, "vocab larger than 2^21 ids");
    // vocabulary as load() rebuilds it (Tokenizer.h:844-861); a pair may only name earlier ids
    std::vector<uint32_t> voff(257 + (size_t)n_merges);
    std::vector<uint8_t> vbytes;
    vbytes.reserve(256 + 8ull * n_merges);
    for (uint32_t i = 0; i < 256; i++) {
        voff[i] = i;
        vbytes.push_back((uint8_t)i);
    }
    voff[256] = 256;
    for (uint32_t i = 0; i < n_merges; i++) {
        uint32_t a = merges[2 * i], b = merges[2 * i + 1];
        if (a >= 256 + i || b >= 256 + i) return set_error(MBPE_E_INVALID, "merge names an id that does not exist yet");
        size_t la = voff[a + 1] - voff[a], lb = voff[b + 1] - voff[b];
        if (vbytes.size() + la + lb >= 0xFFFFFFFFull) return set_error(MBPE_E_INVALID, "vocabulary bytes exceed 4 GiB");
        size_t base = vbytes.size();
        vbytes.resize(base + la + lb);
        memcpy(&vbytes[base], &vbytes[voff[a]], la);
        memcpy(&vbytes[base + la], &vbytes[voff[b]], lb);
        voff[256 + i + 1] = (uint32_t)vbytes.size();
    }
    // lookup table, later duplicates overwrite (Tokenizer.h:835)
    uint32_t cap = 1024;
    while (cap < 4ull * n_merges) cap <<= 1;
    std::vector<uint64_t> slots(cap, ENC_EMPTY);
    for (uint32_t i = 0; i < n_merges; i++) {
        uint64_t key = enc_key(merges[2 * i], merges[2 * i + 1]);
        uint32_t h = enc_hash(key) & (cap - 1);
        while (slots[h] != ENC_EMPTY && (slots[h] >> ENC_ID_BITS) != key) h = (h + 1) & (cap - 1);
        slots[h] = (key << ENC_ID_BITS) | (256 + i);
    }
    int rc = use_device(device);
    if (rc) return rc;
    mbpe_encoder *e = new mbpe_encoder();
    e->device = device;
    e->sms = sm_count(device);
    e->n_merges = n_merges;
    e->vocab_size = 256 + n_merges;
    e->mask = cap - 1;
    MB_CUDA(cudaMalloc(&e->d_slots, (uint64_t)cap * 8));
    MB_CUDA(cudaMemcpy(e->d_slots, slots.data(), (uint64_t)cap * 8, cudaMemcpyHostToDevice));
    MB_CUDA(cudaMalloc(&e->d_voff, voff.size() * 4));
    MB_CUDA(cudaMemcpy(e->d_voff, voff.data(), voff.size() * 4, cudaMemcpyHostToDevice));
    MB_CUDA(cudaMalloc(&e->d_vbytes, vbytes.size()));
    MB_CUDA(cudaMemcpy(e->d_vbytes, vbytes.data(), vbytes.size(), cudaMemcpyHostToDevice));
    {
        std::vector<unsigned long long> vpack(256 + (size_t)n_merges);
        for (size_t i = 0; i < vpack.size(); i++) {
            const uint32_t l = voff[i + 1] - voff[i];
            unsigned long long v = 0xFF;
            if (l <= 7) {
                v = l;
                for (uint32_t q = 0; q < l; q++) v |= (unsigned long long)vbytes[voff[i] + q] << (8 * (q + 1));
            }
            vpack[i] = v;
        }
        MB_CUDA(cudaMalloc(&e->d_vpack, vpack.size() * 8));
        MB_CUDA(cudaMemcpy(e->d_vpack, vpack.data(), vpack.size() * 8, cudaMemcpyHostToDevice));
    }
    MB_CUDA(cudaMalloc(&e->d_small, 16));
    MB_CUDA(cudaMalloc(&e->d_n_out, 8));
    MB_CUDA(cudaMalloc(&e->d_sp_ids, 4));
    MB_CUDA(cudaMalloc(&e->d_sp_off, 8));
    MB_CUDA(cudaMalloc(&e->d_sp_bytes, 1));
    const char *cache_env = getenv("MBPE_ENCODE_CACHE"); // "0" disables; otherwise log2 of the slot count (default 22 = 256 MB)
    int cache_log2 = cache_env && *cache_env ? atoi(cache_env) : 22;
    if (cache_log2 >= 10 && cache_log2 <= 26) {
        e->cache_slots = 1u << cache_log2;
        e->cache_log_cap = 1u << 19;
        e->cache_arena_cap = 1u << 22;
        MB_CUDA(cudaMalloc(&e->d_cache, (uint64_t)e->cache_slots * sizeof(CacheSlot)));
        MB_CUDA(cudaMemset(e->d_cache, 0, (uint64_t)e->cache_slots * sizeof(CacheSlot)));
        MB_CUDA(cudaMalloc(&e->d_cache_log, (uint64_t)e->cache_log_cap * sizeof(CacheLogEntry)));
        MB_CUDA(cudaMalloc(&e->d_cache_arena, (uint64_t)e->cache_arena_cap * 4));
        MB_CUDA(cudaMalloc(&e->d_cache_ctr, 16));
        MB_CUDA(cudaMemset(e->d_cache_ctr, 0, 16));
    }
    const char *sb_env = getenv("MBPE_ENCODE_SUBBATCH");
    if (sb_env && *sb_env) e->sub_batch_chunks = std::max<uint64_t>(4096, strtoull(sb_env, nullptr, 10)) / 4096 * 4096;
    if (!getenv("MBPE_NO_L2_PERSIST")) {
        int max_persist = 0, max_window = 0;
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, device);
        cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, device);
        if (max_persist > 0 && max_window > 0 &&
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)max_persist) == cudaSuccess) {
            e->l2_window_max = (size_t)max_window;
            e->l2_persist_bytes = (size_t)max_persist;
        } else {
            cudaGetLastError();
        }
    }
    const char *cfg_env = getenv("MBPE_ENC_CFG");
    e->cfg = cfg_env && *cfg_env ? std::min(std::max(atoi(cfg_env), 0), N_ENC_CONFIGS - 1) : 0;
    for (int i = 0; i < N_ENC_CONFIGS; i++)
        MB_CUDA(cudaFuncSetAttribute(enc_configs[i].kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)enc_configs[i].smem));
    *out = e;
    return MBPE_OK;
}

extern "C" void mbpe_encoder_destroy(mbpe_encoder *e) {
    if (!e) return;
    cudaSetDevice(e->device);
    void *ps[] = {e->d_slots, e->d_voff, e->d_vbytes, e->d_vpack, e->d_sp_ids, e->d_sp_off, e->d_sp_bytes, e->d_status, e->d_small,
                  e->d_n_out, e->d_long_list, e->d_scratch_a, e->d_scratch_b, e->d_cache, e->d_cache_log, e->d_cache_ctr, e->d_cache_arena};
    for (void *p : ps) cudaFree(p);
    delete e;
}

extern "C" int mbpe_encoder_set_specials(mbpe_encoder *e, const uint32_t *ids, const uint8_t *bytes,
                                         const uint64_t *off, uint32_t n) {
    if (!e || (n && (!ids || !bytes || !off))) return set_error(MBPE_E_INVALID, "null argument");
    int rc = use_device(e->device);
    if (rc) return rc;
    // sorted by id; for a duplicated id the LAST one wins, as unordered_map assignment does (Tokenizer.h:484)
    std::vector<uint32_t> order(n);
    for (uint32_t i = 0; i < n; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return ids[x] < ids[y]; });
    std::vector<uint32_t> sid, soff{0};
    std::vector<uint8_t> sb;
    for (uint32_t k = 0; k < n; k++) {
        uint32_t i = order[k];
        if (k + 1 < n && ids[order[k + 1]] == ids[i]) continue;
        sid.push_back(ids[i]);
        sb.insert(sb.end(), bytes + off[i], bytes + off[i + 1]);
        soff.push_back((uint32_t)sb.size());
    }
    cudaFree(e->d_sp_ids);
    cudaFree(e->d_sp_off);
    cudaFree(e->d_sp_bytes);
    e->n_sp = (uint32_t)sid.size();
    MB_CUDA(cudaMalloc(&e->d_sp_ids, std::max<size_t>(sid.size(), 1) * 4));
    MB_CUDA(cudaMalloc(&e->d_sp_off, soff.size() * 4));
    MB_CUDA(cudaMalloc(&e->d_sp_bytes, std::max<size_t>(sb.size(), 1)));
    MB_CUDA(cudaMemcpy(e->d_sp_ids, sid.data(), sid.size() * 4, cudaMemcpyHostToDevice));
    MB_CUDA(cudaMemcpy(e->d_sp_off, soff.data(), soff.size() * 4, cudaMemcpyHostToDevice));
    MB_CUDA(cudaMemcpy(e->d_sp_bytes, sb.data(), sb.size(), cudaMemcpyHostToDevice));
    return MBPE_OK;
}

// Special tokens on the encode side (Tokenizer.h:605-650, :667-671): a special token is a chunk of its own that becomes
// one ready-made id. The device front end marks those chunks (mbpe_pretok_split_device_parts); here their bytes are put
// into the chunk cache as "these bytes -> this id", so the tile kernel needs no special case. An ordinary chunk can never
// carry the bytes of a special token (the splitter would have cut it out). The cache is emptied first: it may have
// learned those bytes as ordinary text before. MBPE_E_UNSUPPORTED: no cache, or a token longer than the cache's keys.
extern "C" int mbpe_encoder_seed_special_chunks(mbpe_encoder *e, const uint32_t *ids, const uint8_t *bytes, const uint64_t *off,
                                                uint32_t n) {
    if (!e || (n && (!ids || !bytes || !off))) return set_error(MBPE_E_INVALID, "null argument");
    int rc = use_device(e->device);
    if (rc) return rc;
    e->specials_seeded = false;
    if (!e->d_cache) return set_error(MBPE_E_UNSUPPORTED, "the chunk cache is disabled");
    std::vector<CacheLogEntry> log(n);
    for (uint32_t i = 0; i < n; i++) {
        const uint64_t len = off[i + 1] - off[i];
        if (len == 0 || len > CACHE_MAX_LEN) return set_error(MBPE_E_UNSUPPORTED, "special token longer than 31 bytes");
        if (n > e->cache_log_cap) return set_error(MBPE_E_UNSUPPORTED, "too many special tokens");
        CacheLogEntry &le = log[i];
        memset(&le, 0, sizeof le);
        for (uint64_t q = 0; q < len; q++) le.k[q >> 3] |= (uint64_t)bytes[off[i] + q] << ((q & 7) * 8);
        le.k[3] |= len << 56;
        le.n = 1;
        le.ids[0] = ids[i];
    }
    MB_CUDA(cudaMemset(e->d_cache, 0, (uint64_t)e->cache_slots * sizeof(CacheSlot)));
    MB_CUDA(cudaMemset(e->d_cache_ctr, 0, 16));
    e->chunks_seen = 0;
    if (n) {
        MB_CUDA(cudaMemcpy(e->d_cache_log, log.data(), n * sizeof(CacheLogEntry), cudaMemcpyHostToDevice));
        const uint32_t cnt = n;
        MB_CUDA(cudaMemcpy(e->d_cache_ctr, &cnt, 4, cudaMemcpyHostToDevice));
        ChunkCache cc{e->d_cache, e->cache_slots - 1, e->d_cache_log, e->d_cache_ctr, e->cache_log_cap, e->d_cache_ctr + 1,
                      e->d_cache_arena, e->d_cache_ctr + 2, e->cache_arena_cap};
        k_cache_insert<<<1, 256>>>(cc);
        MB_CUDA(cudaGetLastError());
        MB_CUDA(cudaDeviceSynchronize());
    }
    e->specials_seeded = true;
    return MBPE_OK;
}

static int ensure_status(mbpe_encoder *e, uint64_t n_tiles) {
    if (n_tiles > e->status_cap) {
        cudaFree(e->d_status);
        e->status_cap = n_tiles + n_tiles / 4 + 64;
        MB_CUDA(cudaMalloc(&e->d_status, e->status_cap * 8));
    }
    return MBPE_OK;
}

extern "C" int mbpe_encode_reserve(mbpe_encoder *e, uint64_t n_bytes, uint64_t n_chunks) {
    if (!e) return set_error(MBPE_E_INVALID, "null argument");
    (void)n_bytes;
    int rc = use_device(e->device);
    if (rc) return rc;
    const uint64_t tile_chunks = (uint64_t)enc_configs[e->cfg].threads * enc_configs[e->cfg].cpt;
    const uint64_t max_sb = e->sub_batch_chunks ? e->sub_batch_chunks : ENC_MAX_SUBBATCH;
    if ((rc = ensure_status(e, (std::min(n_chunks, max_sb) + tile_chunks - 1) / tile_chunks + 1))) return rc;
    if (e->long_cap == 0) {
        e->long_cap = 1 << 16;
        MB_CUDA(cudaMalloc(&e->d_long_list, e->long_cap * 4));
    }
    return MBPE_OK;
}

// d_out_off: optional device u64[n_chunks + 1]
static int encode_device_impl(mbpe_encoder *e, const uint8_t *d_bytes, uint64_t n_bytes, const uint32_t *d_off,
                              uint64_t n_chunks, uint32_t *d_out, uint64_t out_cap, uint64_t *d_n_out,
                              unsigned long long *d_out_off, cudaStream_t st) {
    if (n_bytes >= (1ull << 32)) return set_error(MBPE_E_INVALID, "a device batch must be < 4 GiB of text");
    int rc = mbpe_encode_reserve(e, n_bytes, n_chunks);
    if (rc) return rc;
    MB_CUDA(cudaMemsetAsync(e->d_small, 0, 16, st));
    MB_CUDA(cudaMemsetAsync(d_n_out, 0, 8, st));
    if (n_chunks == 0) {
        if (d_out_off) MB_CUDA(cudaMemsetAsync(d_out_off, 0, 8, st));
        return MBPE_OK;
    }
    EncTable tab{e->d_slots, e->mask};
    // long chunks (rare: > 64 bytes) are found on the device; the count comes back to size the scratch path
    unsigned grid = (unsigned)std::min<uint64_t>((n_chunks + 255) / 256, (uint64_t)e->sms * 8);
    k_encode_find_long<<<grid, 256, 0, st>>>(d_off, n_chunks, e->d_long_list, e->d_small + 1, (uint32_t)e->long_cap);
    e->launches++;
    uint32_t n_long = 0;
    MB_CUDA(cudaMemcpyAsync(&n_long, e->d_small + 1, 4, cudaMemcpyDeviceToHost, st));
    MB_CUDA(cudaStreamSynchronize(st));
    if (n_long > e->long_cap) { // list overflowed: grow and redo
        cudaFree(e->d_long_list);
        e->long_cap = n_long + 1024;
        MB_CUDA(cudaMalloc(&e->d_long_list, e->long_cap * 4));
        MB_CUDA(cudaMemsetAsync(e->d_small + 1, 0, 4, st));
        k_encode_find_long<<<grid, 256, 0, st>>>(d_off, n_chunks, e->d_long_list, e->d_small + 1, (uint32_t)e->long_cap);
        e->launches++;
    }
    if (n_long) {
        if (n_bytes + 1 > e->scratch_cap) {
            cudaFree(e->d_scratch_a);
            cudaFree(e->d_scratch_b);
            e->scratch_cap = n_bytes + 1;
            MB_CUDA(cudaMalloc(&e->d_scratch_a, e->scratch_cap * 4));
            MB_CUDA(cudaMalloc(&e->d_scratch_b, e->scratch_cap * 4));
        }
        k_encode_long<<<(n_long + 63) / 64, 64, 0, st>>>(tab, d_bytes, d_off, e->d_long_list, n_long, e->d_scratch_a,
                                                         e->d_scratch_b);
        e->launches++;
    }
    EncArgs a{};
    a.tab = tab;
    a.cache = ChunkCache{e->d_cache, e->cache_slots ? e->cache_slots - 1 : 0, e->d_cache_log, e->d_cache_ctr,
                         e->cache_log_cap, e->d_cache_ctr ? e->d_cache_ctr + 1 : nullptr, e->d_cache_arena,
                         e->d_cache_ctr ? e->d_cache_ctr + 2 : nullptr, e->cache_arena_cap};
    a.bytes = d_bytes;
    a.n_bytes_total = n_bytes;
    a.off = d_off;
    a.n_chunks = n_chunks;
    a.out = d_out;
    a.out_cap = out_cap;
    a.d_n_out = (unsigned long long *)d_n_out;
    a.stream_base = (const unsigned long long *)d_n_out;
    a.out_off = d_out_off;
    a.status = e->d_status;
    a.ticket = e->d_small;
    a.scratch_a = e->d_scratch_a;
    a.scratch_b = e->d_scratch_b;
    a.overflow = e->d_small + 2;
    a.miss_count = e->d_small + 3;
    if (const char *ab = getenv("MBPE_ENC_ABLATE")) a.ablate = (uint32_t)atoi(ab);
    // Sub-batches of whole tiles: the cache learns from one sub-batch before the next one starts, and the ids of
    // sub-batch i+1 continue the stream where sub-batch i ended (*d_n_out).
    const EncConfig &kc = enc_configs[e->cfg];
    const uint64_t ET_CHUNKS = (uint64_t)kc.threads * kc.cpt;
    // Sub-batch schedule: the cache learns between sub-batches, so a cold encoder starts with small ones (1 M chunks)
    // and doubles; a warm one (chunks_seen) goes straight to large sub-batches, which amortise the launch tail.
    uint64_t sb = e->sub_batch_chunks ? e->sub_batch_chunks
                                      : std::min<uint64_t>(ENC_MAX_SUBBATCH, std::max<uint64_t>(1u << 20, e->chunks_seen));
    sb = (sb + ET_CHUNKS - 1) / ET_CHUNKS * ET_CHUNKS;
    for (uint64_t cb = 0; cb < n_chunks;) {
        a.chunk0 = cb;
        a.chunk1 = std::min(n_chunks, cb + sb);
        cb = a.chunk1;
        e->chunks_seen += a.chunk1 - a.chunk0;
        if (!e->sub_batch_chunks) sb = std::min<uint64_t>(ENC_MAX_SUBBATCH, sb * 2);
        const uint64_t n_tiles = (a.chunk1 - a.chunk0 + ET_CHUNKS - 1) / ET_CHUNKS;
        a.n_tiles = (uint32_t)n_tiles;
        MB_CUDA(cudaMemsetAsync(e->d_status, 0, n_tiles * 8, st));
        MB_CUDA(cudaMemsetAsync(e->d_small, 0, 4, st)); // ticket
        if (a.cache.slots) {
            k_cache_reset_log<<<1, 1, 0, st>>>(a.cache);
            e->launches++;
        }
        unsigned g2 = (unsigned)std::min<uint64_t>(n_tiles, (uint64_t)e->sms * kc.ctas);
        cudaLaunchConfig_t lc{};
        lc.gridDim = dim3(g2);
        lc.blockDim = dim3(kc.threads);
        lc.dynamicSmemBytes = kc.smem;
        lc.stream = st;
        cudaLaunchAttribute lattr[1];
        lc.attrs = lattr;
        lc.numAttrs = 0;
        if (a.cache.slots && e->l2_window_max) {
            // the text, boundaries and ids stream through L2 once; the randomly probed chunk cache is what must stay
            const size_t bytes = std::min((size_t)e->cache_slots * sizeof(CacheSlot), e->l2_window_max);
            lattr[0].id = cudaLaunchAttributeAccessPolicyWindow;
            lattr[0].val.accessPolicyWindow.base_ptr = e->d_cache;
            lattr[0].val.accessPolicyWindow.num_bytes = bytes;
            lattr[0].val.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)e->l2_persist_bytes / (double)bytes);
            lattr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            lattr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            lc.numAttrs = 1;
        }
        MB_CUDA(cudaLaunchKernelEx(&lc, kc.kernel, a));
        e->launches++;
        if (a.cache.slots) {
            k_cache_insert<<<e->sms * 2, 256, 0, st>>>(a.cache);
            e->launches++;
        }
    }
    MB_CUDA(cudaGetLastError());
    if (getenv("MBPE_DEBUG")) {
        uint32_t small[4];
        uint32_t ctr[2] = {0, 0};
        MB_CUDA(cudaMemcpyAsync(small, e->d_small, 16, cudaMemcpyDeviceToHost, st));
        if (e->d_cache_ctr) MB_CUDA(cudaMemcpyAsync(ctr, e->d_cache_ctr, 8, cudaMemcpyDeviceToHost, st));
        MB_CUDA(cudaStreamSynchronize(st));
        fprintf(stderr, "[mbpe] encode: %llu chunks, %u scanned (cache misses), %u long, cache entries %u of %u\n",
                (unsigned long long)n_chunks, small[3], small[1], ctr[1], e->cache_slots);
    }
    return MBPE_OK;
}

extern "C" int mbpe_encode_device(mbpe_encoder *e, const uint8_t *d_bytes, uint64_t n_bytes,
                                  const uint32_t *d_chunk_off32, uint64_t n_chunks, uint32_t *d_out_tokens,
                                  uint64_t out_cap, uint64_t *d_n_out, void *stream) {
    if (!e || !d_chunk_off32 || !d_n_out || (n_bytes && !d_bytes)) return set_error(MBPE_E_INVALID, "null argument");
    int rc = use_device(e->device);
    if (rc) return rc;
    return encode_device_impl(e, d_bytes, n_bytes, d_chunk_off32, n_chunks, d_out_tokens, out_cap, d_n_out, nullptr,
                              (cudaStream_t)stream);
}

extern "C" int mbpe_encode(mbpe_encoder *e, const uint8_t *bytes, uint64_t n_bytes, const uint64_t *chunk_off,
                           uint64_t n_chunks, uint32_t *out_tokens, uint64_t out_cap, uint64_t *n_out,
                           uint64_t *out_off) {
    if (!e || !chunk_off || !n_out || (n_bytes && !bytes)) return set_error(MBPE_E_INVALID, "null argument");
    if (chunk_off[0] != 0 || chunk_off[n_chunks] != n_bytes)
        return set_error(MBPE_E_INVALID, "chunk_off must start at 0 and end at n_bytes");
    int rc = use_device(e->device);
    if (rc) return rc;
    *n_out = 0;
    // batches of whole chunks, < 2 GiB of text each, so offsets fit u32 on the device
    const uint64_t BATCH = 1ull << 31;
    uint64_t c0 = 0, produced = 0;
    std::vector<uint32_t> off32;
    uint8_t *d_bytes = nullptr;
    uint32_t *d_off = nullptr, *d_out = nullptr;
    unsigned long long *d_out_off = nullptr;
    uint64_t cap_bytes = 0, cap_chunks = 0;
    int status = MBPE_OK;
    auto cleanup = [&]() {
        cudaFree(d_bytes);
        cudaFree(d_off);
        cudaFree(d_out);
        cudaFree(d_out_off);
    };
    if (out_off) out_off[0] = 0;
    while (c0 < n_chunks) {
        uint64_t b0 = chunk_off[c0], c1 = c0;
        while (c1 < n_chunks && chunk_off[c1 + 1] - b0 <= BATCH) c1++;
        if (c1 == c0) {
            if (chunk_off[c0 + 1] - b0 >= (1ull << 32)) {
                cleanup();
                return set_error(MBPE_E_INVALID, "a single chunk of 4 GiB or more is not supported");
            }
            c1 = c0 + 1;
        }
        uint64_t nb = chunk_off[c1] - b0, nc = c1 - c0;
        off32.resize(nc + 1);
        for (uint64_t i = 0; i <= nc; i++) {
            if (i && chunk_off[c0 + i] < chunk_off[c0 + i - 1]) {
                cleanup();
                return set_error(MBPE_E_INVALID, "chunk_off not monotonic");
            }
            off32[i] = (uint32_t)(chunk_off[c0 + i] - b0);
        }
        if (nb > cap_bytes || nc > cap_chunks) {
            cleanup();
            cap_bytes = std::max(nb, cap_bytes);
            cap_chunks = std::max(nc, cap_chunks);
            MB_CUDA(cudaMalloc(&d_bytes, std::max<uint64_t>(cap_bytes, 1)));
            MB_CUDA(cudaMalloc(&d_off, (cap_chunks + 1) * 4));
            MB_CUDA(cudaMalloc(&d_out, std::max<uint64_t>(cap_bytes, 1) * 4));
            if (out_off) MB_CUDA(cudaMalloc(&d_out_off, (cap_chunks + 1) * 8));
        }
        MB_CUDA(cudaMemcpy(d_bytes, bytes + b0, nb, cudaMemcpyHostToDevice));
        MB_CUDA(cudaMemcpy(d_off, off32.data(), (nc + 1) * 4, cudaMemcpyHostToDevice));
        rc = encode_device_impl(e, d_bytes, nb, d_off, nc, d_out, nb, (uint64_t *)e->d_n_out, d_out_off, nullptr);
        if (rc) {
            cleanup();
            return rc;
        }
        uint64_t n = 0;
        MB_CUDA(cudaMemcpy(&n, e->d_n_out, 8, cudaMemcpyDeviceToHost));
        if (out_tokens && produced + n <= out_cap)
            MB_CUDA(cudaMemcpy(out_tokens + produced, d_out, n * 4, cudaMemcpyDeviceToHost));
        else
            status = MBPE_E_CAPACITY;
        if (out_off) {
            MB_CUDA(cudaMemcpy(out_off + c0, d_out_off, (nc + 1) * 8, cudaMemcpyDeviceToHost));
            if (produced)
                for (uint64_t i = 0; i <= nc; i++) out_off[c0 + i] += produced;
        }
        produced += n;
        c0 = c1;
    }
    cleanup();
    *n_out = produced;
    if (status == MBPE_E_CAPACITY) return set_error(status, "out_tokens too small; *n_out holds the needed count");
    return MBPE_OK;
}

static int decode_launch(mbpe_encoder *e, const uint32_t *d_ids, uint64_t n_ids, uint8_t *d_out, uint64_t out_cap,
                         unsigned long long *d_n_out, cudaStream_t st) {
    const uint64_t n_tiles = (n_ids + DEC_IDS - 1) / DEC_IDS;
    if (n_tiles >= 0xFFFFFFFFull) return set_error(MBPE_E_INVALID, "too many ids in one call");
    int rc = ensure_status(e, n_tiles);
    if (rc) return rc;
    DecArgs a{};
    a.ids = d_ids;
    a.n_ids = n_ids;
    a.v_off = e->d_voff;
    a.v_bytes = e->d_vbytes;
    a.v_pack = getenv("MBPE_DEC_NOPACK") ? nullptr : e->d_vpack;
    a.vocab_size = e->vocab_size;
    a.sp_ids = e->d_sp_ids;
    a.sp_off = e->d_sp_off;
    a.sp_bytes = e->d_sp_bytes;
    a.n_sp = e->n_sp;
    a.d_n_out = d_n_out;
    a.status = e->d_status;
    a.ticket = e->d_small;
    a.n_tiles = (uint32_t)n_tiles;
    a.out = d_out;
    a.out_cap = d_out ? out_cap : 0;
    MB_CUDA(cudaMemsetAsync(e->d_small, 0, 16, st));
    MB_CUDA(cudaMemsetAsync(e->d_status, 0, n_tiles * 8, st));
    const unsigned grid = (unsigned)std::min<uint64_t>(n_tiles, (uint64_t)e->sms * 6);
    k_decode_tiles<<<grid, DEC_THREADS, 0, st>>>(a);
    e->launches++;
    MB_CUDA(cudaGetLastError());
    return MBPE_OK;
}

// resident ids -> resident bytes in one pass. *d_n_out receives the decoded size; bytes that do not fit out_cap are
// dropped (compare the two to detect it). d_out == NULL: size only. d_ids must be 16-byte aligned.
extern "C" int mbpe_decode_device(mbpe_encoder *e, const uint32_t *d_ids, uint64_t n_ids, uint8_t *d_out, uint64_t out_cap,
                                  uint64_t *d_n_out, void *stream) {
    if (!e || !d_n_out || (n_ids && !d_ids)) return set_error(MBPE_E_INVALID, "null argument");
    if (((uintptr_t)d_ids & 15) != 0) return set_error(MBPE_E_INVALID, "d_ids must be 16-byte aligned");
    int rc = use_device(e->device);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    MB_CUDA(cudaMemsetAsync(d_n_out, 0, 8, st));
    if (n_ids == 0) return MBPE_OK;
    return decode_launch(e, d_ids, n_ids, d_out, out_cap, (unsigned long long *)d_n_out, st);
}

extern "C" int mbpe_decode(mbpe_encoder *e, const uint32_t *ids, uint64_t n_ids, uint8_t *out, uint64_t out_cap,
                           uint64_t *n_out) {
    if (!e || !n_out || (n_ids && !ids)) return set_error(MBPE_E_INVALID, "null argument");
    int rc = use_device(e->device);
    if (rc) return rc;
    *n_out = 0;
    if (n_ids == 0) return MBPE_OK;
    uint32_t *d_ids = nullptr;
    uint8_t *d_out = nullptr;
    MB_CUDA(cudaMalloc(&d_ids, n_ids * 4));
    cudaError_t ce = cudaMemcpy(d_ids, ids, n_ids * 4, cudaMemcpyHostToDevice);
    // pass 1 sizes the output, pass 2 writes it (the caller may also stop after pass 1 with out == NULL)
    if (ce == cudaSuccess) rc = decode_launch(e, d_ids, n_ids, nullptr, 0, e->d_n_out, nullptr);
    if (ce == cudaSuccess && rc == MBPE_OK) ce = cudaMemcpy(n_out, e->d_n_out, 8, cudaMemcpyDeviceToHost);
    if (ce == cudaSuccess && rc == MBPE_OK && out) {
        if (*n_out > out_cap) {
            rc = set_error(MBPE_E_CAPACITY, "out too small; *n_out holds the needed size");
        } else if ((ce = cudaMalloc(&d_out, std::max<uint64_t>(*n_out, 1))) == cudaSuccess) {
            rc = decode_launch(e, d_ids, n_ids, d_out, *n_out, e->d_n_out, nullptr);
            if (rc == MBPE_OK) ce = cudaMemcpy(out, d_out, *n_out, cudaMemcpyDeviceToHost);
        }
    }
    cudaFree(d_ids);
    cudaFree(d_out);
    if (rc == MBPE_OK && ce != cudaSuccess) rc = cuda_fail(ce, "decode", __FILE__, __LINE__);
    return rc;
}

// ---------------------------------------------------------------------------------------------------------
// .enc file -> text file in blocks: a reader thread fills pinned id blocks, the calling thread uploads
// a block, sizes and gathers it on the device and downloads the bytes into a pinned block, a writer thread drains
// those into the output file. Blocks of ids are independent, so there is no boundary problem.
// ---------------------------------------------------------------------------------------------------------
namespace {
struct BlockGate {
    std::mutex mu;
    std::condition_variable cv;
    long long ready = -1;
    bool failed = false;
    void publish(long long k) {
        {
            std::lock_guard<std::mutex> lk(mu);
            ready = k;
        }
        cv.notify_all();
    }
    void fail() {
        {
            std::lock_guard<std::mutex> lk(mu);
            failed = true;
        }
        cv.notify_all();
    }
    bool wait_for(long long k) {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return ready >= k || failed; });
        return ready >= k;
    }
};
} // namespace

extern "C" int mbpe_decode_file(mbpe_encoder *e, const char *in_path, const char *out_path, uint64_t *n_ids_out,
                                uint64_t *n_bytes_out) {
    if (!e || !in_path || !out_path) return set_error(MBPE_E_INVALID, "null argument");
    int rc = use_device(e->device);
    if (rc) return rc;
    FILE *fin = fopen(in_path, "rb");
    if (!fin) return set_error(MBPE_E_IO, std::string("cannot open ") + in_path);
    FILE *fout = fopen(out_path, "wb");
    if (!fout) {
        fclose(fin);
        return set_error(MBPE_E_IO, std::string("cannot open ") + out_path);
    }
    constexpr uint64_t BLK = 4u << 20; // ids per block (16 MiB)
    constexpr int NB = 3;
    uint64_t out_cap = BLK * 8; // bytes per block; grows when a block decodes to more
    uint32_t *h_in[NB] = {nullptr, nullptr, nullptr}, *d_ids = nullptr;
    uint8_t *h_out[NB] = {nullptr, nullptr, nullptr}, *d_out = nullptr;
    unsigned long long *d_n = nullptr;
    bool ok = cudaMalloc(&d_ids, BLK * 4) == cudaSuccess && cudaMalloc(&d_out, out_cap) == cudaSuccess && cudaMalloc(&d_n, 8) == cudaSuccess;
    for (int i = 0; i < NB && ok; i++) ok = cudaMallocHost(&h_in[i], BLK * 4) == cudaSuccess && cudaMallocHost(&h_out[i], out_cap) == cudaSuccess;
    auto release = [&]() {
        for (int i = 0; i < NB; i++) {
            cudaFreeHost(h_in[i]);
            cudaFreeHost(h_out[i]);
        }
        cudaFree(d_ids);
        cudaFree(d_out);
        cudaFree(d_n);
        fclose(fin);
        fclose(fout);
    };
    if (!ok) {
        cudaGetLastError();
        release();
        return set_error(MBPE_E_CUDA, "out of (pinned) memory");
    }
    std::mutex meta_mu;
    std::vector<uint64_t> ids_in_block, bytes_in_block;
    std::vector<char> last_block;
    BlockGate read_gate, in_free, write_gate, out_free;
    in_free.publish(NB - 1);
    out_free.publish(NB - 1);
    std::thread reader([&]() {
        for (long long k = 0;; k++) {
            if (!in_free.wait_for(k)) return;
            const size_t got = fread(h_in[k % NB], 4, BLK, fin); // a trailing partial word is dropped (minbpe-cc.cpp:79)
            {
                std::lock_guard<std::mutex> lk(meta_mu);
                ids_in_block.push_back(got);
                last_block.push_back(got < BLK);
            }
            read_gate.publish(k);
            if (got < BLK) return;
        }
    });
    bool write_failed = false;
    long long n_blocks_total = -1;
    std::mutex total_mu;
    std::thread writer([&]() {
        for (long long k = 0;; k++) {
            {
                std::lock_guard<std::mutex> lk(total_mu);
                if (n_blocks_total >= 0 && k >= n_blocks_total) return;
            }
            if (!write_gate.wait_for(k)) return;
            uint64_t n;
            {
                std::lock_guard<std::mutex> lk(meta_mu);
                n = bytes_in_block[k];
            }
            if (fwrite(h_out[k % NB], 1, n, fout) != n) {
                write_failed = true;
                out_free.fail();
                return;
            }
            out_free.publish(k + NB);
        }
    });
    uint64_t total_ids = 0, total_bytes = 0;
    cudaError_t ce = cudaSuccess;
    for (long long k = 0;; k++) {
        if (!read_gate.wait_for(k)) {
            rc = set_error(MBPE_E_IO, "read failed");
            break;
        }
        uint64_t n;
        bool is_last;
        {
            std::lock_guard<std::mutex> lk(meta_mu);
            n = ids_in_block[k];
            is_last = last_block[k];
        }
        if (!out_free.wait_for(k)) {
            rc = set_error(MBPE_E_IO, std::string("write failed: ") + out_path);
            break;
        }
        unsigned long long n_bytes = 0;
        if (n) {
            if ((ce = cudaMemcpy(d_ids, h_in[k % NB], n * 4, cudaMemcpyHostToDevice)) != cudaSuccess) break;
            if ((rc = decode_launch(e, d_ids, n, d_out, out_cap, d_n, nullptr))) break;
            if ((ce = cudaMemcpy(&n_bytes, d_n, 8, cudaMemcpyDeviceToHost)) != cudaSuccess) break;
            if (n_bytes > out_cap) { // a vocabulary of very long tokens: not worth growing pinned buffers mid-stream
                rc = set_error(MBPE_E_UNSUPPORTED, "a block decodes to more than 8 bytes per id: use the whole-file path");
                break;
            }
            if ((ce = cudaMemcpy(h_out[k % NB], d_out, n_bytes, cudaMemcpyDeviceToHost)) != cudaSuccess) break;
        }
        in_free.publish(k + NB);
        {
            std::lock_guard<std::mutex> lk(meta_mu);
            bytes_in_block.push_back(n_bytes);
        }
        total_ids += n;
        total_bytes += n_bytes;
        if (is_last) {
            std::lock_guard<std::mutex> lk(total_mu);
            n_blocks_total = k + 1;
        }
        write_gate.publish(k);
        if (is_last) break;
    }
    if (rc != MBPE_OK || ce != cudaSuccess) {
        in_free.fail();
        write_gate.fail();
    }
    reader.join();
    writer.join();
    if (rc == MBPE_OK && ce != cudaSuccess) rc = cuda_fail(ce, "decode_file", __FILE__, __LINE__);
    if (rc == MBPE_OK && (write_failed || fflush(fout) != 0)) rc = set_error(MBPE_E_IO, std::string("write failed: ") + out_path);
    release();
    if (rc == MBPE_OK) {
        if (n_ids_out) *n_ids_out = total_ids;
        if (n_bytes_out) *n_bytes_out = total_bytes;
    }
    return rc;
}
