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
#include <atomic>
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

} // namespace mbpe

#include "encode_tiles.cuh"
#include "encode_hot.cuh"

namespace mbpe {
// ---------------------------------------------------------------------------------------------------------
// Long chunks (> ENC_SHORT_MAX bytes; encoder "basic": the whole text is one chunk), one thread each, in place in a
// scratch stream. The tile kernel reports them (optimistic first launch); k_encode_find_long only runs when that
// report overflowed. Result: tokens at scratch_a[off[c] ..], count at scratch_b[off[c]].
// ---------------------------------------------------------------------------------------------------------
__global__ void k_encode_find_long(const uint32_t *off, uint64_t n_chunks, uint32_t *long_list, uint32_t *n_long,
                                   uint32_t cap) {
    for (uint64_t c = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; c < n_chunks; c += (uint64_t)gridDim.x * blockDim.x)
        if (off[c + 1] - off[c] > ENC_SHORT_MAX) {
            uint32_t k = atomicAdd(n_long, 1u);
            if (k < cap) long_list[k] = (uint32_t)c;
        }
}

__global__ void k_encode_long(EncTable tab, EncSpecials sp, const uint8_t *bytes, const uint32_t *off, const uint32_t *long_list,
                              uint32_t n_long, uint32_t *scratch_a, uint32_t *scratch_b) {
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n_long; k += gridDim.x * blockDim.x) {
        uint32_t c = long_list[k], o = off[c], len = off[c + 1] - o;
        uint32_t *t = scratch_a + o;
        const uint32_t sid = special_match(sp, len, [&](uint32_t i) { return __ldg(&bytes[o + i]); });
        if (sid != ENC_NONE) { // a special token longer than ENC_SHORT_MAX bytes (Tokenizer.h:667-671)
            t[0] = sid;
            scratch_b[o] = 1;
            continue;
        }
        for (uint32_t i = 0; i < len; i++) t[i] = bytes[o + i];
        bool merged = true;
        while (merged && len >= 2) len = enc_pass(tab, t, len, merged); // in place: write index never passes read index
        scratch_b[o] = len;
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
    // k_decode_lean: per id, byte 0 = length if <= 7; 0x80 | length if 8..126; 0xFF = look it up (127 bytes or more, or a
    // special token: then bit 63 is set and bits 8..39 hold the index of the special token); bytes 1..7 = the token's first
    // 7 bytes. v_pack2: the token's bytes 7..14.
    const unsigned long long *v_lean, *v_pack2;
    uint32_t pf_ahead; // k_decode_lean: ids of the tile this many tickets ahead are prefetched into L2 (0 = off)
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

// A tile = THREADS x DEC_IPT consecutive ids, one CTA. Each thread owns DEC_IPT consecutive ids: coalesced 16-byte id
// loads; one 8-byte word per id from a packed table answers "how long, which bytes" for tokens of up to 7 bytes (nearly
// all of them). The first TAB_IDS entries of that table live in SHARED memory (ids are handed out in merge order, so the
// low ids are the frequent ones): a random 8-byte gather costs a few shared-memory wavefronts there, against one L1 tag
// look-up per lane on the global path, which is what bounded the round-1 kernel. Block scan of the lengths; then warp 0
// resolves the tile's place in the byte stream by decoupled look-back (128 predecessors per round trip) and takes the NEXT
// tile's ticket WHILE the other warps assemble the tile's bytes in shared memory at tile-local offsets -- each thread
// shifts its tokens into a 64-bit register and stores whole words (the first and the last word of its run, which it shares
// with its neighbours, byte by byte). The tile leaves as aligned 4-byte words (128 bytes per warp store): the
// misalignment of the tile's base is absorbed by a funnel shift over the shared-memory image. Tiles whose bytes do not
// fit the staging buffer store directly.
constexpr int DEC_IPT = 8;

// The CTA's threads work in GROUPS of GROUP threads: every group runs its own stream of tiles (own ticket, own image, own
// named barrier) and all of them share the one table. Measured (1 GiB of text, ms): groups of 1024: 4.05, 512: 3.95,
// 256: 4.87, 128: 5.57 -- every tile pays a fixed price (look-back walk, ticket, two barriers), so small tiles lose.
template <int GROUP>
struct DecGroupSmemT {
    static constexpr uint32_t STAGE = GROUP * DEC_IPT * 3; // bytes staged per tile (average ~2.4 bytes per id)
    alignas(16) uint32_t stage[STAGE / 4 + 4];
    uint32_t s_warp[32];
    uint32_t s_tile[2];
    unsigned long long s_base;
};
template <int THREADS, int TAB_IDS, int GROUP>
struct DecSmemT {
    alignas(16) unsigned long long tab[TAB_IDS ? TAB_IDS : 1];
    DecGroupSmemT<GROUP> grp[THREADS / GROUP];
};
__device__ __forceinline__ void group_barrier(uint32_t id, uint32_t n_threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory");
}

// LANES: the ids of a tile are dealt to the lanes round-robin (warp w, round j, lane l: id w * 256 + j * 32 + l) instead of 8
// consecutive ids per thread: every lane places ONE token per round with predicated byte stores -- no per-thread
// accumulator, no data-dependent loops, all 32 lanes busy (the consecutive layout ran at 6 warp instructions per id with
// 20 of 32 lanes active: ncu in profiles/r2).
template <int THREADS, int TAB_IDS, int MIN_CTAS, int GROUP, bool LANES = false>
__global__ void __launch_bounds__(THREADS, MIN_CTAS) k_decode_tiles(const __grid_constant__ DecArgs a) {
    static_assert(THREADS % GROUP == 0 && GROUP % 32 == 0 && THREADS / GROUP <= 15, "one named barrier per group");
    constexpr int NW = GROUP / 32;
    constexpr uint32_t DEC_IDS = GROUP * DEC_IPT, DEC_STAGE = DecGroupSmemT<GROUP>::STAGE;
    extern __shared__ __align__(16) unsigned char dec_smem_raw[];
    DecSmemT<THREADS, TAB_IDS, GROUP> &cta = *reinterpret_cast<DecSmemT<THREADS, TAB_IDS, GROUP> *>(dec_smem_raw);
    DecGroupSmemT<GROUP> &sm = cta.grp[threadIdx.x / GROUP];
    const uint32_t tid = threadIdx.x % GROUP, lane = tid & 31, warp = tid >> 5; // (within the group)
    const uint32_t bar_id = 1 + threadIdx.x / GROUP;
    uint8_t *const stage8 = reinterpret_cast<uint8_t *>(sm.stage);
    if (TAB_IDS && a.v_pack)
        for (uint32_t i = threadIdx.x; i < (uint32_t)TAB_IDS; i += THREADS) cta.tab[i] = i < a.vocab_size ? __ldg(&a.v_pack[i]) : 0xFFull;
    if (tid == 0) sm.s_tile[0] = atomicAdd(a.ticket, 1u); // tiles start in order: look-back cannot deadlock
    __syncthreads();
    uint32_t nxt[DEC_IPT];
    auto load_ids = [&](uint32_t t, uint32_t *dst) { // this thread's ids of tile t (0xFFFFFFFF past the end)
        if (LANES) {
            const uint64_t k = (uint64_t)t * DEC_IDS + (uint64_t)warp * (32 * DEC_IPT) + lane;
#pragma unroll
            for (int j = 0; j < DEC_IPT; j++) dst[j] = (t < a.n_tiles && k + j * 32 < a.n_ids) ? __ldcs(a.ids + k + j * 32) : 0xFFFFFFFFu;
            return;
        }
        const uint64_t k = (uint64_t)t * DEC_IDS + (uint64_t)tid * DEC_IPT;
        if (t < a.n_tiles && k + DEC_IPT <= a.n_ids) { // ids is 16-byte aligned (device allocation), k a multiple of 8
            const uint4 q0 = __ldcs(reinterpret_cast<const uint4 *>(a.ids + k));
            const uint4 q1 = __ldcs(reinterpret_cast<const uint4 *>(a.ids + k + 4));
            dst[0] = q0.x, dst[1] = q0.y, dst[2] = q0.z, dst[3] = q0.w, dst[4] = q1.x, dst[5] = q1.y, dst[6] = q1.z, dst[7] = q1.w;
        } else {
#pragma unroll
            for (int j = 0; j < DEC_IPT; j++) dst[j] = (t < a.n_tiles && k + j < a.n_ids) ? __ldcs(a.ids + k + j) : 0xFFFFFFFFu;
        }
    };
    for (uint32_t it = 0;; it++) {
        const uint32_t tile = sm.s_tile[it & 1];
        if (tile >= a.n_tiles) return;
        // index of this thread's j-th id: k0 + j * KS
        const uint64_t k0 = (uint64_t)tile * DEC_IDS + (LANES ? (uint64_t)warp * (32 * DEC_IPT) + lane : (uint64_t)tid * DEC_IPT);
        constexpr uint32_t KS = LANES ? 32 : 1;
        uint32_t id[DEC_IPT];
        if (it == 0) load_ids(tile, nxt);
#pragma unroll
        for (int j = 0; j < DEC_IPT; j++) id[j] = nxt[j]; // (loaded while the previous tile was being stored)
        unsigned long long pk[DEC_IPT]; // low byte: length (0xFF: long / special -> resolved through the index), then the bytes
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
        for (int j = 0; j < DEC_IPT; j++) {
            pk[j] = 0xFFull;
            if (a.v_pack && k0 + j * KS < a.n_ids && id[j] < a.vocab_size)
                pk[j] = (TAB_IDS && id[j] < (uint32_t)TAB_IDS) ? cta.tab[id[j]] : __ldg(&a.v_pack[id[j]]);
        }
#pragma unroll
        for (int j = 0; j < DEC_IPT; j++) {
            len[j] = 0;
            if (k0 + j * KS >= a.n_ids) continue;
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
        uint32_t loc[DEC_IPT]; // LANES: offset of token j within the warp's part of the tile (id order = round major)
        if (LANES) {
            // two rounds per scan, 16 bits each (a warp's round total stays below 65536 while every token is shorter than
            // 2048 bytes; a longer special token takes the unpacked scans)
            const bool wide = __any_sync(0xffffffffu, sum >= 2048u);
            uint32_t run = 0;
            if (!wide) {
#pragma unroll
                for (int j = 0; j < DEC_IPT; j += 2) {
                    uint32_t pi = len[j] | (len[j + 1] << 16);
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t v = __shfl_up_sync(0xffffffffu, pi, d);
                        if (lane >= d) pi += v;
                    }
                    const uint32_t tot = __shfl_sync(0xffffffffu, pi, 31);
                    loc[j] = run + (pi & 0xFFFFu) - len[j];
                    run += tot & 0xFFFFu;
                    loc[j + 1] = run + (pi >> 16) - len[j + 1];
                    run += tot >> 16;
                }
            } else {
#pragma unroll
                for (int j = 0; j < DEC_IPT; j++) {
                    uint32_t pi = len[j];
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t v = __shfl_up_sync(0xffffffffu, pi, d);
                        if (lane >= d) pi += v;
                    }
                    loc[j] = run + pi - len[j];
                    run += __shfl_sync(0xffffffffu, pi, 31);
                }
            }
            incl = run; // (every lane holds the warp's total)
        } else {
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += v;
            }
        }
        if (lane == 31) sm.s_warp[warp] = incl;
        group_barrier(bar_id, GROUP);
        uint32_t warp_base, total;
        { // every warp scans the warp sums itself (lane w holds warp w's)
            const uint32_t mine = lane < NW ? sm.s_warp[lane] : 0u;
            uint32_t wi = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= d) wi += v;
            }
            total = __shfl_sync(0xffffffffu, wi, 31);
            warp_base = __shfl_sync(0xffffffffu, wi - mine, warp);
        }
        // warp 0 announces the tile's size, assembles its share like everybody else, and only then walks the predecessors
        uint64_t base0 = 0;
        if (warp == 0) base0 = lookback_announce(a.status, tile, total);
        const bool via_smem = a.out != nullptr && total <= DEC_STAGE;
        const uint32_t loc0 = warp_base + (incl - sum);
        // the bytes of a token that is not in the packed word (longer than 7 bytes, or a special token)
        auto long_src = [&](int j) -> const uint8_t * {
            return (pk[j] & PK_SPECIAL) ? a.sp_bytes + __ldg(&a.sp_off[(uint32_t)(pk[j] >> 8)]) : a.v_bytes + __ldg(&a.v_off[id[j]]);
        };
        if (LANES && via_smem) {
#pragma unroll
            for (int j = 0; j < DEC_IPT; j++) {
                uint8_t *const d = stage8 + warp_base + loc[j];
                const bool packed = (pk[j] & 0xFF) != 0xFF;
                const bool any_hi = __any_sync(0xffffffffu, packed && len[j] > 3); // (asked by all lanes, before they diverge)
                if (packed) { // <= 7 bytes, in the register: predicated byte stores, the same code in every lane
                    const unsigned long long v = pk[j] >> 8;
                    const uint32_t l = len[j];
                    if (l > 0) d[0] = (uint8_t)v;
                    if (l > 1) d[1] = (uint8_t)(v >> 8);
                    if (l > 2) d[2] = (uint8_t)(v >> 16);
                    if (any_hi) {
                        if (l > 3) d[3] = (uint8_t)(v >> 24);
                        if (l > 4) d[4] = (uint8_t)(v >> 32);
                        if (l > 5) d[5] = (uint8_t)(v >> 40);
                        if (l > 6) d[6] = (uint8_t)(v >> 48);
                    }
                } else if (len[j]) { // rare: longer than 7 bytes, or a special token
                    const uint8_t *src = long_src(j);
                    for (uint32_t i = 0; i < len[j]; i++) d[i] = __ldg(src + i);
                }
            }
        } else if (via_smem) {
            // This thread's tokens -> stage8[loc0 ..): bytes are shifted into `acc` behind the `fill` (< 4) bytes that precede
            // them in the current word; whole words are stored as words, except the first one of the run if it starts
            // inside a word (`head`: those leading bytes belong to the neighbour), and the tail: byte by byte.
            uint32_t wpos = loc0 & ~3u, fill = loc0 & 3u, head = fill;
            unsigned long long acc = 0;
            auto put_word = [&](uint32_t w, uint32_t upto) { // bytes [head, upto) of the word at wpos
                if (head == 0 && upto == 4) {
                    sm.stage[wpos >> 2] = w;
                } else {
                    for (uint32_t b = head; b < upto; b++) stage8[wpos + b] = (uint8_t)(w >> (8 * b));
                }
                head = 0;
            };
#pragma unroll
            for (int j = 0; j < DEC_IPT; j++) {
                if (len[j] == 0) continue;
                if ((pk[j] & 0xFF) != 0xFF) { // <= 7 bytes, in the register (zero padded)
                    const unsigned long long v = pk[j] >> 8;
                    acc |= v << (8 * fill);
                    uint32_t over = fill ? (uint32_t)(v >> (64 - 8 * fill)) : 0u; // (what the shift pushed out: fill + len <= 10 bytes)
                    fill += len[j];
                    while (fill >= 4) {
                        put_word((uint32_t)acc, 4);
                        acc = (acc >> 32) | ((unsigned long long)over << 32);
                        over = 0;
                        fill -= 4;
                        wpos += 4;
                    }
                } else { // rare: pending bytes out, then the token byte by byte, then start over at its end
                    put_word((uint32_t)acc, fill);
                    const uint8_t *src = long_src(j);
                    const uint32_t at = wpos + fill;
                    for (uint32_t i = 0; i < len[j]; i++) stage8[at + i] = __ldg(src + i);
                    const uint32_t end = at + len[j];
                    wpos = end & ~3u;
                    fill = head = end & 3u;
                    acc = 0;
                }
            }
            put_word((uint32_t)acc, fill);
        }
        if (warp == 0) {
            const uint64_t b = lookback_resolve<4>(a.status, tile, total, base0);
            if (lane == 0) {
                sm.s_base = b;
                sm.s_tile[(it + 1) & 1] = atomicAdd(a.ticket, 1u); // (this tile's size is published: successors do not wait for us)
            }
        }
        group_barrier(bar_id, GROUP);
        const uint64_t base = sm.s_base;
        load_ids(sm.s_tile[(it + 1) & 1], nxt); // the next tile's ids arrive while this one is stored
        if (a.out) {
            if (via_smem && base + total <= a.out_cap) {
                // global word k of the tile = bytes [4k - pad, 4k - pad + 4) of the image; whole words by funnel shift
                const uint32_t pad = (uint32_t)(base & 3), sh = ((4 - pad) & 3) * 8;
                const uint64_t w0 = base - pad; // 4-byte aligned (out is a device allocation)
                const uint32_t n_words = (pad + total + 3) >> 2;
                uint32_t *const gw = reinterpret_cast<uint32_t *>(a.out + w0);
                for (uint32_t k = tid; k < n_words; k += GROUP) {
                    const int lo = (int)(4 * k) - (int)pad; // tile-local offset of the word's first byte
                    if (lo >= 0 && (uint32_t)lo + 4 <= total) {
                        const uint32_t j = (uint32_t)lo >> 2;
                        const uint32_t v = pad ? __funnelshift_r(sm.stage[j], sm.stage[j + 1], sh) : sm.stage[j];
                        __stcs(gw + k, v);
                    } else { // first / last word of the tile: shared with the neighbouring tiles, byte stores
                        for (int i = lo < 0 ? 0 : lo; i < lo + 4 && (uint32_t)i < total; i++) a.out[base + i] = stage8[i];
                    }
                }
            } else if (via_smem) { // the end of a buffer that is too small: what fits, byte by byte
                for (uint32_t i = tid; i < total; i += GROUP)
                    if (base + i < a.out_cap) a.out[base + i] = stage8[i];
            } else if (LANES) { // a tile with more bytes than the image holds: every lane stores its tokens, byte by byte
#pragma unroll
                for (int j = 0; j < DEC_IPT; j++) {
                    const uint64_t at = base + warp_base + loc[j];
                    if (len[j] == 0 || at + len[j] > a.out_cap) continue;
                    if ((pk[j] & 0xFF) != 0xFF) {
                        unsigned long long v = pk[j] >> 8;
                        for (uint32_t i = 0; i < len[j]; i++, v >>= 8) a.out[at + i] = (uint8_t)v;
                    } else {
                        const uint8_t *src = long_src(j);
                        for (uint32_t i = 0; i < len[j]; i++) a.out[at + i] = __ldg(src + i);
                    }
                }
            } else { // a tile with more bytes than the image holds: every thread stores its own, byte by byte
                const uint64_t at = base + loc0;
                const uint64_t room = at < a.out_cap ? a.out_cap - at : 0;
                uint8_t *const dst = a.out + at;
                uint32_t loc = 0;
#pragma unroll
                for (int j = 0; j < DEC_IPT; j++) {
                    if (len[j] == 0) continue;
                    if (loc + len[j] <= room) {
                        if ((pk[j] & 0xFF) != 0xFF) {
                            unsigned long long v = pk[j] >> 8;
                            for (uint32_t i = 0; i < len[j]; i++, v >>= 8) dst[loc + i] = (uint8_t)v;
                        } else {
                            const uint8_t *src = long_src(j);
                            for (uint32_t i = 0; i < len[j]; i++) dst[loc + i] = __ldg(src + i);
                        }
                    }
                    loc += len[j];
                }
            }
        }
        if (tile == a.n_tiles - 1 && tid == 0) *a.d_n_out = base + total;
        // (the image is rewritten only after the next tile's first barrier, which every thread reaches after its stores)
    }
}

// ---------------------------------------------------------------------------------------------------------
// k_decode_lean: the same tile scheme as k_decode_tiles (ticket, block scan, announce / walk, staged image, aligned word
// stores), with the per-id work cut to what ncu shows is needed -- k_decode_tiles runs 6 warp instructions per id at 20 of
// 32 lanes. One id per lane and round (ids dealt round-robin within a warp: 32 consecutive ids per round, coalesced 4-byte
// loads); ONE table word answers length and first 7 bytes for every id below the vocabulary size (special tokens inside
// that range are patched into the table by mbpe_encoder_set_specials, so nobody searches for them; ids beyond the
// vocabulary take the out-of-line path); no per-id index checks (ids past the end of the stream are loaded as 0xFFFFFFFF =
// unknown = nothing); tokens of 8..15 bytes take their second word from a second table instead of a byte loop; bytes leave
// as predicated byte stores, the same instructions in every lane.
// ---------------------------------------------------------------------------------------------------------
constexpr unsigned long long PKL_SPECIAL = 1ull << 63;

// ids beyond the vocabulary: a special token (Tokenizer.h:733-736) or unknown (nothing, :739-742); -> table word
__device__ __noinline__ unsigned long long dec_lean_outside(const DecArgs &a, uint32_t idv) {
    int lo = 0, hi = (int)a.n_sp - 1;
    while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        const uint32_t v = __ldg(&a.sp_ids[mid]);
        if (v == idv) return PKL_SPECIAL | ((unsigned long long)(uint32_t)mid << 8) | 0xFFull;
        if (v < idv)
            lo = mid + 1;
        else
            hi = mid - 1;
    }
    return 0ull;
}
// a token the table does not hold whole: its length and where its bytes are
__device__ __forceinline__ const uint8_t *dec_lean_src(const DecArgs &a, unsigned long long pk, uint32_t idv, uint32_t &len) {
    if (pk & PKL_SPECIAL) {
        const uint32_t sp = (uint32_t)(pk >> 8);
        const uint32_t o = __ldg(&a.sp_off[sp]);
        len = __ldg(&a.sp_off[sp + 1]) - o;
        return a.sp_bytes + o;
    }
    const uint32_t o = __ldg(&a.v_off[idv]);
    len = __ldg(&a.v_off[idv + 1]) - o;
    return a.v_bytes + o;
}

// (Tried: warp 0 of a group as a control warp that carries no ids and walks the predecessors while the others place their
// tokens -- 3.4 / 3.6 ms against 2.9 / 3.0: the walk starts before the predecessors have announced themselves and the next
// ticket is taken earlier, i.e. held idle longer.)
template <int THREADS, int TAB_IDS, int GROUP>
__global__ void __launch_bounds__(THREADS, 1) k_decode_lean(const __grid_constant__ DecArgs a) {
    static_assert(THREADS % GROUP == 0 && GROUP % 32 == 0 && THREADS / GROUP <= 15 && TAB_IDS > 0, "one named barrier per group");
    constexpr int NW = GROUP / 32;
    constexpr uint32_t DEC_IDS = GROUP * DEC_IPT, DEC_STAGE = DecGroupSmemT<GROUP>::STAGE;
    extern __shared__ __align__(16) unsigned char dec_smem_raw[];
    DecSmemT<THREADS, TAB_IDS, GROUP> &cta = *reinterpret_cast<DecSmemT<THREADS, TAB_IDS, GROUP> *>(dec_smem_raw);
    DecGroupSmemT<GROUP> &sm = cta.grp[threadIdx.x / GROUP];
    const uint32_t tid = threadIdx.x % GROUP, lane = tid & 31, warp = tid >> 5; // (within the group)
    const uint32_t bar_id = 1 + threadIdx.x / GROUP;
    uint8_t *const stage8 = reinterpret_cast<uint8_t *>(sm.stage);
    for (uint32_t i = threadIdx.x; i < (uint32_t)TAB_IDS; i += THREADS) cta.tab[i] = i < a.vocab_size ? __ldg(&a.v_lean[i]) : 0ull;
    if (tid == 0) sm.s_tile[0] = atomicAdd(a.ticket, 1u); // tiles start in order: look-back cannot deadlock
    __syncthreads();
    const uint32_t vocab = a.vocab_size, tab_n = min((uint32_t)TAB_IDS, vocab); // ids the shared-memory table answers (vocab >= 256)
    uint32_t nxt[DEC_IPT];
    // round j, lane l: id (warp * 8 + j) * 32 + l of tile t; returns the mask of the rounds that hold an id at all (any u32 may
    // be the id of a special token, so "no id here" cannot be a value)
    auto load_ids = [&](uint32_t t, uint32_t *dst) -> uint32_t {
        const uint64_t k = (uint64_t)t * DEC_IDS + (uint64_t)warp * (32 * DEC_IPT) + lane;
        if (t < a.n_tiles && (uint64_t)(t + 1) * DEC_IDS <= a.n_ids) {
#pragma unroll
            for (int j = 0; j < DEC_IPT; j++) dst[j] = __ldcs(a.ids + k + j * 32);
            return 0xFFu;
        }
        uint32_t m = 0;
#pragma unroll
        for (int j = 0; j < DEC_IPT; j++) {
            const bool ok = t < a.n_tiles && k + j * 32 < a.n_ids;
            dst[j] = ok ? __ldcs(a.ids + k + j * 32) : 0u;
            m |= ok ? 1u << j : 0u;
        }
        return m;
    };
    uint32_t nxt_mask = 0;
    for (uint32_t it = 0;; it++) {
        const uint32_t tile = sm.s_tile[it & 1];
        if (tile >= a.n_tiles) return;
        if (it == 0) nxt_mask = load_ids(tile, nxt);
        const uint32_t have = nxt_mask;
        uint32_t id[DEC_IPT];
        unsigned long long pk[DEC_IPT];
        uint32_t len[DEC_IPT];
        bool odd = false; // some token of this lane is not answered by the table word alone
#pragma unroll
        for (int j = 0; j < DEC_IPT; j++) {
            id[j] = nxt[j]; // (loaded while the previous tile was being stored)
            pk[j] = cta.tab[min(id[j], tab_n - 1)];
            if (id[j] >= tab_n) pk[j] = id[j] < vocab ? __ldg(&a.v_lean[id[j]]) : dec_lean_outside(a, id[j]);
        }
        if (have != 0xFFu) { // the stream's last tile only
#pragma unroll
            for (int j = 0; j < DEC_IPT; j++)
                if (!((have >> j) & 1u)) pk[j] = 0ull;
        }
#pragma unroll
        for (int j = 0; j < DEC_IPT; j++) {
            const uint32_t b0 = (uint32_t)pk[j] & 0xFFu;
            len[j] = b0 & 0x7Fu;
            odd = odd || b0 == 0xFFu;
            // (bytes 7..14 of a longer token are fetched when it is placed, one round after the other: ask L1 for them now)
            if (b0 > 0x87u && b0 != 0xFFu) asm volatile("prefetch.global.L1 [%0];" ::"l"(&a.v_pack2[id[j]]));
        }
        if (__any_sync(0xffffffffu, odd)) { // 127 bytes or more, or a special token: the length comes from the index
#pragma unroll
            for (int j = 0; j < DEC_IPT; j++)
                if (((uint32_t)pk[j] & 0xFFu) == 0xFFu) dec_lean_src(a, pk[j], id[j], len[j]);
        }
        // ---- place of every token within the warp's part of the tile (id order = round major) ---------------------------
        uint32_t loc[DEC_IPT], run = 0;
        {
            uint32_t mx = 0;
#pragma unroll
            for (int j = 0; j < DEC_IPT; j++) mx = max(mx, len[j]);
            if (!__any_sync(0xffffffffu, mx >= 2048u)) { // two rounds per scan, 16 bits each
#pragma unroll
                for (int j = 0; j < DEC_IPT; j += 2) {
                    uint32_t pi = len[j] | (len[j + 1] << 16);
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t v = __shfl_up_sync(0xffffffffu, pi, d);
                        if (lane >= d) pi += v;
                    }
                    const uint32_t tot = __shfl_sync(0xffffffffu, pi, 31);
                    loc[j] = run + (pi & 0xFFFFu) - len[j];
                    run += tot & 0xFFFFu;
                    loc[j + 1] = run + (pi >> 16) - len[j + 1];
                    run += tot >> 16;
                }
            } else {
#pragma unroll
                for (int j = 0; j < DEC_IPT; j++) {
                    uint32_t pi = len[j];
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t v = __shfl_up_sync(0xffffffffu, pi, d);
                        if (lane >= d) pi += v;
                    }
                    loc[j] = run + pi - len[j];
                    run += __shfl_sync(0xffffffffu, pi, 31);
                }
            }
        }
        if (lane == 31) sm.s_warp[warp] = run;
        group_barrier(bar_id, GROUP);
        uint32_t warp_base, total;
        { // every warp scans the warp sums itself (lane w holds warp w's)
            const uint32_t mine = lane < NW ? sm.s_warp[lane] : 0u;
            uint32_t wi = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= d) wi += v;
            }
            total = __shfl_sync(0xffffffffu, wi, 31);
            warp_base = __shfl_sync(0xffffffffu, wi - mine, warp);
        }
        uint64_t base0 = 0;
        if (warp == 0) base0 = lookback_announce(a.status, tile, total);
        const bool via_smem = a.out != nullptr && total <= DEC_STAGE;
        if (via_smem) {
#pragma unroll
            for (int j = 0; j < DEC_IPT; j++) {
                uint8_t *const d = stage8 + warp_base + loc[j];
                const uint32_t lo = (uint32_t)(pk[j] >> 8), hi = (uint32_t)(pk[j] >> 40), l = len[j];
                const uint32_t b0 = (uint32_t)pk[j] & 0xFFu;
                const bool plain = b0 != 0xFFu;          // the first min(l, 7) bytes are in the table word
                const bool tail = plain && l > 7;          // bytes 7.. come from the second table (8..15) / the index (more)
                const bool any_tail = __any_sync(0xffffffffu, tail || !plain);
                if (plain) { // (some lane of 32 nearly always has more than 3 bytes: no vote to skip bytes 3..6)
                    if (l > 0) d[0] = (uint8_t)lo;
                    if (l > 1) d[1] = (uint8_t)(lo >> 8);
                    if (l > 2) d[2] = (uint8_t)(lo >> 16);
                    if (l > 3) d[3] = (uint8_t)(lo >> 24);
                    if (l > 4) d[4] = (uint8_t)hi;
                    if (l > 5) d[5] = (uint8_t)(hi >> 8);
                    if (l > 6) d[6] = (uint8_t)(hi >> 16);
                }
                if (any_tail) {
                    if (tail) {
                        const unsigned long long w2 = __ldg(&a.v_pack2[id[j]]);
                        const uint32_t lo2 = (uint32_t)w2, hi2 = (uint32_t)(w2 >> 32);
                        d[7] = (uint8_t)lo2;
                        if (l > 8) d[8] = (uint8_t)(lo2 >> 8);
                        if (l > 9) d[9] = (uint8_t)(lo2 >> 16);
                        if (l > 10) d[10] = (uint8_t)(lo2 >> 24);
                        if (l > 11) {
                            d[11] = (uint8_t)hi2;
                            if (l > 12) d[12] = (uint8_t)(hi2 >> 8);
                            if (l > 13) d[13] = (uint8_t)(hi2 >> 16);
                            if (l > 14) d[14] = (uint8_t)(hi2 >> 24);
                            if (l > 15) {
                                const uint8_t *src = a.v_bytes + __ldg(&a.v_off[id[j]]);
#pragma unroll 1
                                for (uint32_t i = 15; i < l; i++) d[i] = __ldg(src + i);
                            }
                        }
                    } else if (!plain && l) {
                        uint32_t ll;
                        const uint8_t *src = dec_lean_src(a, pk[j], id[j], ll);
#pragma unroll 1
                        for (uint32_t i = 0; i < l; i++) d[i] = __ldg(src + i);
                    }
                }
            }
        }
        if (warp == 0) {
            const uint64_t b = lookback_resolve<4>(a.status, tile, total, base0);
            if (lane == 0) {
                sm.s_base = b;
                sm.s_tile[(it + 1) & 1] = atomicAdd(a.ticket, 1u); // (this tile's size is published: successors do not wait for us)
            }
        }
        group_barrier(bar_id, GROUP);
        const uint64_t base = sm.s_base;
        nxt_mask = load_ids(sm.s_tile[(it + 1) & 1], nxt); // the next tile's ids arrive while this one is stored
        // ... and the tile one round of tickets further on (this group's, or a neighbour's) is asked into L2: one line per thread
        if (a.pf_ahead && tid < DEC_IDS / 32) {
            const uint64_t k = ((uint64_t)sm.s_tile[(it + 1) & 1] + a.pf_ahead) * DEC_IDS + (uint64_t)tid * 32;
            if (k < a.n_ids) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.ids + k));
        }
        if (a.out) {
            if (via_smem && base + total <= a.out_cap && (reinterpret_cast<uintptr_t>(a.out) & 15) == 0) {
                // global 16-byte vector v of the tile = bytes [16v - pad, 16v - pad + 16) of the image: five image words, four
                // funnel shifts (the misalignment of the tile's base is the same for every vector), one 16-byte store
                const uint32_t pad = (uint32_t)(base & 15), sh = ((16 - pad) & 3) * 8;
                const uint32_t n_vec = (pad + total + 15) >> 4;
                uint4 *const gv = reinterpret_cast<uint4 *>(a.out + (base - pad));
                for (uint32_t v = tid; v < n_vec; v += GROUP) {
                    const int lo = (int)(16 * v) - (int)pad; // tile-local offset of the vector's first byte
                    if (lo >= 0 && (uint32_t)lo + 16 <= total) {
                        const uint32_t j = (uint32_t)lo >> 2;
                        const uint32_t w0 = sm.stage[j], w1 = sm.stage[j + 1], w2 = sm.stage[j + 2], w3 = sm.stage[j + 3], w4 = sm.stage[j + 4];
                        __stcs(gv + v, make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh),
                                                  __funnelshift_r(w3, w4, sh)));
                    } else { // first / last vector of the tile: shared with the neighbouring tiles, byte stores
                        for (int i = lo < 0 ? 0 : lo; i < lo + 16 && (uint32_t)i < total; i++) a.out[base + i] = stage8[i];
                    }
                }
            } else if (via_smem && base + total <= a.out_cap) {
                // global word k of the tile = bytes [4k - pad, 4k - pad + 4) of the image; whole words by funnel shift
                const uint32_t pad = (uint32_t)(base & 3), sh = ((4 - pad) & 3) * 8;
                const uint64_t w0 = base - pad; // 4-byte aligned (out is a device allocation)
                const uint32_t n_words = (pad + total + 3) >> 2;
                uint32_t *const gw = reinterpret_cast<uint32_t *>(a.out + w0);
                for (uint32_t k = tid; k < n_words; k += GROUP) {
                    const int lo = (int)(4 * k) - (int)pad; // tile-local offset of the word's first byte
                    if (lo >= 0 && (uint32_t)lo + 4 <= total) {
                        const uint32_t j = (uint32_t)lo >> 2;
                        const uint32_t v = pad ? __funnelshift_r(sm.stage[j], sm.stage[j + 1], sh) : sm.stage[j];
                        __stcs(gw + k, v);
                    } else { // first / last word of the tile: shared with the neighbouring tiles, byte stores
                        for (int i = lo < 0 ? 0 : lo; i < lo + 4 && (uint32_t)i < total; i++) a.out[base + i] = stage8[i];
                    }
                }
            } else if (via_smem) { // the end of a buffer that is too small: what fits, byte by byte
                for (uint32_t i = tid; i < total; i += GROUP)
                    if (base + i < a.out_cap) a.out[base + i] = stage8[i];
            } else { // a tile with more bytes than the image holds: every lane stores its tokens, byte by byte
#pragma unroll
                for (int j = 0; j < DEC_IPT; j++) {
                    const uint64_t at = base + warp_base + loc[j];
                    if (len[j] == 0 || at + len[j] > a.out_cap) continue;
                    uint32_t ll;
                    const uint8_t *src = dec_lean_src(a, (pk[j] & PKL_SPECIAL) ? pk[j] : 0ull, id[j], ll);
                    for (uint32_t i = 0; i < len[j]; i++) a.out[at + i] = __ldg(src + i);
                }
            }
        }
        if (tile == a.n_tiles - 1 && tid == 0) *a.d_n_out = base + total;
        // (the image is rewritten only after the next tile's first barrier, which every thread reaches after its stores)
    }
}

struct DecConfig {
    int threads, group, ctas, tile_ids; // tile_ids: ids of one tile of one group
    void (*kernel)(const DecArgs);
    size_t smem;
};
#define DEC_CFG(T, I, M, G) DecConfig{T, G, M, G * DEC_IPT, k_decode_tiles<T, I, M, G>, sizeof(DecSmemT<T, I, G>)}
#define DEC_LANES(T, I, M, G) DecConfig{T, G, M, G * DEC_IPT, k_decode_tiles<T, I, M, G, true>, sizeof(DecSmemT<T, I, G>)}
#define DEC_LEAN(T, I, G) DecConfig{T, G, 1, G * DEC_IPT, k_decode_lean<T, I, G>, sizeof(DecSmemT<T, I, G>)}
// (threads, packed-vocabulary entries in shared memory, CTAs per SM, threads per tile group); 0 = default, the others for
// A/B runs (MBPE_DEC_CFG)
static const DecConfig dec_configs[] = {DEC_CFG(1024, 24576, 1, 512), DEC_CFG(1024, 24576, 1, 1024), DEC_CFG(1024, 24576, 1, 256),
                                        DEC_CFG(256, 0, 4, 256),       DEC_CFG(512, 12288, 2, 256),
                                        DEC_LANES(1024, 24576, 1, 512), DEC_LANES(1024, 24576, 1, 1024), DEC_LANES(1024, 24576, 1, 256),
                                        DEC_LEAN(1024, 24576, 512),     DEC_LEAN(1024, 24576, 1024),      DEC_LEAN(1024, 24576, 256),
                                        DEC_LEAN(1024, 20480, 512),     DEC_LEAN(1024, 16384, 512),       DEC_LEAN(1024, 12288, 512),
                                        DEC_LEAN(1024, 8192, 512),      DEC_LEAN(1024, 16384, 256)};
// 0-4 k_decode_tiles, 5-7 its lane-per-token variant, 8-15 k_decode_lean. Default: 16384 table words in shared memory and four
// groups of 256 threads -- a smaller table leaves L1 to the ids and to the global half of the table (2.79 -> 2.70 ms per GiB
// from 24576 to 20480 words, flat down to 8192), and with it four groups beat two (2.67).
constexpr int DEC_DEFAULT_CFG = 15;
constexpr int N_DEC_CONFIGS = sizeof(dec_configs) / sizeof(dec_configs[0]);
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
    unsigned long long *d_vlean = nullptr, *d_vpack2 = nullptr; // k_decode_lean: see DecArgs::v_lean
    std::vector<unsigned long long> h_vlean;                    // without special tokens (they are patched in on the device copy)
    uint32_t *d_sp_ids = nullptr, *d_sp_off = nullptr;
    uint8_t *d_sp_bytes = nullptr;
    uint32_t n_sp = 0;
    // encode-side special tokens: "a chunk with exactly these bytes is this id" (Tokenizer.h:667-671)
    uint32_t *d_esp_ids = nullptr, *d_esp_off = nullptr;
    uint8_t *d_esp_bytes = nullptr;
    uint32_t n_esp = 0;
    unsigned long long esp_len_mask = 0;
    // scratch (grown on demand)
    unsigned long long *d_status = nullptr; // look-back words of the tile kernels, one per tile of a launch
    uint32_t *d_spill = nullptr;            // encode: parking overflow, one block per CTA of the tile kernel
    uint64_t spill_words = 0;
    uint64_t status_cap = 0;
    uint32_t *d_small = nullptr; // [0] ticket, [1] n_long, [2] overflow, [3] scanned chunks
    unsigned long long *d_prof = nullptr; // MBPE_DEBUG: cycles per phase of k_encode_tiles (ENC_PROF_*)
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
    uint32_t cache_log2 = 0, cache_slots = 0, cache_log_cap = 0, cache_arena_cap = 0;
    uint32_t *d_cache_ctr = nullptr; // [0] log count, [1] used slots, [3] arena cursor
    uint64_t sub_batch_chunks = 0; // fixed sub-batch size (MBPE_ENCODE_SUBBATCH), 0 = geometric schedule
    uint64_t chunks_seen = 0;      // chunks encoded so far with this handle: how warm the caches are
    int cfg = 0; // kernel shape, see enc_configs
    // k_encode_hot: image of the shared-memory table of the hottest chunks, and the scratch its rebuild needs
    uint4 *d_hot_img = nullptr;
    uint32_t *d_hot_count = nullptr;
    unsigned long long *d_hot_best = nullptr;
    bool hot_valid = false;
    uint64_t hot_built_at = 0; // chunks_seen when the image was last rebuilt
    int dec_cfg = 0; // see dec_configs
    size_t l2_window_max = 0, l2_persist_bytes = 0;
    // mbpe_encode (host buffers): pinned staging + device buffers of the segment pipeline, two of each
    struct HostPipe {
        uint8_t *h_text = nullptr, *d_text = nullptr;
        uint32_t *h_off = nullptr, *d_off = nullptr, *h_ids = nullptr, *d_ids = nullptr;
        unsigned long long *h_out_off = nullptr, *d_out_off = nullptr;
    } pipe[2];
    uint64_t pipe_bytes = 0, pipe_chunks = 0;
    cudaStream_t pipe_stream = nullptr;
};

namespace mbpe {
struct EncConfig {
    int threads, cpt, ctas;
    void (*kernel)(const EncArgs);
    size_t smem;
    bool hot; // k_encode_hot (falls back to configuration 0 for the launches it does not take, see encode_device_impl)
};
#define ENC_CFG(T, C, M, P, L, R) EncConfig{T, C, M, k_encode_tiles<T, C, M, P, L, R>, sizeof(EncSmemT<T, C>), false}
#define ENC_HOT(T, C, P) EncConfig{T, C, 1, k_encode_hot<T, C, P>, sizeof(HotSmemT<T, C>), true}
// (threads, chunks per thread, CTAs per SM, probes in flight per thread, L1 policy of the probe loads, L2 prefetch pass); 0 = default, the others for A/B runs (MBPE_ENC_CFG)
static const EncConfig enc_configs[] = {ENC_CFG(256, 4, 4, 1, 0, 0), ENC_CFG(256, 4, 4, 1, 0, 1), ENC_CFG(256, 4, 3, 1, 0, 0), ENC_CFG(512, 4, 2, 1, 0, 0),
                                        ENC_CFG(256, 8, 2, 1, 0, 0), ENC_CFG(128, 4, 8, 1, 0, 0), ENC_CFG(512, 4, 2, 1, 0, 1), ENC_CFG(256, 8, 2, 1, 0, 1),
                                        ENC_HOT(1024, 4, 2), ENC_HOT(512, 8, 4)}; // 8, 9: k_encode_hot (experimental, measured slower: DESIGN.md section 3)
constexpr int N_ENC_CONFIGS = sizeof(enc_configs) / sizeof(enc_configs[0]);

static ChunkCache cache_view(const mbpe_encoder *e) {
    ChunkCache cc{};
    cc.slots = e->d_cache;
    cc.shift = 32 - e->cache_log2;
    cc.mask = e->cache_slots ? e->cache_slots - 1 : 0;
    cc.log = e->d_cache_log;
    cc.log_count = e->d_cache_ctr;
    cc.log_cap = e->cache_log_cap;
    cc.used = e->d_cache_ctr ? e->d_cache_ctr + 1 : nullptr;
    cc.arena = e->d_cache_arena;
    cc.arena_used = e->d_cache_ctr ? e->d_cache_ctr + 3 : nullptr;
    cc.arena_cap = e->cache_arena_cap;
    return cc;
}
static EncSpecials specials_view(const mbpe_encoder *e) {
    return EncSpecials{e->d_esp_ids, e->d_esp_off, e->d_esp_bytes, e->n_esp, e->esp_len_mask};
}
} // namespace mbpe

static int encoder_create_impl(mbpe_encoder *e, const uint32_t *merges, uint32_t n_merges, int device) {
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
        std::vector<unsigned long long> vpack2(vpack.size(), 0ull);
        e->h_vlean.assign(vpack.size(), 0ull);
        for (size_t i = 0; i < vpack.size(); i++) {
            const uint32_t l = voff[i + 1] - voff[i];
            unsigned long long v = l <= 7 ? l : l <= 126 ? (0x80u | l) : 0xFFu;
            for (uint32_t q = 0; q < std::min(l, 7u); q++) v |= (unsigned long long)vbytes[voff[i] + q] << (8 * (q + 1));
            e->h_vlean[i] = v;
            for (uint32_t q = 7; q < std::min(l, 15u); q++) vpack2[i] |= (unsigned long long)vbytes[voff[i] + q] << (8 * (q - 7));
        }
        MB_CUDA(cudaMalloc(&e->d_vlean, vpack.size() * 8));
        MB_CUDA(cudaMemcpy(e->d_vlean, e->h_vlean.data(), vpack.size() * 8, cudaMemcpyHostToDevice));
        MB_CUDA(cudaMalloc(&e->d_vpack2, vpack.size() * 8));
        MB_CUDA(cudaMemcpy(e->d_vpack2, vpack2.data(), vpack.size() * 8, cudaMemcpyHostToDevice));
    }
    MB_CUDA(cudaMalloc(&e->d_small, 16));
    MB_CUDA(cudaMalloc(&e->d_n_out, 8));
    if (getenv("MBPE_DEBUG")) {
        MB_CUDA(cudaMalloc(&e->d_prof, ENC_PROF_N * 8));
        MB_CUDA(cudaMemset(e->d_prof, 0, ENC_PROF_N * 8));
    }
    MB_CUDA(cudaMalloc(&e->d_sp_ids, 4));
    MB_CUDA(cudaMalloc(&e->d_sp_off, 8));
    MB_CUDA(cudaMalloc(&e->d_sp_bytes, 1));
    // "0" disables the cache; otherwise log2 of the slot count (default 22 = 256 MB)
    const char *cache_env = getenv("MBPE_ENCODE_CACHE");
    int cache_log2 = cache_env && *cache_env ? atoi(cache_env) : 22;
    if (cache_log2 >= 10 && cache_log2 <= 26) {
        e->cache_log2 = (uint32_t)cache_log2;
        e->cache_slots = 1u << cache_log2;
        e->cache_log_cap = 1u << 19;
        e->cache_arena_cap = 1u << 22;
        MB_CUDA(cudaMalloc(&e->d_cache, (uint64_t)e->cache_slots * sizeof(CacheSlot)));
        MB_CUDA(cudaMemset(e->d_cache, 0, (uint64_t)e->cache_slots * sizeof(CacheSlot)));
        MB_CUDA(cudaMalloc(&e->d_cache_log, (uint64_t)e->cache_log_cap * sizeof(CacheLogEntry)));
        MB_CUDA(cudaMalloc(&e->d_cache_arena, (uint64_t)e->cache_arena_cap * 4));
        MB_CUDA(cudaMalloc(&e->d_cache_ctr, 16));
        MB_CUDA(cudaMemset(e->d_cache_ctr, 0, 16));
        MB_CUDA(cudaMalloc(&e->d_hot_img, (size_t)HOT_N * sizeof(uint4)));
        MB_CUDA(cudaMemset(e->d_hot_img, 0, (size_t)HOT_N * sizeof(uint4)));
        MB_CUDA(cudaMalloc(&e->d_hot_count, (size_t)4 << HOT_COUNT_LOG2));
        MB_CUDA(cudaMalloc(&e->d_hot_best, (size_t)HOT_N * 8));
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
    const char *dcfg_env = getenv("MBPE_DEC_CFG");
    e->dec_cfg = dcfg_env && *dcfg_env ? std::min(std::max(atoi(dcfg_env), 0), N_DEC_CONFIGS - 1) : DEC_DEFAULT_CFG;
    for (int i = 0; i < N_DEC_CONFIGS; i++)
        MB_CUDA(cudaFuncSetAttribute(dec_configs[i].kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dec_configs[i].smem));
    for (int i = 0; i < N_ENC_CONFIGS; i++) {
        MB_CUDA(cudaFuncSetAttribute(enc_configs[i].kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)enc_configs[i].smem));
        if (const char *co = getenv("MBPE_ENC_CARVEOUT")) // percent of the L1/shared array given to shared memory (A/B runs)
            MB_CUDA(cudaFuncSetAttribute(enc_configs[i].kernel, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(co)));
    }
    return MBPE_OK;
}

extern "C" int mbpe_encoder_create(const uint32_t *merges, uint32_t n_merges, int device, mbpe_encoder **out) {
    if (!out || (n_merges && !merges)) return set_error(MBPE_E_INVALID, "null argument");
    *out = nullptr;
    if (256ull + n_merges > (1ull << ENC_ID_BITS)) return set_error(MBPE_E_INVALID, "vocab larger than 2^21 ids");
    mbpe_encoder *e = new mbpe_encoder();
    const int rc = encoder_create_impl(e, merges, n_merges, device);
    if (rc) { // nothing half-built survives a failure
        mbpe_encoder_destroy(e);
        return rc;
    }
    *out = e;
    return MBPE_OK;
}

static void free_host_pipe(mbpe_encoder *e) {
    for (auto &p : e->pipe) {
        cudaFreeHost(p.h_text);
        cudaFreeHost(p.h_off);
        cudaFreeHost(p.h_ids);
        cudaFreeHost(p.h_out_off);
        cudaFree(p.d_text);
        cudaFree(p.d_off);
        cudaFree(p.d_ids);
        cudaFree(p.d_out_off);
        p = mbpe_encoder::HostPipe();
    }
    e->pipe_bytes = e->pipe_chunks = 0;
}

extern "C" void mbpe_encoder_destroy(mbpe_encoder *e) {
    if (!e) return;
    cudaSetDevice(e->device);
    void *ps[] = {e->d_slots, e->d_voff, e->d_vbytes, e->d_vpack, e->d_sp_ids, e->d_sp_off, e->d_sp_bytes, e->d_status, e->d_small,
                  e->d_n_out, e->d_long_list, e->d_scratch_a, e->d_scratch_b, e->d_cache, e->d_cache_log,
                  e->d_cache_ctr, e->d_cache_arena, e->d_esp_ids, e->d_esp_off, e->d_esp_bytes, e->d_prof, e->d_spill,
                  e->d_hot_img, e->d_hot_count, e->d_hot_best, e->d_vlean, e->d_vpack2};
    for (void *p : ps) cudaFree(p);
    free_host_pipe(e);
    if (e->pipe_stream) cudaStreamDestroy(e->pipe_stream);
    cudaGetLastError();
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
    e->d_sp_ids = e->d_sp_off = nullptr;
    e->d_sp_bytes = nullptr;
    e->n_sp = 0;
    MB_CUDA(cudaMalloc(&e->d_sp_ids, std::max<size_t>(sid.size(), 1) * 4));
    MB_CUDA(cudaMalloc(&e->d_sp_off, soff.size() * 4));
    MB_CUDA(cudaMalloc(&e->d_sp_bytes, std::max<size_t>(sb.size(), 1)));
    MB_CUDA(cudaMemcpy(e->d_sp_ids, sid.data(), sid.size() * 4, cudaMemcpyHostToDevice));
    MB_CUDA(cudaMemcpy(e->d_sp_off, soff.data(), soff.size() * 4, cudaMemcpyHostToDevice));
    MB_CUDA(cudaMemcpy(e->d_sp_bytes, sb.data(), sb.size(), cudaMemcpyHostToDevice));
    e->n_sp = (uint32_t)sid.size();
    // k_decode_lean finds a special token that overrides a vocabulary id (Tokenizer.h:733-736) in the table word itself
    if (e->d_vlean) {
        std::vector<unsigned long long> vl = e->h_vlean;
        for (size_t k = 0; k < sid.size(); k++)
            if (sid[k] < vl.size()) vl[sid[k]] = PKL_SPECIAL | ((unsigned long long)k << 8) | 0xFFull;
        MB_CUDA(cudaMemcpy(e->d_vlean, vl.data(), vl.size() * 8, cudaMemcpyHostToDevice));
    }
    return MBPE_OK;
}

// Special tokens on the encode side (Tokenizer.h:605-650, :667-671): a special token is a chunk of its own that becomes
// one ready-made id. The device front end marks those chunks (mbpe_pretok_split_device_parts); the merge scan matches
// them by exact compare in its slow path (tile kernel and long-chunk kernel alike), so the ids never depend on a cache.
// For speed the tokens of up to 31 bytes are also put into the chunk caches as "these bytes -> this id". An ordinary
// chunk can never carry the bytes of a special token (the splitter would have cut it out). The caches are emptied
// first: they may have learned those bytes as ordinary text before. Any token length, any count.
extern "C" int mbpe_encoder_seed_special_chunks(mbpe_encoder *e, const uint32_t *ids, const uint8_t *bytes, const uint64_t *off,
                                                uint32_t n) {
    if (!e || (n && (!ids || !bytes || !off))) return set_error(MBPE_E_INVALID, "null argument");
    int rc = use_device(e->device);
    if (rc) return rc;
    std::vector<uint32_t> soff(n + 1, 0);
    unsigned long long len_mask = 0;
    for (uint32_t i = 0; i < n; i++) {
        const uint64_t len = off[i + 1] - off[i];
        if (len == 0 || off[i + 1] - off[0] >= (1ull << 31)) return set_error(MBPE_E_INVALID, "empty or oversized special token");
        soff[i + 1] = (uint32_t)(off[i + 1] - off[0]);
        len_mask |= 1ull << std::min<uint64_t>(len, 63);
    }
    MB_CUDA(cudaDeviceSynchronize()); // nothing may still be reading the old tables
    cudaFree(e->d_esp_ids);
    cudaFree(e->d_esp_off);
    cudaFree(e->d_esp_bytes);
    e->d_esp_ids = e->d_esp_off = nullptr;
    e->d_esp_bytes = nullptr;
    e->n_esp = 0;
    e->esp_len_mask = 0;
    if (n) {
        MB_CUDA(cudaMalloc(&e->d_esp_ids, (size_t)n * 4));
        MB_CUDA(cudaMalloc(&e->d_esp_off, ((size_t)n + 1) * 4));
        MB_CUDA(cudaMalloc(&e->d_esp_bytes, std::max<size_t>(soff[n], 1)));
        MB_CUDA(cudaMemcpy(e->d_esp_ids, ids, (size_t)n * 4, cudaMemcpyHostToDevice));
        MB_CUDA(cudaMemcpy(e->d_esp_off, soff.data(), ((size_t)n + 1) * 4, cudaMemcpyHostToDevice));
        MB_CUDA(cudaMemcpy(e->d_esp_bytes, bytes + off[0], soff[n], cudaMemcpyHostToDevice));
        e->n_esp = n;
        e->esp_len_mask = len_mask;
    }
    if (!e->d_cache) return MBPE_OK;
    std::vector<CacheLogEntry> log;
    for (uint32_t i = 0; i < n && log.size() < e->cache_log_cap; i++) {
        const uint64_t len = off[i + 1] - off[i];
        if (len > CACHE_MAX_LEN) continue; // matched by the slow path only
        CacheLogEntry le;
        memset(&le, 0, sizeof le);
        for (uint64_t q = 0; q < len; q++) le.k[q >> 3] |= (uint64_t)bytes[off[i] + q] << ((q & 7) * 8);
        le.k[3] |= len << 56;
        le.n = 1;
        le.ids[0] = ids[i];
        log.push_back(le);
    }
    MB_CUDA(cudaMemset(e->d_cache, 0, (uint64_t)e->cache_slots * sizeof(CacheSlot)));
    MB_CUDA(cudaMemset(e->d_cache_ctr, 0, 16));
    MB_CUDA(cudaMemset(e->d_hot_img, 0, (size_t)HOT_N * sizeof(uint4))); // (its entries come from the cache)
    e->hot_valid = false;
    e->hot_built_at = 0;
    e->chunks_seen = 0;
    if (!log.empty()) {
        MB_CUDA(cudaMemcpy(e->d_cache_log, log.data(), log.size() * sizeof(CacheLogEntry), cudaMemcpyHostToDevice));
        const uint32_t cnt = (uint32_t)log.size();
        MB_CUDA(cudaMemcpy(e->d_cache_ctr, &cnt, 4, cudaMemcpyHostToDevice));
        k_cache_insert<<<1, 256>>>(cache_view(e));
        MB_CUDA(cudaGetLastError());
        MB_CUDA(cudaDeviceSynchronize());
    }
    return MBPE_OK;
}

static int ensure_status(mbpe_encoder *e, uint64_t n_tiles) {
    if (n_tiles > e->status_cap) {
        cudaFree(e->d_status);
        e->d_status = nullptr;
        e->status_cap = 0;
        const uint64_t cap = n_tiles + n_tiles / 4 + 64;
        MB_CUDA(cudaMalloc(&e->d_status, cap * 8));
        e->status_cap = cap;
    }
    return MBPE_OK;
}

extern "C" uint64_t mbpe_encoder_launches(const mbpe_encoder *e) { return e ? e->launches : 0; }

extern "C" int mbpe_encode_reserve(mbpe_encoder *e, uint64_t n_bytes, uint64_t n_chunks) {
    if (!e) return set_error(MBPE_E_INVALID, "null argument");
    (void)n_bytes;
    int rc = use_device(e->device);
    if (rc) return rc;
    uint64_t tile_chunks = ~0ull, sm_chunks = 0; // smallest tile of the configurations a launch may use; most chunks in flight per SM
    for (int c : {e->cfg, 0}) {
        tile_chunks = std::min<uint64_t>(tile_chunks, (uint64_t)enc_configs[c].threads * enc_configs[c].cpt);
        sm_chunks = std::max<uint64_t>(sm_chunks, (uint64_t)enc_configs[c].threads * enc_configs[c].cpt * enc_configs[c].ctas);
    }
    const uint64_t max_sb = e->sub_batch_chunks ? e->sub_batch_chunks : ENC_MAX_SUBBATCH;
    if ((rc = ensure_status(e, (std::min(n_chunks, max_sb) + tile_chunks - 1) / tile_chunks + 1))) return rc;
    if (e->long_cap == 0) {
        MB_CUDA(cudaMalloc(&e->d_long_list, (1 << 16) * 4));
        e->long_cap = 1 << 16;
    }
    // parking overflow of the tile kernel: one block of TILE * ENC_SHORT_MAX words per CTA (the worst case; untouched pages
    // cost nothing but address space)
    const uint64_t need = (uint64_t)e->sms * sm_chunks * ENC_SHORT_MAX;
    if (need > e->spill_words) {
        cudaFree(e->d_spill);
        e->d_spill = nullptr;
        e->spill_words = 0;
        MB_CUDA(cudaMalloc(&e->d_spill, need * 4));
        e->spill_words = need;
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
    if (n_chunks == 0) {
        MB_CUDA(cudaMemsetAsync(d_n_out, 0, 8, st));
        if (d_out_off) MB_CUDA(cudaMemsetAsync(d_out_off, 0, 8, st));
        return MBPE_OK;
    }
    EncArgs a{};
    a.tab = EncTable{e->d_slots, e->mask};
    a.cache = cache_view(e);
    a.sp = specials_view(e);
    a.bytes = d_bytes;
    a.n_bytes_total = n_bytes;
    a.off = d_off;
    a.n_chunks = n_chunks;
    a.out = d_out;
    a.out_cap = out_cap;
    a.d_n_out = (unsigned long long *)d_n_out;
    a.out_off = d_out_off;
    a.stream_base = (const unsigned long long *)d_n_out;
    a.status = e->d_status;
    a.ticket = e->d_small;
    a.long_list = e->d_long_list;
    a.n_long = e->d_small + 1;
    a.long_cap = (uint32_t)e->long_cap;
    a.overflow = e->d_small + 2;
    a.miss_count = e->d_small + 3;
    a.spill = e->d_spill;
    a.prof = e->d_prof;
    // cp.async.bulk needs 16-byte aligned sources (device allocations are; offsets into them may not be)
    a.bulk = ((((uintptr_t)d_bytes) | ((uintptr_t)d_off)) & 15) == 0 && !getenv("MBPE_ENC_NO_BULK");
    a.out_aligned = (((uintptr_t)d_out) & 15) == 0;
    if (const char *ab = getenv("MBPE_ENC_ABLATE")) a.ablate = (uint32_t)atoi(ab);
    if (const char *pf = getenv("MBPE_ENC_PF")) a.pf_ahead = (uint32_t)atoi(pf);
    if (a.pf_ahead == 1) a.pf_ahead = (uint32_t)(e->sms * enc_configs[e->cfg].ctas); // one round of tickets
    a.hot_img = e->d_hot_img;
    uint32_t small[4] = {0, 0, 0, 0};
    // First launch sequence is optimistic: no chunk is longer than ENC_SHORT_MAX bytes (true for the regex patterns on
    // ordinary text), so the boundaries are read exactly once. The tile kernel reports the long chunks it meets
    // instead of encoding them; if there are any, they go through k_encode_long and the sequence runs again.
    for (int attempt = 0; attempt < 2; attempt++) {
        // k_encode_hot takes the common case: aligned buffers, chunk cache on, no per-chunk offsets wanted, no long chunks
        // known (it reports them like k_encode_tiles does; the repeat run with the splice enabled is k_encode_tiles')
        const bool use_hot = enc_configs[e->cfg].hot && a.bulk && a.cache.slots && !a.out_off && !a.scratch_a && !a.ablate;
        const EncConfig &kc = enc_configs[use_hot || !enc_configs[e->cfg].hot ? e->cfg : 0];
        const uint64_t ET_CHUNKS = (uint64_t)kc.threads * kc.cpt;
        MB_CUDA(cudaMemsetAsync(e->d_small, 0, 16, st));
        MB_CUDA(cudaMemsetAsync(d_n_out, 0, 8, st));
        // Launches of whole tiles: the cache learns from one launch before the next one starts (a cold encoder starts
        // with small launches, 1 M chunks, and doubles; a warm one goes straight to large ones), and the ids of launch
        // i+1 continue the stream where launch i ended (*d_n_out).
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
            // the hot table follows the text: rebuilt from a sample of this launch's chunks at the start of a call and
            // whenever the encoder has seen twice as many chunks as at the last rebuild (the chunk cache it draws the ids
            // from is still filling up then)
            const uint64_t seen_before = e->chunks_seen - (a.chunk1 - a.chunk0);
            if (use_hot && (!e->hot_valid || (a.chunk0 == 0 && a.chunk1 >= (1u << 16)) || seen_before >= 2 * e->hot_built_at)) {
                HotBuild hb{};
                hb.bytes = d_bytes;
                hb.off = d_off;
                hb.chunk0 = a.chunk0;
                hb.n_sample = (uint32_t)std::min<uint64_t>(a.chunk1 - a.chunk0, 1u << 18);
                hb.stride = (a.chunk1 - a.chunk0) / hb.n_sample;
                hb.count = e->d_hot_count;
                hb.best = e->d_hot_best;
                hb.img = e->d_hot_img;
                MB_CUDA(cudaMemsetAsync(e->d_hot_count, 0, (size_t)4 << HOT_COUNT_LOG2, st));
                MB_CUDA(cudaMemsetAsync(e->d_hot_best, 0, (size_t)HOT_N * 8, st));
                MB_CUDA(cudaMemsetAsync(e->d_hot_img, 0, (size_t)HOT_N * sizeof(uint4), st));
                const unsigned gb = (hb.n_sample + 255) / 256;
                k_hot_count<<<gb, 256, 0, st>>>(hb);
                k_hot_elect<<<gb, 256, 0, st>>>(hb);
                k_hot_fill<<<gb, 256, 0, st>>>(hb);
                k_hot_resolve<<<HOT_N / 256, 256, 0, st>>>(hb, a.cache);
                e->launches += 4;
                e->hot_valid = true;
                e->hot_built_at = seen_before;
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
                // the text, boundaries and ids stream through L2 once; the randomly probed chunk cache is what should stay
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
            if (a.cache.slots) { // what the launch had to scan goes into the cache before the next one starts
                k_cache_insert<<<e->sms * 2, 256, 0, st>>>(a.cache);
                e->launches++;
            }
        }
        MB_CUDA(cudaGetLastError());
        MB_CUDA(cudaMemcpyAsync(small, e->d_small, 16, cudaMemcpyDeviceToHost, st));
        MB_CUDA(cudaStreamSynchronize(st));
        const uint32_t n_long = small[1];
        if (attempt == 1 || n_long == 0) break;
        if (n_long > e->long_cap) { // the report overflowed: find them all again with a list that is large enough
            cudaFree(e->d_long_list);
            e->d_long_list = nullptr;
            e->long_cap = 0;
            MB_CUDA(cudaMalloc(&e->d_long_list, ((uint64_t)n_long + 1024) * 4));
            e->long_cap = (uint64_t)n_long + 1024;
            a.long_list = e->d_long_list;
            a.long_cap = (uint32_t)e->long_cap;
            MB_CUDA(cudaMemsetAsync(e->d_small + 1, 0, 4, st));
            unsigned grid = (unsigned)std::min<uint64_t>((n_chunks + 255) / 256, (uint64_t)e->sms * 8);
            k_encode_find_long<<<grid, 256, 0, st>>>(d_off, n_chunks, e->d_long_list, e->d_small + 1, (uint32_t)e->long_cap);
            e->launches++;
        }
        if (n_bytes + 1 > e->scratch_cap) {
            cudaFree(e->d_scratch_a);
            cudaFree(e->d_scratch_b);
            e->d_scratch_a = e->d_scratch_b = nullptr;
            e->scratch_cap = 0;
            MB_CUDA(cudaMalloc(&e->d_scratch_a, (n_bytes + 1) * 4));
            MB_CUDA(cudaMalloc(&e->d_scratch_b, (n_bytes + 1) * 4));
            e->scratch_cap = n_bytes + 1;
        }
        k_encode_long<<<(n_long + 63) / 64, 64, 0, st>>>(a.tab, a.sp, d_bytes, d_off, e->d_long_list, n_long, e->d_scratch_a,
                                                         e->d_scratch_b);
        e->launches++;
        a.scratch_a = e->d_scratch_a;
        a.scratch_b = e->d_scratch_b;
    }
    if (getenv("MBPE_DEBUG")) {
        uint32_t ctr[4] = {0, 0, 0, 0};
        if (e->d_cache_ctr) MB_CUDA(cudaMemcpy(ctr, e->d_cache_ctr, 16, cudaMemcpyDeviceToHost));
        fprintf(stderr, "[mbpe] encode: %llu chunks, %u scanned (cache misses), %u long, cache entries %u of %u%s\n",
                (unsigned long long)n_chunks, small[3], small[1], ctr[1], e->cache_slots,
                a.bulk ? ", bulk staging" : ", cooperative staging (unaligned buffers)");
        if (e->d_prof) {
            unsigned long long pr[ENC_PROF_N];
            MB_CUDA(cudaMemcpy(pr, e->d_prof, sizeof pr, cudaMemcpyDeviceToHost));
            MB_CUDA(cudaMemset(e->d_prof, 0, sizeof pr));
            if (enc_configs[e->cfg].hot) {
                const double w = pr[11] ? (double)pr[11] : 1.0;
                fprintf(stderr, "[mbpe] k_encode_hot: hot table %llu, home slot %llu, rest of the probe sequence %llu, slow path %llu chunks; warp tiles %llu, "
                                "cycles per warp tile: keys+hot %.0f, chunk cache %.0f, scan+place %.0f, gather %.0f, wait for previous base %.0f, "
                                "store previous %.0f, all %.0f\n",
                        pr[0], pr[1], pr[3], pr[2], pr[11], pr[4] / w, pr[5] / w, pr[6] / w, pr[7] / w, pr[8] / w, pr[9] / w, pr[10] / w);
                return MBPE_OK;
            }
            const double t = pr[7] ? (double)pr[7] : 1.0;
            fprintf(stderr, "[mbpe] encode tiles %llu, cycles per tile: wait data %.0f, probes %.0f, scan list %.0f, count scan %.0f, "
                            "look-back+gather %.0f, fetch next %.0f, store %.0f\n",
                    pr[7], pr[0] / t, pr[1] / t, pr[2] / t, pr[3] / t, pr[4] / t, pr[5] / t, pr[6] / t);
            fprintf(stderr, "[mbpe]   thread 0: look-back %.0f, gather %.0f, barrier %.0f; thread 32: gather %.0f, barrier %.0f\n", pr[8] / t,
                    pr[9] / t, pr[10] / t, pr[11] / t, pr[12] / t);
        }
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

// ---------------------------------------------------------------------------------------------------------
// mbpe_encode: host bytes + host u64 chunk offsets in, host ids out. The chunk list is cut into segments of whole
// chunks; a segment is staged into pinned memory by a few host threads (text copied, offsets narrowed to u32 relative
// to the segment), goes up, through the merge scan and down into pinned memory, and is copied out to the caller's
// buffer -- staging of segment k+1 and copy-out of segment k-1 overlap the device work of segment k.
// ---------------------------------------------------------------------------------------------------------
namespace {
constexpr uint64_t HP_SEG_BYTES = 32ull << 20, HP_SEG_CHUNKS = 8ull << 20;

int ensure_host_pipe(mbpe_encoder *e, uint64_t seg_bytes, uint64_t seg_chunks, bool want_off) {
    if (!e->pipe_stream) MB_CUDA(cudaStreamCreateWithFlags(&e->pipe_stream, cudaStreamNonBlocking));
    const bool have_off = e->pipe[0].h_out_off != nullptr;
    if (seg_bytes <= e->pipe_bytes && seg_chunks <= e->pipe_chunks && (!want_off || have_off)) return MBPE_OK;
    const uint64_t nb = std::max(seg_bytes, e->pipe_bytes), nc = std::max(seg_chunks, e->pipe_chunks);
    free_host_pipe(e);
    for (auto &p : e->pipe) {
        MB_CUDA(cudaMallocHost(&p.h_text, std::max<uint64_t>(nb, 16)));
        MB_CUDA(cudaMallocHost(&p.h_off, (nc + 1) * 4));
        MB_CUDA(cudaMallocHost(&p.h_ids, std::max<uint64_t>(nb, 16) * 4));
        MB_CUDA(cudaMalloc(&p.d_text, std::max<uint64_t>(nb, 16)));
        MB_CUDA(cudaMalloc(&p.d_off, (nc + 1) * 4));
        MB_CUDA(cudaMalloc(&p.d_ids, std::max<uint64_t>(nb, 16) * 4));
        if (want_off || have_off) {
            MB_CUDA(cudaMallocHost(&p.h_out_off, (nc + 1) * 8));
            MB_CUDA(cudaMalloc(&p.d_out_off, (nc + 1) * 8));
        }
    }
    e->pipe_bytes = nb;
    e->pipe_chunks = nc;
    return MBPE_OK;
}

// run f(part, n_parts) on a few threads (the calling thread is one of them)
template <class F>
void on_threads(unsigned n, const F &f) {
    std::vector<std::thread> th;
    for (unsigned i = 1; i < n; i++) th.emplace_back([&f, i, n]() { f(i, n); });
    f(0u, n);
    for (auto &t : th) t.join();
}
} // namespace

extern "C" int mbpe_encode(mbpe_encoder *e, const uint8_t *bytes, uint64_t n_bytes, const uint64_t *chunk_off,
                           uint64_t n_chunks, uint32_t *out_tokens, uint64_t out_cap, uint64_t *n_out,
                           uint64_t *out_off) {
    if (!e || !chunk_off || !n_out || (n_bytes && !bytes)) return set_error(MBPE_E_INVALID, "null argument");
    if (chunk_off[0] != 0 || chunk_off[n_chunks] != n_bytes)
        return set_error(MBPE_E_INVALID, "chunk_off must start at 0 and end at n_bytes");
    int rc = use_device(e->device);
    if (rc) return rc;
    *n_out = 0;
    if (out_off) out_off[0] = 0;
    // segments of whole chunks: [seg[k], seg[k+1]) in chunk indices
    std::vector<uint64_t> seg{0};
    uint64_t max_b = 0, max_c = 0;
    for (uint64_t c0 = 0; c0 < n_chunks;) {
        const uint64_t b0 = chunk_off[c0];
        // last chunk that still ends inside the byte budget (binary search: the offsets are checked for order below)
        uint64_t lo = c0 + 1, hi = std::min(n_chunks, c0 + HP_SEG_CHUNKS);
        while (lo < hi) {
            const uint64_t mid = (lo + hi + 1) / 2;
            if (chunk_off[mid] >= b0 && chunk_off[mid] - b0 <= HP_SEG_BYTES)
                lo = mid;
            else
                hi = mid - 1;
        }
        const uint64_t c1 = lo;
        if (chunk_off[c1] < b0) return set_error(MBPE_E_INVALID, "chunk_off not monotonic");
        if (chunk_off[c1] - b0 >= (1ull << 32)) return set_error(MBPE_E_INVALID, "a single chunk of 4 GiB or more is not supported");
        max_b = std::max(max_b, chunk_off[c1] - b0);
        max_c = std::max(max_c, c1 - c0);
        seg.push_back(c1);
        c0 = c1;
    }
    const size_t n_seg = seg.size() - 1;
    if (n_seg == 0) return MBPE_OK;
    if ((rc = ensure_host_pipe(e, max_b, max_c, out_off != nullptr))) return rc;
    cudaStream_t st = e->pipe_stream;
    const unsigned n_thr = (unsigned)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    std::atomic<int> bad_order{0};
    auto stage_in = [&](size_t k) {
        auto &p = e->pipe[k & 1];
        const uint64_t c0 = seg[k], c1 = seg[k + 1], b0 = chunk_off[c0], nb = chunk_off[c1] - b0, nc = c1 - c0;
        on_threads(nb + nc * 12 > (4u << 20) ? n_thr : 1u, [&](unsigned part, unsigned parts) {
            const uint64_t x0 = nb * part / parts, x1 = nb * (part + 1) / parts;
            memcpy(p.h_text + x0, bytes + b0 + x0, x1 - x0);
            const uint64_t i0 = (nc + 1) * part / parts, i1 = (nc + 1) * (part + 1) / parts;
            bool bad = false;
            for (uint64_t i = i0; i < i1; i++) {
                const uint64_t v = chunk_off[c0 + i];
                bad |= (i && v < chunk_off[c0 + i - 1]) || v < b0 || v - b0 > nb;
                p.h_off[i] = (uint32_t)(v - b0);
            }
            if (bad) bad_order.store(1);
        });
    };
    uint64_t produced = 0;
    int status = MBPE_OK;
    struct Done {
        uint64_t n_ids, at;
    };
    std::vector<Done> done(n_seg);
    auto stage_out = [&](size_t k) {
        auto &p = e->pipe[k & 1];
        const uint64_t c0 = seg[k], nc = seg[k + 1] - c0, n = done[k].n_ids, at = done[k].at;
        on_threads(n * 4 + (out_off ? nc * 8 : 0) > (4u << 20) ? n_thr : 1u, [&](unsigned part, unsigned parts) {
            if (out_tokens && at + n <= out_cap) {
                const uint64_t x0 = n * part / parts, x1 = n * (part + 1) / parts;
                memcpy(out_tokens + at + x0, p.h_ids + x0, (x1 - x0) * 4);
            }
            if (out_off) {
                const uint64_t i0 = (nc + 1) * part / parts, i1 = (nc + 1) * (part + 1) / parts;
                for (uint64_t i = i0; i < i1; i++) out_off[c0 + i] = p.h_out_off[i] + at;
            }
        });
    };
    stage_in(0);
    std::thread helper;
    for (size_t k = 0; k < n_seg && rc == MBPE_OK; k++) {
        if (bad_order.load()) {
            rc = set_error(MBPE_E_INVALID, "chunk_off not monotonic");
            break;
        }
        auto &p = e->pipe[k & 1];
        const uint64_t c0 = seg[k], c1 = seg[k + 1], nb = chunk_off[c1] - chunk_off[c0], nc = c1 - c0;
        // the other buffer pair is free: segment k-1 was copied out before the device work of k-1 ... see below
        helper = std::thread([&, k]() {
            if (k >= 1) stage_out(k - 1);
            if (k + 1 < n_seg) stage_in(k + 1);
        });
        cudaError_t ce = cudaMemcpyAsync(p.d_text, p.h_text, nb, cudaMemcpyHostToDevice, st);
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(p.d_off, p.h_off, (nc + 1) * 4, cudaMemcpyHostToDevice, st);
        if (ce == cudaSuccess) rc = encode_device_impl(e, p.d_text, nb, p.d_off, nc, p.d_ids, std::max<uint64_t>(nb, 1), (uint64_t *)e->d_n_out, p.d_out_off && out_off ? p.d_out_off : nullptr, st);
        uint64_t n = 0;
        if (ce == cudaSuccess && rc == MBPE_OK) ce = cudaMemcpyAsync(&n, e->d_n_out, 8, cudaMemcpyDeviceToHost, st);
        if (ce == cudaSuccess && rc == MBPE_OK) ce = cudaStreamSynchronize(st);
        if (ce == cudaSuccess && rc == MBPE_OK && n) ce = cudaMemcpyAsync(p.h_ids, p.d_ids, n * 4, cudaMemcpyDeviceToHost, st);
        if (ce == cudaSuccess && rc == MBPE_OK && out_off) ce = cudaMemcpyAsync(p.h_out_off, p.d_out_off, (nc + 1) * 8, cudaMemcpyDeviceToHost, st);
        if (ce == cudaSuccess && rc == MBPE_OK) ce = cudaStreamSynchronize(st);
        helper.join();
        if (rc == MBPE_OK && ce != cudaSuccess) rc = cuda_fail(ce, "mbpe_encode pipeline", __FILE__, __LINE__);
        if (rc) break;
        done[k] = Done{n, produced};
        if (!out_tokens || produced + n > out_cap) status = MBPE_E_CAPACITY;
        produced += n;
    }
    if (helper.joinable()) helper.join();
    if (rc) return rc;
    if (bad_order.load()) return set_error(MBPE_E_INVALID, "chunk_off not monotonic");
    stage_out(n_seg - 1);
    *n_out = produced;
    if (status == MBPE_E_CAPACITY) return set_error(status, "out_tokens too small; *n_out holds the needed count");
    return MBPE_OK;
}

static int decode_launch(mbpe_encoder *e, const uint32_t *d_ids, uint64_t n_ids, uint8_t *d_out, uint64_t out_cap,
                         unsigned long long *d_n_out, cudaStream_t st) {
    const bool nopack = getenv("MBPE_DEC_NOPACK") != nullptr;
    const DecConfig &dc = dec_configs[nopack ? 3 : e->dec_cfg]; // (configuration 3 keeps no table in shared memory)
    const uint64_t tile_ids = (uint64_t)dc.tile_ids;
    const uint64_t n_tiles = (n_ids + tile_ids - 1) / tile_ids;
    if (n_tiles >= 0xFFFFFFFFull) return set_error(MBPE_E_INVALID, "too many ids in one call");
    int rc = ensure_status(e, n_tiles);
    if (rc) return rc;
    DecArgs a{};
    a.ids = d_ids;
    a.n_ids = n_ids;
    a.v_off = e->d_voff;
    a.v_bytes = e->d_vbytes;
    a.v_pack = nopack ? nullptr : e->d_vpack;
    a.v_lean = e->d_vlean;
    {
        const char *pf = getenv("MBPE_DEC_PF");
        a.pf_ahead = pf && *pf ? (uint32_t)atoi(pf) : 1u; // default: one round of tickets ahead (2.89 -> 2.79 ms per GiB of text)
        if (a.pf_ahead == 1) a.pf_ahead = (uint32_t)(e->sms * (dc.threads / dc.group)); // one round of tickets
    }
    a.v_pack2 = e->d_vpack2;
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
    const uint64_t groups_per_cta = dc.threads / dc.group;
    const unsigned grid = (unsigned)std::min<uint64_t>((n_tiles + groups_per_cta - 1) / groups_per_cta, (uint64_t)e->sms * dc.ctas);
    dc.kernel<<<grid, dc.threads, dc.smem, st>>>(a);
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
