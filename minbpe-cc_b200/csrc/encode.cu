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
// k_encode_tiles: one thread per chunk (chunks <= ENC_SHORT_MAX bytes), tokens in a per-thread buffer; the flat
// output position comes from a block scan + decoupled look-back over tiles, so the stream is produced in ONE
// pass: every text byte and boundary is read once, every id written once (SURVEY 8(d) B_enc).
// Chunks longer than ENC_SHORT_MAX (encoder "basic": the whole text is one chunk) are encoded first by
// k_encode_long into a scratch stream and spliced in by k_encode_tiles.
#include <algorithm>
#include <vector>

#include "common.cuh"

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
// decoupled look-back over tiles (status word = flag:2 | value:62)
// ---------------------------------------------------------------------------------------------------------
constexpr uint64_t LB_AGG = 1ull << 62, LB_PREFIX = 2ull << 62, LB_VAL = (1ull << 62) - 1;

__device__ __forceinline__ uint64_t lookback_base(unsigned long long *status, uint32_t tile, uint64_t total) {
    if (tile == 0) {
        atomicExch(&status[0], LB_PREFIX | total);
        return 0;
    }
    atomicExch(&status[tile], LB_AGG | total);
    uint64_t acc = 0;
    for (int64_t j = (int64_t)tile - 1;; j--) {
        unsigned long long v;
        do {
            v = *((volatile unsigned long long *)&status[j]);
        } while ((v >> 62) == 0);
        acc += v & LB_VAL;
        if (v & LB_PREFIX) break;
    }
    atomicExch(&status[tile], LB_PREFIX | (acc + total));
    return acc;
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
struct EncArgs {
    EncTable tab;
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
};

// A tile = ET_CHUNKS consecutive chunks. Its text is staged into shared memory with coalesced 16-byte loads,
// one u32 token slot per byte; each thread owns ET_CPT consecutive chunks (a private, contiguous slot range) and
// runs the passes in place. Per pass the pair lookups of up to 8 positions are issued together (independent
// read-only loads in flight), then resolved left to right.
constexpr int ET_CPT = 4;
constexpr int ET_CHUNKS = ENC_THREADS * ET_CPT;
constexpr int ET_CAP = 10240; // staged bytes per tile; a tile with more text takes the unstaged path
constexpr uint32_t ENC_NONE = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t enc_lookup_id(const EncTable &t, uint32_t a, uint32_t b) {
    uint32_t id;
    return enc_lookup(t, a, b, id) ? id : ENC_NONE;
}

// passes over t[0..len) (shared memory, private to the thread), in place. Returns the final length.
__device__ __forceinline__ uint32_t enc_chunk_smem(const EncTable &tab, uint32_t *t, uint32_t len) {
    while (len >= 2) {
        uint32_t w = 0;
        bool skip = false, merged = false;
        for (uint32_t base = 0; base < len; base += 8) {
            uint32_t v[9], id[8];
#pragma unroll
            for (int i = 0; i < 9; i++) v[i] = (base + i < len) ? t[base + i] : 0u;
#pragma unroll
            for (int i = 0; i < 8; i++) id[i] = (base + i + 1 < len) ? enc_lookup_id(tab, v[i], v[i + 1]) : ENC_NONE;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (base + i < len) {
                    if (skip) {
                        skip = false;
                    } else if (id[i] != ENC_NONE) {
                        t[w++] = id[i];
                        skip = true;
                        merged = true;
                    } else {
                        t[w++] = v[i];
                    }
                }
            }
        }
        len = w;
        if (!merged) break;
    }
    return len;
}

__global__ void __launch_bounds__(ENC_THREADS, 4) k_encode_tiles(const EncArgs a) {
    __shared__ uint32_t s_off[ET_CHUNKS + 1];
    __shared__ uint32_t s_tok[ET_CAP];
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_warp[ENC_THREADS / 32];
    __shared__ unsigned long long s_base;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_tile = atomicAdd(a.ticket, 1u); // tiles start in order: look-back cannot deadlock
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= a.n_tiles) return;
        const uint64_t c0 = (uint64_t)tile * ET_CHUNKS;
        const uint32_t nc = (uint32_t)min((uint64_t)ET_CHUNKS, a.n_chunks - c0);
        for (uint32_t i = tid; i <= nc; i += ENC_THREADS) s_off[i] = __ldg(&a.off[c0 + i]);
        __syncthreads();
        const uint32_t b0 = s_off[0], b1 = s_off[nc], nb = b1 - b0;
        const bool staged = nb <= ET_CAP;
        if (staged) {
            const uint32_t a0 = b0 & ~15u; // 16-byte aligned window start (cudaMalloc'd buffers are 256-aligned)
            for (uint32_t v = tid; a0 + v * 16 < b1; v += ENC_THREADS) {
                const uint32_t g0 = a0 + v * 16;
                if (g0 + 16 <= a.n_bytes_total) {
                    uint4 q = __ldg(reinterpret_cast<const uint4 *>(a.bytes + g0));
                    uint32_t wds[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        uint32_t g = g0 + j;
                        if (g >= b0 && g < b1) s_tok[g - b0] = (wds[j >> 2] >> ((j & 3) * 8)) & 0xFFu;
                    }
                } else {
                    for (uint32_t g = max(g0, b0); g < b1 && g < g0 + 16; g++) s_tok[g - b0] = __ldg(&a.bytes[g]);
                }
            }
            __syncthreads();
        }
        uint32_t cnt[ET_CPT], sum = 0;
#pragma unroll
        for (int j = 0; j < ET_CPT; j++) {
            const uint32_t k = tid * ET_CPT + j;
            cnt[j] = 0;
            if (k < nc) {
                const uint32_t o = s_off[k], len = s_off[k + 1] - o;
                if (len > ENC_SHORT_MAX) {
                    cnt[j] = a.scratch_b[o]; // encoded by k_encode_long
                } else if (staged) {
                    cnt[j] = enc_chunk_smem(a.tab, &s_tok[o - b0], len);
                } else { // unstaged tile: count now, encode again when writing
                    uint32_t t[ENC_SHORT_MAX];
                    for (uint32_t i = 0; i < len; i++) t[i] = __ldg(&a.bytes[o + i]);
                    uint32_t l2 = len;
                    bool merged = true;
                    while (merged && l2 >= 2) l2 = enc_pass(a.tab, t, l2, merged);
                    cnt[j] = l2;
                }
                sum += cnt[j];
            }
        }
        // block exclusive scan of the per-thread sums
        uint32_t incl = sum;
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t warp_base = 0, total = 0;
#pragma unroll
        for (int w = 0; w < ENC_THREADS / 32; w++) {
            uint32_t v = s_warp[w];
            if (w < (int)warp) warp_base += v;
            total += v;
        }
        if (tid == 0) s_base = lookback_base(a.status, tile, total);
        __syncthreads();
        const uint64_t base = s_base;
        uint64_t dst = base + warp_base + (incl - sum);
#pragma unroll
        for (int j = 0; j < ET_CPT; j++) {
            const uint32_t k = tid * ET_CPT + j;
            if (k < nc) {
                const uint32_t o = s_off[k], len = s_off[k + 1] - o, n = cnt[j];
                if (a.out_off) a.out_off[c0 + k] = dst;
                if (dst + n <= a.out_cap) {
                    if (len > ENC_SHORT_MAX) {
                        for (uint32_t i = 0; i < n; i++) a.out[dst + i] = a.scratch_a[o + i];
                    } else if (staged) {
                        for (uint32_t i = 0; i < n; i++) a.out[dst + i] = s_tok[o - b0 + i];
                    } else {
                        uint32_t t[ENC_SHORT_MAX];
                        for (uint32_t i = 0; i < len; i++) t[i] = __ldg(&a.bytes[o + i]);
                        uint32_t l2 = len;
                        bool merged = true;
                        while (merged && l2 >= 2) l2 = enc_pass(a.tab, t, l2, merged);
                        for (uint32_t i = 0; i < n; i++) a.out[dst + i] = t[i];
                    }
                } else if (n) {
                    *a.overflow = 1;
                }
                dst += n;
            }
        }
        if (tile == a.n_tiles - 1 && tid == 0) {
            *a.d_n_out = base + total;
            if (a.out_off) a.out_off[a.n_chunks] = base + total;
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

__global__ void __launch_bounds__(ENC_THREADS) k_decode_tiles(const DecArgs a) {
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_warp[ENC_THREADS / 32];
    __shared__ unsigned long long s_base;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_tile = atomicAdd(a.ticket, 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= a.n_tiles) return;
        const uint64_t k = (uint64_t)tile * ENC_THREADS + threadIdx.x;
        const uint8_t *src = nullptr;
        uint32_t len = 0;
        if (k < a.n_ids) {
            uint32_t id = __ldg(&a.ids[k]);
            int lo = 0, hi = (int)a.n_sp - 1, hit = -1; // special tokens override the vocabulary (Tokenizer.h:733)
            while (lo <= hi) {
                int mid = (lo + hi) >> 1;
                uint32_t v = __ldg(&a.sp_ids[mid]);
                if (v == id) {
                    hit = mid;
                    break;
                }
                if (v < id)
                    lo = mid + 1;
                else
                    hi = mid - 1;
            }
            if (hit >= 0) {
                uint32_t so = __ldg(&a.sp_off[hit]);
                src = a.sp_bytes + so;
                len = __ldg(&a.sp_off[hit + 1]) - so;
            } else if (id < a.vocab_size) {
                uint32_t vo = __ldg(&a.v_off[id]);
                src = a.v_bytes + vo;
                len = __ldg(&a.v_off[id + 1]) - vo;
            }
        }
        uint32_t incl = len;
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t warp_base = 0, total = 0;
        for (int w = 0; w < ENC_THREADS / 32; w++) {
            uint32_t v = s_warp[w];
            if (w < (int)warp) warp_base += v;
            total += v;
        }
        if (threadIdx.x == 0) s_base = lookback_base(a.status, tile, total);
        __syncthreads();
        const uint64_t dst = s_base + warp_base + (incl - len);
        if (a.out && dst + len <= a.out_cap)
            for (uint32_t i = 0; i < len; i++) a.out[dst + i] = __ldg(&src[i]);
        if (tile == a.n_tiles - 1 && threadIdx.x == 0) *a.d_n_out = s_base + total;
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
};

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
    MB_CUDA(cudaMalloc(&e->d_small, 16));
    MB_CUDA(cudaMalloc(&e->d_n_out, 8));
    MB_CUDA(cudaMalloc(&e->d_sp_ids, 4));
    MB_CUDA(cudaMalloc(&e->d_sp_off, 8));
    MB_CUDA(cudaMalloc(&e->d_sp_bytes, 1));
    *out = e;
    return MBPE_OK;
}

extern "C" void mbpe_encoder_destroy(mbpe_encoder *e) {
    if (!e) return;
    cudaSetDevice(e->device);
    void *ps[] = {e->d_slots, e->d_voff, e->d_vbytes, e->d_sp_ids, e->d_sp_off, e->d_sp_bytes, e->d_status, e->d_small,
                  e->d_n_out, e->d_long_list, e->d_scratch_a, e->d_scratch_b};
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
    if ((rc = ensure_status(e, (n_chunks + ENC_THREADS - 1) / ENC_THREADS + 1))) return rc;
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
    uint64_t n_tiles = (n_chunks + ET_CHUNKS - 1) / ET_CHUNKS;
    if (n_tiles >= 0xFFFFFFFFull) return set_error(MBPE_E_INVALID, "too many chunks in one batch");
    MB_CUDA(cudaMemsetAsync(e->d_small, 0, 16, st));
    if (n_chunks == 0) {
        MB_CUDA(cudaMemsetAsync(d_n_out, 0, 8, st));
        if (d_out_off) MB_CUDA(cudaMemsetAsync(d_out_off, 0, 8, st));
        return MBPE_OK;
    }
    MB_CUDA(cudaMemsetAsync(e->d_status, 0, n_tiles * 8, st));
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
    a.bytes = d_bytes;
    a.n_bytes_total = n_bytes;
    a.off = d_off;
    a.n_chunks = n_chunks;
    a.out = d_out;
    a.out_cap = out_cap;
    a.d_n_out = (unsigned long long *)d_n_out;
    a.out_off = d_out_off;
    a.status = e->d_status;
    a.ticket = e->d_small;
    a.n_tiles = (uint32_t)n_tiles;
    a.scratch_a = e->d_scratch_a;
    a.scratch_b = e->d_scratch_b;
    a.overflow = e->d_small + 2;
    unsigned g2 = (unsigned)std::min<uint64_t>(n_tiles, (uint64_t)e->sms * 8);
    k_encode_tiles<<<g2, ENC_THREADS, 0, st>>>(a);
    e->launches++;
    MB_CUDA(cudaGetLastError());
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

extern "C" int mbpe_decode(mbpe_encoder *e, const uint32_t *ids, uint64_t n_ids, uint8_t *out, uint64_t out_cap,
                           uint64_t *n_out) {
    if (!e || !n_out || (n_ids && !ids)) return set_error(MBPE_E_INVALID, "null argument");
    int rc = use_device(e->device);
    if (rc) return rc;
    *n_out = 0;
    if (n_ids == 0) return MBPE_OK;
    uint64_t n_tiles = (n_ids + ENC_THREADS - 1) / ENC_THREADS;
    if (n_tiles >= 0xFFFFFFFFull) return set_error(MBPE_E_INVALID, "too many ids in one call");
    if ((rc = ensure_status(e, n_tiles))) return rc;
    uint32_t *d_ids = nullptr;
    uint8_t *d_out = nullptr;
    MB_CUDA(cudaMalloc(&d_ids, n_ids * 4));
    MB_CUDA(cudaMemcpy(d_ids, ids, n_ids * 4, cudaMemcpyHostToDevice));
    DecArgs a{};
    a.ids = d_ids;
    a.n_ids = n_ids;
    a.v_off = e->d_voff;
    a.v_bytes = e->d_vbytes;
    a.vocab_size = e->vocab_size;
    a.sp_ids = e->d_sp_ids;
    a.sp_off = e->d_sp_off;
    a.sp_bytes = e->d_sp_bytes;
    a.n_sp = e->n_sp;
    a.d_n_out = e->d_n_out;
    a.status = e->d_status;
    a.ticket = e->d_small;
    a.n_tiles = (uint32_t)n_tiles;
    unsigned grid = (unsigned)std::min<uint64_t>(n_tiles, (uint64_t)e->sms * 8);
    // pass 1 sizes the output, pass 2 writes it (the caller may also stop after pass 1 with out == NULL)
    for (int pass = 0; pass < 2; pass++) {
        MB_CUDA(cudaMemset(e->d_small, 0, 16));
        MB_CUDA(cudaMemset(e->d_status, 0, n_tiles * 8));
        a.out = pass ? d_out : nullptr;
        a.out_cap = pass ? *n_out : 0;
        k_decode_tiles<<<grid, ENC_THREADS>>>(a);
        e->launches++;
        MB_CUDA(cudaGetLastError());
        if (pass == 0) {
            MB_CUDA(cudaMemcpy(n_out, e->d_n_out, 8, cudaMemcpyDeviceToHost));
            if (!out) break;
            if (*n_out > out_cap) {
                cudaFree(d_ids);
                return set_error(MBPE_E_CAPACITY, "out too small; *n_out holds the needed size");
            }
            MB_CUDA(cudaMalloc(&d_out, std::max<uint64_t>(*n_out, 1)));
        } else {
            MB_CUDA(cudaMemcpy(out, d_out, *n_out, cudaMemcpyDeviceToHost));
        }
    }
    cudaFree(d_ids);
    cudaFree(d_out);
    return MBPE_OK;
}
