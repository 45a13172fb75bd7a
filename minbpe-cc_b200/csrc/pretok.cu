// pretok.cu -- GPU pre-tokeniser for the GPT-4 / GPT-2 split patterns and GPU chunk dedup: the two host stages
// in front of the merge loop (Tokenizer.h:500-544 regex split; chunk -> count) moved next to it, so that
// text -> chunks -> unique chunks -> merge loop never leaves HBM.
//
//   k_pretok_mark     one thread per 32-byte window of text runs pretok_window() (pretok_core.cuh) and sets one bit per
//                     chunk start in a bitmap (1 bit per text byte). Windows are independent; overlap is idempotent.
//   k_bits_compact    bitmap -> ascending list of set-bit positions (= chunk offsets), single pass: popc per word,
//                     block scan, decoupled look-back for the tile base, positions staged in shared memory and stored
//                     as whole lines. Used twice: chunk starts, and first occurrences of unique chunks.
//   k_dedup_insert    every chunk into an open-addressed table keyed by its bytes (compared against the
//                     representative's bytes in the text itself): slot = {tag | smallest chunk index, count}.
//   k_dedup_first     table -> bitmap over chunk indices (bit c = chunk c is the first occurrence of its bytes).
//   k_dedup_emit      unique chunks in first-appearance order: weight, length; then token offsets (scan) and the
//                     bytes widened to u32 -- the trainer's input layout (train_cuda.cu), written in place on the GPU.
// Texts the matcher cannot take (malformed UTF-8, or a stretch without letters/blanks longer than the crawl limit)
// come back as MBPE_E_UNSUPPORTED and the caller uses the PCRE2 path.
// Host-facing entry points built on these: mbpe_pretok_corpus (text -> unique-chunk corpus on the device, upload
// overlapped with marking), mbpe_encode_text (text -> ids, three-stream pipeline over 64 MiB segments) and
// mbpe_encode_file (file -> .enc file in blocks with a reader and a writer thread).
#include <string.h>

#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"
#include "lookback.cuh"
#include "pretok_core.cuh"

namespace mbpe {

struct DevText {
    const uint8_t *p;
    __device__ __forceinline__ uint8_t operator[](uint64_t i) const { return __ldg(p + i); }
};

constexpr int PM_THREADS = 256;
constexpr uint32_t PT_WINDOW = 32; // bytes per thread = one bitmap word

// windows [w_first, ceil(len / 32)) of the subject text[0, len)
__global__ void __launch_bounds__(PM_THREADS) k_pretok_mark(const uint8_t *text, uint64_t len, const uint8_t *table,
                                                             uint32_t *bitmap, uint32_t *err, uint64_t max_crawl, uint32_t kind,
                                                             uint64_t w_first) {
    const uint64_t n_win = (len + PT_WINDOW - 1) / PT_WINDOW;
    PretokIn<DevText> in{DevText{text}, len, table, err, kind};
    for (uint64_t w = w_first + blockIdx.x * (uint64_t)PM_THREADS + threadIdx.x; w < n_win; w += (uint64_t)gridDim.x * PM_THREADS) {
        uint64_t cur = ~0ull;
        uint32_t bits = 0;
        pretok_window(in, w * PT_WINDOW, (w + 1) * PT_WINDOW, max_crawl, [&](uint64_t p) {
            const uint64_t wi = p >> 5;
            if (wi != cur) {
                if (bits) atomicOr(&bitmap[cur], bits);
                cur = wi;
                bits = 0;
            }
            bits |= 1u << (p & 31);
        });
        if (bits) atomicOr(&bitmap[cur], bits);
    }
}

// the same with the text cut into independent subjects by special-token occurrences (pretok_window_parts)
__global__ void __launch_bounds__(PM_THREADS) k_pretok_mark_parts(const uint8_t *text, uint64_t len, const uint8_t *table,
                                                                   uint32_t *bitmap, uint32_t *err, uint64_t max_crawl, uint32_t kind,
                                                                   const uint32_t *sp_b, const uint32_t *sp_e, uint32_t n_sp) {
    const uint64_t n_win = (len + PT_WINDOW - 1) / PT_WINDOW;
    PretokIn<DevText> in{DevText{text}, len, table, err, kind, 0};
    for (uint64_t w = blockIdx.x * (uint64_t)PM_THREADS + threadIdx.x; w < n_win; w += (uint64_t)gridDim.x * PM_THREADS) {
        uint64_t cur = ~0ull;
        uint32_t bits = 0;
        pretok_window_parts(in, sp_b, sp_e, n_sp, len, w * PT_WINDOW, (w + 1) * PT_WINDOW, max_crawl, [&](uint64_t p) {
            const uint64_t wi = p >> 5;
            if (wi != cur) {
                if (bits) atomicOr(&bitmap[cur], bits);
                cur = wi;
                bits = 0;
            }
            bits |= 1u << (p & 31);
        });
        if (bits) atomicOr(&bitmap[cur], bits);
    }
}

// ---------------------------------------------------------------------------------------------------------
// bitmap -> positions
// ---------------------------------------------------------------------------------------------------------
constexpr int BC_THREADS = 256, BC_WPT = 4, BC_WORDS = BC_THREADS * BC_WPT; // 1024 words = 32 Ki bits per tile
constexpr uint32_t BC_STAGE = 12288;                                        // positions staged per tile (48 KB)

struct BcSmem {
    uint32_t stage[BC_STAGE];
    uint32_t warp_sum[BC_THREADS / 32];
    uint32_t tile;
    unsigned long long base;
};

// out[k] = position of the k-th set bit (k counted over the whole bitmap); *n_out = number of set bits. Positions
// are 32-bit (bitmaps cover < 2^32 bits).
__global__ void __launch_bounds__(BC_THREADS) k_bits_compact(const uint32_t *bitmap, uint64_t n_words, uint32_t *out,
                                                              uint64_t out_cap, unsigned long long *status, uint32_t *ticket,
                                                              uint32_t n_tiles, unsigned long long *n_out, uint32_t *overflow) {
    extern __shared__ __align__(16) unsigned char bc_raw[];
    BcSmem &sm = *reinterpret_cast<BcSmem *>(bc_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (;;) {
        __syncthreads();
        if (tid == 0) sm.tile = atomicAdd(ticket, 1u);
        __syncthreads();
        const uint32_t tile = sm.tile;
        if (tile >= n_tiles) return;
        const uint64_t w0 = (uint64_t)tile * BC_WORDS + (uint64_t)tid * BC_WPT;
        uint32_t w[BC_WPT];
        if (w0 + BC_WPT <= n_words) {
            const uint4 q = __ldcs(reinterpret_cast<const uint4 *>(bitmap + w0));
            w[0] = q.x, w[1] = q.y, w[2] = q.z, w[3] = q.w;
        } else {
#pragma unroll
            for (int j = 0; j < BC_WPT; j++) w[j] = w0 + j < n_words ? __ldcs(bitmap + w0 + j) : 0u;
        }
        uint32_t sum = 0;
#pragma unroll
        for (int j = 0; j < BC_WPT; j++) sum += __popc(w[j]);
        uint32_t incl = sum;
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        if (lane == 31) sm.warp_sum[warp] = incl;
        __syncthreads();
        uint32_t warp_base = 0, total = 0;
#pragma unroll
        for (int k = 0; k < BC_THREADS / 32; k++) {
            const uint32_t v = sm.warp_sum[k];
            if (k < (int)warp) warp_base += v;
            total += v;
        }
        if (warp == 0) {
            const uint64_t b = lookback_base(status, tile, total);
            if (lane == 0) sm.base = b;
        }
        __syncthreads();
        const uint64_t base = sm.base;
        const bool fits = base + total <= out_cap;
        if (!fits && tid == 0) *overflow = 1;
        const bool via_smem = total <= BC_STAGE;
        uint32_t loc = warp_base + (incl - sum);
        if (fits) {
#pragma unroll
            for (int j = 0; j < BC_WPT; j++) {
                uint32_t bits = w[j];
                const uint32_t pos0 = (uint32_t)((w0 + j) << 5);
                while (bits) {
                    const uint32_t b = __ffs(bits) - 1;
                    bits &= bits - 1;
                    if (via_smem)
                        sm.stage[loc] = pos0 + b;
                    else
                        out[base + loc] = pos0 + b;
                    loc++;
                }
            }
            if (via_smem) {
                __syncthreads();
                for (uint32_t i = tid; i < total; i += BC_THREADS) __stcs(&out[base + i], sm.stage[i]);
            }
        }
        if (tile == n_tiles - 1 && tid == 0) *n_out = base + total;
    }
}

// ---------------------------------------------------------------------------------------------------------
// chunk dedup: unique chunk -> count, unique chunks in first-appearance order (host twin: chunker.cpp)
// ---------------------------------------------------------------------------------------------------------
constexpr unsigned long long DD_EMPTY = ~0ull;
constexpr uint32_t DD_MAX_PROBES = 1u << 14;

// The text may live in several resident segments (each < 4 GiB, so offsets stay 32-bit); chunk indices are global.
constexpr int DD_MAX_SEGS = 16;
struct DdSeg {
    const uint8_t *text;
    const uint32_t *off; // this segment's chunk offsets (n + 1), relative to its text
    uint64_t chunk0;     // global index of its first chunk
    uint64_t n_bytes;    // readable bytes of `text` (the 16-byte key loads never cross it)
    const uint32_t *weight; // multiplicity of each chunk of this segment (merging already deduplicated corpora); null = 1
};
struct DedupArgs {
    DdSeg seg[DD_MAX_SEGS];
    uint32_t n_segs;
    uint64_t n_chunks;
    unsigned long long *words; // slot: tag(31) << 32 | smallest chunk index with these bytes; DD_EMPTY = free
    uint32_t *counts;          // slot: occurrences
    uint32_t mask;
    uint32_t *used;     // claimed slots
    uint32_t *overflow; // probe limit hit: table too small
    uint32_t fast;      // every segment's text is 4-byte aligned: chunks of <= 16 bytes are read as five aligned words
};

__device__ __forceinline__ uint64_t dd_hash(const uint8_t *text, uint32_t o, uint32_t len) {
    uint64_t h = 0x9E3779B97F4A7C15ull ^ len;
    uint32_t i = 0;
    for (; i + 8 <= len; i += 8) {
        uint64_t v = 0;
#pragma unroll
        for (int q = 0; q < 8; q++) v |= (uint64_t)__ldg(text + o + i + q) << (8 * q);
        h = (h ^ v) * 0xff51afd7ed558ccdULL;
        h ^= h >> 29;
    }
    uint64_t v = 0;
    for (uint32_t q = 0; i + q < len; q++) v |= (uint64_t)__ldg(text + o + i + q) << (8 * q);
    h = (h ^ v) * 0xc4ceb9fe1a85ec53ULL;
    h ^= h >> 32;
    h *= 0xff51afd7ed558ccdULL;
    h ^= h >> 29;
    return h;
}

// dd_hash() of a chunk of <= 16 bytes given as two little-endian words (bytes past the chunk zeroed): same value
__device__ __forceinline__ uint64_t dd_hash16(uint64_t k0, uint64_t k1, uint32_t len) {
    uint64_t h = 0x9E3779B97F4A7C15ull ^ len;
    uint64_t v = k0;
    if (len >= 8) {
        h = (h ^ k0) * 0xff51afd7ed558ccdULL;
        h ^= h >> 29;
        v = k1;
        if (len == 16) {
            h = (h ^ k1) * 0xff51afd7ed558ccdULL;
            h ^= h >> 29;
            v = 0;
        }
    }
    h = (h ^ v) * 0xc4ceb9fe1a85ec53ULL;
    h ^= h >> 32;
    h *= 0xff51afd7ed558ccdULL;
    h ^= h >> 29;
    return h;
}
// the first 16 bytes of the chunk at o (len <= 16) as two words, bytes past the chunk zeroed. text is 4-byte
// aligned and o + 20 <= readable bytes: five aligned word loads in flight at once instead of a byte loop
__device__ __forceinline__ void dd_load16(const uint8_t *text, uint32_t o, uint32_t len, uint64_t &k0, uint64_t &k1) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(text + (o & ~3u));
    const uint32_t sh = (o & 3) * 8;
    const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2), w3 = __ldg(w + 3), w4 = __ldg(w + 4);
    const uint32_t v0 = __funnelshift_r(w0, w1, sh), v1 = __funnelshift_r(w1, w2, sh);
    const uint32_t v2 = __funnelshift_r(w2, w3, sh), v3 = __funnelshift_r(w3, w4, sh);
    k0 = ((uint64_t)v1 << 32) | v0;
    k1 = ((uint64_t)v3 << 32) | v2;
    if (len < 8) {
        k0 &= (1ull << (len * 8)) - 1;
        k1 = 0;
    } else if (len < 16) {
        k1 &= (1ull << ((len - 8) * 8)) - 1;
    }
}

__device__ __forceinline__ const DdSeg &dd_seg(const DedupArgs &a, uint64_t c) {
    uint32_t k = 0;
    while (k + 1 < a.n_segs && a.seg[k + 1].chunk0 <= c) k++;
    return a.seg[k];
}

__device__ __forceinline__ bool dd_equal(const DedupArgs &a, const uint8_t *text, uint32_t o, uint32_t len, uint32_t rep) {
    const DdSeg &rs = dd_seg(a, rep);
    const uint32_t ro = __ldg(rs.off + (rep - rs.chunk0));
    if (__ldg(rs.off + (rep - rs.chunk0) + 1) - ro != len) return false;
    for (uint32_t i = 0; i < len; i++)
        if (__ldg(text + o + i) != __ldg(rs.text + ro + i)) return false;
    return true;
}

__global__ void __launch_bounds__(256) k_dedup_insert(const DedupArgs a) {
    for (uint64_t c = blockIdx.x * 256ull + threadIdx.x; c < a.n_chunks; c += (uint64_t)gridDim.x * 256) {
        const DdSeg &sg = dd_seg(a, c);
        const uint32_t o = __ldg(sg.off + (c - sg.chunk0)), len = __ldg(sg.off + (c - sg.chunk0) + 1) - o;
        const bool fast = a.fast && len <= 16 && (uint64_t)o + 20 <= sg.n_bytes;
        const uint32_t wt = sg.weight ? __ldg(sg.weight + (c - sg.chunk0)) : 1u;
        uint64_t k0 = 0, k1 = 0;
        if (fast) dd_load16(sg.text, o, len, k0, k1);
        const uint64_t h = fast ? dd_hash16(k0, k1, len) : dd_hash(sg.text, o, len);
        const unsigned long long mine = ((h >> 33) << 32) | (uint32_t)c;
        uint32_t s = (uint32_t)h & a.mask;
        for (uint32_t probes = 0;; probes++) {
            if (probes > DD_MAX_PROBES) {
                *a.overflow = 1;
                break;
            }
            unsigned long long w = *((volatile unsigned long long *)&a.words[s]);
            if (w == DD_EMPTY) {
                w = atomicCAS(&a.words[s], DD_EMPTY, mine);
                if (w == DD_EMPTY) {
                    atomicAdd(a.used, 1u);
                    atomicAdd(&a.counts[s], wt);
                    break;
                }
            }
            if ((w >> 32) == (mine >> 32)) {
                const uint32_t rep = (uint32_t)w;
                bool same = rep == (uint32_t)c;
                if (!same && fast) { // both chunks as 16-byte keys when the representative allows it too
                    const DdSeg &rs = dd_seg(a, rep);
                    const uint32_t ro = __ldg(rs.off + (rep - rs.chunk0)), rl = __ldg(rs.off + (rep - rs.chunk0) + 1) - ro;
                    if (rl != len) {
                        same = false;
                    } else if ((uint64_t)ro + 20 <= rs.n_bytes) {
                        uint64_t r0, r1;
                        dd_load16(rs.text, ro, len, r0, r1);
                        same = r0 == k0 && r1 == k1;
                    } else {
                        same = dd_equal(a, sg.text, o, len, rep);
                    }
                } else if (!same) {
                    same = dd_equal(a, sg.text, o, len, rep);
                }
                if (same) {
                    if ((uint32_t)c < rep) atomicMin(&a.words[s], mine); // keep the FIRST occurrence as representative
                    atomicAdd(&a.counts[s], wt);
                    break;
                }
            }
            s = (s + 1) & a.mask;
        }
    }
}

// bit c of `first` = chunk c is the first occurrence of its bytes
__global__ void k_dedup_first(const unsigned long long *words, uint64_t n_slots, uint32_t *first) {
    for (uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; s < n_slots; s += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long w = words[s];
        if (w != DD_EMPTY) atomicOr(&first[(uint32_t)w >> 5], 1u << ((uint32_t)w & 31));
    }
}

// unique u = chunk first_idx[u]: weight, and its place in the token stream (exclusive scan of the lengths)
constexpr int DE_THREADS = 256;
__global__ void __launch_bounds__(DE_THREADS) k_dedup_emit(const DedupArgs a, const uint32_t *first_idx, uint32_t n_unique,
                                                            uint32_t *weight, unsigned long long *tok_off,
                                                            unsigned long long *status, uint32_t *ticket, uint32_t n_tiles) {
    __shared__ uint32_t warp_sum[DE_THREADS / 32];
    __shared__ uint32_t s_tile;
    __shared__ unsigned long long s_base;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_tile = atomicAdd(ticket, 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= n_tiles) return;
        const uint32_t u = tile * DE_THREADS + tid;
        uint32_t len = 0;
        if (u < n_unique) {
            const uint32_t c = __ldg(first_idx + u);
            const DdSeg &sg = dd_seg(a, c);
            const uint32_t o = __ldg(sg.off + (c - sg.chunk0));
            len = __ldg(sg.off + (c - sg.chunk0) + 1) - o;
            const uint64_t h = dd_hash(sg.text, o, len);
            uint32_t s = (uint32_t)h & a.mask;
            while ((uint32_t)a.words[s] != c || a.words[s] == DD_EMPTY) s = (s + 1) & a.mask; // the slot this chunk represents
            weight[u] = a.counts[s];
        }
        uint32_t incl = len;
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        if (lane == 31) warp_sum[warp] = incl;
        __syncthreads();
        uint32_t warp_base = 0, total = 0;
#pragma unroll
        for (int k = 0; k < DE_THREADS / 32; k++) {
            const uint32_t v = warp_sum[k];
            if (k < (int)warp) warp_base += v;
            total += v;
        }
        if (warp == 0) {
            const uint64_t b = lookback_base(status, tile, total);
            if (lane == 0) s_base = b;
        }
        __syncthreads();
        if (u < n_unique) tok_off[u] = s_base + warp_base + (incl - len);
        if (tile == n_tiles - 1 && tid == 0) tok_off[n_unique] = s_base + total;
    }
}

// bytes of the unique chunks widened to u32 tokens (Tokenizer.h:85-100), one warp per chunk
__global__ void __launch_bounds__(256) k_dedup_tokens(const DedupArgs a, const uint32_t *first_idx, uint32_t n_unique,
                                                       const unsigned long long *tok_off, uint32_t *tokens) {
    const uint32_t lane = threadIdx.x & 31;
    for (uint64_t u = (blockIdx.x * 256ull + threadIdx.x) >> 5; u < n_unique; u += ((uint64_t)gridDim.x * 256) >> 5) {
        const uint32_t c = __ldg(first_idx + u);
        const DdSeg &sg = dd_seg(a, c);
        const uint32_t o = __ldg(sg.off + (c - sg.chunk0)), len = __ldg(sg.off + (c - sg.chunk0) + 1) - o;
        const unsigned long long t0 = tok_off[u];
        for (uint32_t i = lane; i < len; i += 32) tokens[t0 + i] = __ldg(sg.text + o + i);
    }
}

} // namespace mbpe

using namespace mbpe;

#include <chrono>
static double pt_now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static bool pt_debug() {
    static const bool on = getenv("MBPE_DEBUG") != nullptr;
    return on;
}

struct mbpe_pretok {
    int device = 0, sms = 148;
    uint8_t *d_table = nullptr;
    uint32_t *d_bitmap = nullptr; // 1 bit per text byte
    uint64_t bitmap_words = 0;
    unsigned long long *d_status = nullptr;
    uint64_t status_cap = 0;
    uint32_t *d_small = nullptr;          // [0] err, [1] ticket, [2] overflow
    unsigned long long *d_count = nullptr; // set-bit count
    uint64_t max_crawl = 1u << 16;
    uint32_t kind = PT_GPT4; // which built-in pattern, mbpe_pretok_select
    uint64_t launches = 0;
    uint8_t *d_seg_text = nullptr; // segment buffers of mbpe_pretok_corpus, kept between calls
    uint32_t *d_seg_off = nullptr;
    uint64_t *d_seg_n = nullptr;
    uint64_t seg_cap = 0;
    cudaStream_t st_in = nullptr, st_c = nullptr, st_out = nullptr; // mbpe_encode_text pipeline
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    uint8_t *d_pipe_text[2] = {nullptr, nullptr};
    uint32_t *d_pipe_ids[2] = {nullptr, nullptr}, *d_pipe_off = nullptr;
    uint64_t pipe_cap = 0;
    uint64_t enc_seg_bytes = 64ull << 20;
    void *scratch[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}; // work buffers, grow-only, kept between calls
    uint64_t scratch_cap[6] = {0, 0, 0, 0, 0, 0}; // 0-3 dedup, 4-5 special-token occurrences of a segment
    std::vector<uint8_t> h_table;   // host copy: segment boundaries are chosen at cuts (pretok_core.cuh)
    uint64_t seg_bytes = 1ull << 30; // host text is brought over in segments of about this size
};

extern "C" int mbpe_pretok_create(int device, mbpe_pretok **out) {
    if (!out) return set_error(MBPE_E_INVALID, "null argument");
    *out = nullptr;
    std::vector<uint8_t> table(PT_TABLE_BYTES);
    int rc = mbpe_pretok_class_table(table.data());
    if (rc) return rc;
    if ((rc = use_device(device))) return rc;
    mbpe_pretok *p = new mbpe_pretok();
    p->device = device;
    p->sms = sm_count(device);
    if (const char *mc = getenv("MBPE_PRETOK_MAX_CRAWL")) p->max_crawl = strtoull(mc, nullptr, 10);
    if (const char *sb = getenv("MBPE_PRETOK_SEG_BYTES")) p->seg_bytes = std::max<uint64_t>(64, strtoull(sb, nullptr, 10));
    p->seg_bytes = std::min<uint64_t>(p->seg_bytes, 3ull << 30);
    if (const char *sb = getenv("MBPE_ENCODE_SEG_BYTES")) p->enc_seg_bytes = std::max<uint64_t>(64, strtoull(sb, nullptr, 10));
    p->enc_seg_bytes = std::min<uint64_t>(p->enc_seg_bytes, 3ull << 30);
    p->h_table = table;
    MB_CUDA(cudaMalloc(&p->d_table, PT_TABLE_BYTES));
    MB_CUDA(cudaMemcpy(p->d_table, table.data(), PT_TABLE_BYTES, cudaMemcpyHostToDevice));
    MB_CUDA(cudaMalloc(&p->d_small, 16));
    MB_CUDA(cudaMalloc(&p->d_count, 8));
    MB_CUDA(cudaFuncSetAttribute(k_bits_compact, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BcSmem)));
    *out = p;
    return MBPE_OK;
}

// which of the two built-in patterns the handle matches (Tokenizer.h:59-60); anything else has no device matcher
extern "C" int mbpe_pretok_select(mbpe_pretok *p, const char *pattern) {
    if (!p || !pattern) return set_error(MBPE_E_INVALID, "null argument");
    if (!strcmp(pattern, mbpe_gpt4_split_pattern()))
        p->kind = PT_GPT4;
    else if (!strcmp(pattern, mbpe_gpt2_split_pattern()))
        p->kind = PT_GPT2;
    else
        return set_error(MBPE_E_UNSUPPORTED, "only the GPT-2 and GPT-4 split patterns have a device matcher");
    return MBPE_OK;
}

extern "C" void mbpe_pretok_destroy(mbpe_pretok *p) {
    if (!p) return;
    cudaSetDevice(p->device);
    cudaFree(p->d_table);
    cudaFree(p->d_bitmap);
    cudaFree(p->d_status);
    cudaFree(p->d_small);
    cudaFree(p->d_count);
    cudaFree(p->d_seg_text);
    cudaFree(p->d_seg_off);
    cudaFree(p->d_seg_n);
    for (void *q : p->scratch) cudaFree(q);
    for (int i = 0; i < 2; i++) {
        cudaFree(p->d_pipe_text[i]);
        cudaFree(p->d_pipe_ids[i]);
        if (p->ev_in[i]) cudaEventDestroy(p->ev_in[i]);
        if (p->ev_out[i]) cudaEventDestroy(p->ev_out[i]);
    }
    cudaFree(p->d_pipe_off);
    for (cudaStream_t q : {p->st_in, p->st_c, p->st_out})
        if (q) cudaStreamDestroy(q);
    delete p;
}

namespace mbpe {
// grow-only work buffer k of the handle (driver allocations of this size cost milliseconds to hundreds of them)
static void *pt_scratch(mbpe_pretok *p, int k, uint64_t bytes) {
    if (bytes > p->scratch_cap[k]) {
        cudaFree(p->scratch[k]);
        p->scratch[k] = nullptr;
        p->scratch_cap[k] = 0;
        const uint64_t want = bytes + bytes / 8 + 256;
        if (cudaMalloc(&p->scratch[k], want) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        p->scratch_cap[k] = want;
    }
    return p->scratch[k];
}

// positions of the set bits of bitmap[0 .. n_words) into d_out (ascending); count left in p->d_count
int bits_compact(mbpe_pretok *p, const uint32_t *d_bitmap, uint64_t n_words, uint32_t *d_out, uint64_t out_cap, cudaStream_t st) {
    const uint64_t n_tiles = (n_words + BC_WORDS - 1) / BC_WORDS;
    if (n_tiles + 1 > p->status_cap) {
        cudaFree(p->d_status);
        p->status_cap = n_tiles + n_tiles / 4 + 64;
        MB_CUDA(cudaMalloc(&p->d_status, p->status_cap * 8));
    }
    MB_CUDA(cudaMemsetAsync(p->d_status, 0, std::max<uint64_t>(n_tiles, 1) * 8, st));
    MB_CUDA(cudaMemsetAsync(p->d_small + 1, 0, 8, st)); // ticket, overflow
    MB_CUDA(cudaMemsetAsync(p->d_count, 0, 8, st));
    if (n_tiles == 0) return MBPE_OK;
    const unsigned grid = (unsigned)std::min<uint64_t>(n_tiles, (uint64_t)p->sms * 4);
    k_bits_compact<<<grid, BC_THREADS, sizeof(BcSmem), st>>>(d_bitmap, n_words, d_out, out_cap, p->d_status, p->d_small + 1,
                                                            (uint32_t)n_tiles, p->d_count, p->d_small + 2);
    p->launches++;
    MB_CUDA(cudaGetLastError());
    return MBPE_OK;
}
} // namespace mbpe

namespace mbpe {
static int split_begin(mbpe_pretok *p, uint64_t len, uint64_t *n_words, cudaStream_t st) {
    *n_words = (len + 31) / 32;
    if (*n_words > p->bitmap_words) {
        cudaFree(p->d_bitmap);
        p->bitmap_words = *n_words + *n_words / 8 + 1024;
        MB_CUDA(cudaMalloc(&p->d_bitmap, p->bitmap_words * 4));
    }
    MB_CUDA(cudaMemsetAsync(p->d_bitmap, 0, *n_words * 4, st));
    MB_CUDA(cudaMemsetAsync(p->d_small, 0, 16, st));
    return MBPE_OK;
}
// marks the chunk starts of windows [w_first, ...) of the subject d_text[0, subject_len)
static int split_mark(mbpe_pretok *p, const uint8_t *d_text, uint64_t subject_len, uint64_t w_first, cudaStream_t st) {
    const uint64_t n_win = (subject_len + 31) / 32 - w_first;
    const unsigned grid = (unsigned)std::min<uint64_t>((n_win + PM_THREADS - 1) / PM_THREADS, (uint64_t)p->sms * 32);
    if (grid == 0) return MBPE_OK;
    k_pretok_mark<<<grid, PM_THREADS, 0, st>>>(d_text, subject_len, p->d_table, p->d_bitmap, p->d_small, p->max_crawl, p->kind, w_first);
    p->launches++;
    MB_CUDA(cudaGetLastError());
    return MBPE_OK;
}
// bitmap -> offsets, error flags, end offset
static int split_finish(mbpe_pretok *p, uint64_t len, uint64_t n_words, uint32_t *d_off_out, uint64_t off_cap, uint64_t *n_chunks,
                        cudaStream_t st) {
    int rc = bits_compact(p, p->d_bitmap, n_words, d_off_out, off_cap - 1, st);
    if (rc) return rc;
    uint32_t small[4];
    unsigned long long count = 0;
    MB_CUDA(cudaMemcpyAsync(small, p->d_small, 16, cudaMemcpyDeviceToHost, st));
    MB_CUDA(cudaMemcpyAsync(&count, p->d_count, 8, cudaMemcpyDeviceToHost, st));
    MB_CUDA(cudaStreamSynchronize(st));
    if (small[0] & PT_ERR_UTF8) return set_error(MBPE_E_UNSUPPORTED, "text is not well-formed UTF-8: use the PCRE2 path");
    if (small[0] & PT_ERR_LONG)
        return set_error(MBPE_E_UNSUPPORTED, "a stretch without letters or blanks exceeds the crawl limit: use the PCRE2 path");
    if (small[2]) {
        *n_chunks = count;
        return set_error(MBPE_E_CAPACITY, "offset buffer too small");
    }
    const uint32_t end = (uint32_t)len;
    MB_CUDA(cudaMemcpyAsync(d_off_out + count, &end, 4, cudaMemcpyHostToDevice, st));
    MB_CUDA(cudaStreamSynchronize(st));
    *n_chunks = count;
    return MBPE_OK;
}
} // namespace mbpe

// d_off_out: u32[off_cap]; on success holds n_chunks + 1 offsets (the last one = len)
extern "C" int mbpe_pretok_split_device(mbpe_pretok *p, const uint8_t *d_text, uint64_t len, uint32_t *d_off_out,
                                        uint64_t off_cap, uint64_t *n_chunks, void *stream) {
    if (!p || !n_chunks || !d_off_out || (len && !d_text)) return set_error(MBPE_E_INVALID, "null argument");
    if (len >= (1ull << 32) - 64) return set_error(MBPE_E_INVALID, "a device batch must be < 4 GiB of text");
    int rc = use_device(p->device);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    *n_chunks = 0;
    if (off_cap < 1) return set_error(MBPE_E_CAPACITY, "offset buffer too small");
    if (len == 0) {
        MB_CUDA(cudaMemsetAsync(d_off_out, 0, 4, st));
        MB_CUDA(cudaStreamSynchronize(st));
        return MBPE_OK;
    }
    uint64_t n_words = 0;
    if ((rc = split_begin(p, len, &n_words, st))) return rc;
    if ((rc = split_mark(p, d_text, len, 0, st))) return rc;
    return split_finish(p, len, n_words, d_off_out, off_cap, n_chunks, st);
}

// The same for a text with special tokens (Tokenizer.h:605-650): d_sp_begin / d_sp_end = the occurrences of special
// tokens in text order (device arrays, offsets into d_text). Every ordinary part is split as a subject of its own
// and every occurrence becomes one chunk.
extern "C" int mbpe_pretok_split_device_parts(mbpe_pretok *p, const uint8_t *d_text, uint64_t len, const uint32_t *d_sp_begin,
                                              const uint32_t *d_sp_end, uint32_t n_sp, uint32_t *d_off_out, uint64_t off_cap,
                                              uint64_t *n_chunks, void *stream) {
    if (n_sp == 0) return mbpe_pretok_split_device(p, d_text, len, d_off_out, off_cap, n_chunks, stream);
    if (!p || !n_chunks || !d_off_out || !d_text || !d_sp_begin || !d_sp_end) return set_error(MBPE_E_INVALID, "null argument");
    if (len >= (1ull << 32) - 64) return set_error(MBPE_E_INVALID, "a device batch must be < 4 GiB of text");
    int rc = use_device(p->device);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    *n_chunks = 0;
    if (off_cap < 1 || len == 0) return set_error(MBPE_E_INVALID, "empty text with special occurrences");
    uint64_t n_words = 0;
    if ((rc = split_begin(p, len, &n_words, st))) return rc;
    const unsigned grid = (unsigned)std::min<uint64_t>((n_words + PM_THREADS - 1) / PM_THREADS, (uint64_t)p->sms * 32);
    k_pretok_mark_parts<<<grid, PM_THREADS, 0, st>>>(d_text, len, p->d_table, p->d_bitmap, p->d_small, p->max_crawl, p->kind,
                                                    d_sp_begin, d_sp_end, n_sp);
    p->launches++;
    MB_CUDA(cudaGetLastError());
    return split_finish(p, len, n_words, d_off_out, off_cap, n_chunks, st);
}

// host text in, host offsets out: the GPU counterpart of mbpe_split for the GPT-4 pattern
extern "C" int mbpe_pretok_split(mbpe_pretok *p, const uint8_t *text, uint64_t len, uint64_t *off_out, uint64_t off_cap,
                                 uint64_t *n_chunks) {
    if (!p || !n_chunks || (len && !text)) return set_error(MBPE_E_INVALID, "null argument");
    int rc = use_device(p->device);
    if (rc) return rc;
    uint8_t *d_text = nullptr;
    uint32_t *d_off = nullptr;
    const uint64_t cap = len + 2;
    MB_CUDA(cudaMalloc(&d_text, std::max<uint64_t>(len, 1)));
    if (cudaMalloc(&d_off, cap * 4) != cudaSuccess) {
        cudaFree(d_text);
        return set_error(MBPE_E_CUDA, "out of device memory");
    }
    cudaError_t ce = cudaMemcpy(d_text, text, len, cudaMemcpyHostToDevice);
    rc = ce == cudaSuccess ? mbpe_pretok_split_device(p, d_text, len, d_off, cap, n_chunks, nullptr)
                           : cuda_fail(ce, "H2D text", __FILE__, __LINE__);
    if (rc == MBPE_OK && off_out) {
        if (*n_chunks + 1 > off_cap) {
            rc = set_error(MBPE_E_CAPACITY, "off_out too small");
        } else {
            std::vector<uint32_t> h(*n_chunks + 1);
            ce = cudaMemcpy(h.data(), d_off, h.size() * 4, cudaMemcpyDeviceToHost);
            if (ce != cudaSuccess) rc = cuda_fail(ce, "D2H offsets", __FILE__, __LINE__);
            for (size_t i = 0; i < h.size(); i++) off_out[i] = h[i];
        }
    }
    cudaFree(d_text);
    cudaFree(d_off);
    return rc;
}

// ---------------------------------------------------------------------------------------------------------
// C ABI: dedup of resident chunks, device corpus
// ---------------------------------------------------------------------------------------------------------
// The corpus buffers come from the device's stream-ordered pool (release threshold = keep everything): allocating and
// freeing them costs microseconds; cudaMalloc / cudaFree of the same sizes showed 100 - 600 ms outliers.
static void pool_keep_everything(int device) {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
}

extern "C" void mbpe_device_corpus_free(mbpe_device_corpus *c) {
    if (!c) return;
    if (c->d_tokens || c->d_off || c->d_weight) cudaSetDevice(c->device);
    if (c->d_tokens) cudaFreeAsync(c->d_tokens, nullptr);
    if (c->d_off) cudaFreeAsync(c->d_off, nullptr);
    if (c->d_weight) cudaFreeAsync(c->d_weight, nullptr);
    memset(c, 0, sizeof *c);
}

static int dedup_segments_impl(mbpe_pretok *p, const uint8_t *const *d_texts, const uint64_t *seg_bytes,
                               const uint32_t *const *d_offs, const uint32_t *const *d_weights, const uint64_t *seg_chunks,
                               uint32_t n_segs, mbpe_device_corpus *out, void *stream) {
    if (!p || !out || (n_segs && (!d_texts || !seg_bytes || !d_offs || !seg_chunks))) return set_error(MBPE_E_INVALID, "null argument");
    if (n_segs > (uint32_t)DD_MAX_SEGS) return set_error(MBPE_E_INVALID, "too many text segments");
    uint64_t n_chunks = 0;
    for (uint32_t k = 0; k < n_segs; k++) n_chunks += seg_chunks[k];
    if (n_chunks >= (1ull << 32) - 64) return set_error(MBPE_E_INVALID, "too many chunks (>= 2^32)");
    int rc = use_device(p->device);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    memset(out, 0, sizeof *out);
    out->n_chunks = n_chunks;
    out->device = p->device;
    pool_keep_everything(p->device);
    if (n_chunks == 0) {
        MB_CUDA(cudaMallocAsync(&out->d_tokens, 4, st));
        MB_CUDA(cudaMallocAsync(&out->d_off, 8, st));
        MB_CUDA(cudaMallocAsync(&out->d_weight, 4, st));
        MB_CUDA(cudaMemsetAsync(out->d_off, 0, 8, st));
        MB_CUDA(cudaStreamSynchronize(st));
        return MBPE_OK;
    }
    uint64_t slots = 1u << 16; // raw text repeats its chunks ~100x; corpora that are already deduplicated hardly at all
    while (slots < (d_weights ? n_chunks * 2 : n_chunks / 16) && slots < (1ull << 27)) slots <<= 1;
    unsigned long long *d_words = nullptr;
    uint32_t *d_counts = nullptr, *d_first = nullptr, *d_first_idx = nullptr;
    auto cleanup = [&]() {};
    uint32_t used = 0;
    DedupArgs a{};
    for (;;) {
        d_words = (unsigned long long *)pt_scratch(p, 0, slots * 8);
        d_counts = (uint32_t *)pt_scratch(p, 1, slots * 4);
        if (!d_words || !d_counts) return set_error(MBPE_E_CUDA, "out of device memory");
        MB_CUDA(cudaMemsetAsync(d_words, 0xFF, slots * 8, st));
        MB_CUDA(cudaMemsetAsync(d_counts, 0, slots * 4, st));
        MB_CUDA(cudaMemsetAsync(p->d_small, 0, 16, st));
        a = DedupArgs{};
        a.fast = 1;
        for (uint32_t k = 0, c0 = 0; k < n_segs; c0 += (uint32_t)seg_chunks[k], k++) {
            a.seg[k] = DdSeg{d_texts[k], d_offs[k], c0, seg_bytes[k], d_weights ? d_weights[k] : nullptr};
            if (((uintptr_t)d_texts[k] & 3) != 0) a.fast = 0;
        }
        a.n_segs = n_segs;
        a.n_chunks = n_chunks;
        a.words = d_words;
        a.counts = d_counts;
        a.mask = (uint32_t)(slots - 1);
        a.used = p->d_small + 3;
        a.overflow = p->d_small + 2;
        const unsigned grid = (unsigned)std::min<uint64_t>((n_chunks + 255) / 256, (uint64_t)p->sms * 16);
        k_dedup_insert<<<grid, 256, 0, st>>>(a);
        p->launches++;
        uint32_t small[4];
        MB_CUDA(cudaMemcpyAsync(small, p->d_small, 16, cudaMemcpyDeviceToHost, st));
        MB_CUDA(cudaStreamSynchronize(st));
        used = small[3];
        if (!small[2] && (uint64_t)used * 2 <= slots) break;
        if (slots >= (1ull << 31)) return set_error(MBPE_E_CUDA, "dedup table cannot grow further");
        slots <<= 2; // too full for short probe chains: redo in a larger table
    }
    const uint64_t first_words = (n_chunks + 31) / 32;
    d_first = (uint32_t *)pt_scratch(p, 2, first_words * 4);
    d_first_idx = (uint32_t *)pt_scratch(p, 3, ((uint64_t)used + 1) * 4);
    if (!d_first || !d_first_idx) return set_error(MBPE_E_CUDA, "out of device memory");
    MB_CUDA(cudaMemsetAsync(d_first, 0, first_words * 4, st));
    k_dedup_first<<<(unsigned)std::min<uint64_t>((slots + 255) / 256, (uint64_t)p->sms * 16), 256, 0, st>>>(d_words, slots, d_first);
    p->launches++;
    if ((rc = bits_compact(p, d_first, first_words, d_first_idx, used, st))) {
        cleanup();
        return rc;
    }
    out->n_unique = used;
    if (cudaMallocAsync(&out->d_off, ((uint64_t)used + 1) * 8, st) != cudaSuccess || cudaMallocAsync(&out->d_weight, (uint64_t)used * 4, st) != cudaSuccess) {
        cleanup();
        mbpe_device_corpus_free(out);
        return set_error(MBPE_E_CUDA, "out of device memory");
    }
    const uint32_t n_tiles = (used + DE_THREADS - 1) / DE_THREADS;
    if ((uint64_t)n_tiles + 1 > p->status_cap) {
        cudaFree(p->d_status);
        p->status_cap = n_tiles + 64;
        MB_CUDA(cudaMalloc(&p->d_status, p->status_cap * 8));
    }
    MB_CUDA(cudaMemsetAsync(p->d_status, 0, (uint64_t)n_tiles * 8, st));
    MB_CUDA(cudaMemsetAsync(p->d_small + 1, 0, 4, st));
    k_dedup_emit<<<std::min<unsigned>(n_tiles, p->sms * 4), DE_THREADS, 0, st>>>(
        a, d_first_idx, used, out->d_weight, (unsigned long long *)out->d_off, p->d_status, p->d_small + 1, n_tiles);
    p->launches++;
    unsigned long long n_tokens = 0;
    MB_CUDA(cudaMemcpyAsync(&n_tokens, out->d_off + used, 8, cudaMemcpyDeviceToHost, st));
    MB_CUDA(cudaStreamSynchronize(st));
    out->n_tokens = n_tokens;
    if (cudaMallocAsync(&out->d_tokens, std::max<uint64_t>(n_tokens, 1) * 4, st) != cudaSuccess) {
        cleanup();
        mbpe_device_corpus_free(out);
        return set_error(MBPE_E_CUDA, "out of device memory");
    }
    k_dedup_tokens<<<std::min<unsigned>((used + 7) / 8, p->sms * 16), 256, 0, st>>>(
        a, d_first_idx, used, (const unsigned long long *)out->d_off, out->d_tokens);
    p->launches++;
    cudaError_t ce = cudaStreamSynchronize(st);
    cleanup();
    if (ce != cudaSuccess) {
        mbpe_device_corpus_free(out);
        return cuda_fail(ce, "dedup kernels", __FILE__, __LINE__);
    }
    return MBPE_OK;
}

extern "C" int mbpe_pretok_dedup_segments(mbpe_pretok *p, const uint8_t *const *d_texts, const uint64_t *seg_bytes,
                                          const uint32_t *const *d_offs, const uint64_t *seg_chunks, uint32_t n_segs,
                                          mbpe_device_corpus *out, void *stream) {
    return dedup_segments_impl(p, d_texts, seg_bytes, d_offs, nullptr, seg_chunks, n_segs, out, stream);
}

// Several already deduplicated chunk lists (the ranks of a sharded train front end, in text order) into one: same
// kernels, every chunk counting with its weight. First-appearance order of the result = order over the whole text.
extern "C" int mbpe_pretok_merge_corpora(mbpe_pretok *p, const uint8_t *const *d_texts, const uint64_t *seg_bytes,
                                         const uint32_t *const *d_offs, const uint32_t *const *d_weights,
                                         const uint64_t *seg_chunks, uint32_t n_segs, mbpe_device_corpus *out, void *stream) {
    if (n_segs && !d_weights) return set_error(MBPE_E_INVALID, "null argument");
    return dedup_segments_impl(p, d_texts, seg_bytes, d_offs, d_weights, seg_chunks, n_segs, out, stream);
}

extern "C" int mbpe_pretok_dedup_device(mbpe_pretok *p, const uint8_t *d_text, uint64_t len, const uint32_t *d_off,
                                        uint64_t n_chunks, mbpe_device_corpus *out, void *stream) {
    if (!d_off || (len && !d_text)) return set_error(MBPE_E_INVALID, "null argument");
    return mbpe_pretok_dedup_segments(p, &d_text, &len, &d_off, &n_chunks, 1, out, stream);
}

extern "C" int mbpe_device_corpus_download(const mbpe_device_corpus *c, uint32_t *tokens, uint64_t *off, uint32_t *weight) {
    if (!c) return set_error(MBPE_E_INVALID, "null argument");
    int rc = use_device(c->device);
    if (rc) return rc;
    if (tokens) MB_CUDA(cudaMemcpy(tokens, c->d_tokens, c->n_tokens * 4, cudaMemcpyDeviceToHost));
    if (off) MB_CUDA(cudaMemcpy(off, c->d_off, (c->n_unique + 1) * 8, cudaMemcpyDeviceToHost));
    if (weight) MB_CUDA(cudaMemcpy(weight, c->d_weight, c->n_unique * 4, cudaMemcpyDeviceToHost));
    return MBPE_OK;
}

namespace mbpe {
struct HostText {
    const uint8_t *p;
    uint8_t operator[](uint64_t i) const { return p[i]; }
};

// Host text is brought over in segments that end at cuts: a cut is a match start whatever precedes it, and the match
// that ends there is decided by "the next code point is not of my class", which the end of a segment answers the same
// way -- so segments are split independently and the chunk lists concatenate. Returns false if no cut is near.
static bool plan_segments(const mbpe_pretok *pt, const uint8_t *text, uint64_t len, uint64_t seg_bytes, std::vector<uint64_t> &bounds) {
    bounds.assign(1, 0);
    uint32_t err = 0;
    PretokIn<HostText> in{HostText{text}, len, pt->h_table.data(), &err, pt->kind};
    while (len - bounds.back() > seg_bytes + seg_bytes / 4) {
        const uint64_t lo = bounds.back() + seg_bytes / 2, hi = bounds.back() + seg_bytes;
        uint64_t cut = 0;
        bool bad = false;
        for (uint64_t q = hi; q > lo && !cut && !bad; q--) {
            if ((text[q] & 0xC0) == 0x80) continue;
            const PtCp prev = pt_before(in, q, bad), cur = pt_at(in, q, bad);
            if (!bad && pt_is_cut(prev.cls, cur)) cut = q;
        }
        if (!cut) return false;
        bounds.push_back(cut);
    }
    bounds.push_back(len);
    return true;
}
} // namespace mbpe

static int ensure_segment_buffers(mbpe_pretok *p, uint64_t max_seg) {
    if (max_seg <= p->seg_cap && p->d_seg_n) return MBPE_OK;
    cudaFree(p->d_seg_text);
    cudaFree(p->d_seg_off);
    p->d_seg_text = nullptr, p->d_seg_off = nullptr;
    p->seg_cap = 0;
    const uint64_t cap = max_seg + max_seg / 16 + 64;
    if (cudaMalloc(&p->d_seg_text, cap) != cudaSuccess || cudaMalloc(&p->d_seg_off, (cap + 2) * 4) != cudaSuccess ||
        (!p->d_seg_n && cudaMalloc(&p->d_seg_n, 8) != cudaSuccess)) {
        cudaGetLastError();
        return set_error(MBPE_E_CUDA, "out of device memory");
    }
    p->seg_cap = cap;
    return MBPE_OK;
}

// One segment of host text: its pieces (128 MiB, cut at matcher cuts) are uploaded on one stream while the pieces
// already on the device are marked on another -- a piece's subject ends at its own end, which is a cut, so the marks
// are those of the whole segment; then one compaction of the bitmap.
static int ensure_pipe(mbpe_pretok *p, uint64_t max_seg);
static int upload_and_split(mbpe_pretok *p, const uint8_t *text, uint64_t len, uint8_t *d_text, uint32_t *d_off, uint64_t off_cap,
                            uint64_t *n_chunks) {
    *n_chunks = 0;
    std::vector<uint64_t> pb;
    const uint64_t piece = 128ull << 20;
    if (len < 2 * piece || ensure_pipe(p, 0) != MBPE_OK || !plan_segments(p, text, len, piece, pb)) {
        cudaError_t ce = cudaMemcpy(d_text, text, len, cudaMemcpyHostToDevice);
        if (ce != cudaSuccess) return cuda_fail(ce, "H2D text", __FILE__, __LINE__);
        return mbpe_pretok_split_device(p, d_text, len, d_off, off_cap, n_chunks, nullptr);
    }
    uint64_t n_words = 0;
    int rc = split_begin(p, len, &n_words, p->st_c);
    for (size_t k = 0; k + 1 < pb.size() && rc == MBPE_OK; k++) {
        const uint64_t b = pb[k], e = pb[k + 1];
        cudaError_t ce = cudaMemcpyAsync(d_text + b, text + b, e - b, cudaMemcpyHostToDevice, p->st_in);
        if (ce == cudaSuccess) ce = cudaEventRecord(p->ev_in[k & 1], p->st_in);
        if (ce == cudaSuccess) ce = cudaStreamWaitEvent(p->st_c, p->ev_in[k & 1], 0);
        if (ce != cudaSuccess) {
            rc = cuda_fail(ce, "H2D text", __FILE__, __LINE__);
            break;
        }
        rc = split_mark(p, d_text, e, b / 32, p->st_c); // windows from the one that holds b; the subject ends at e
    }
    if (rc == MBPE_OK) rc = split_finish(p, len, n_words, d_off, off_cap, n_chunks, p->st_c);
    cudaStreamSynchronize(p->st_in);
    cudaStreamSynchronize(p->st_c);
    return rc;
}

// text in host memory -> unique chunks resident on the device (split + dedup), the front end of Tokenizer::train
extern "C" int mbpe_pretok_corpus(mbpe_pretok *p, const uint8_t *text, uint64_t len, mbpe_device_corpus *out) {
    if (!p || !out || (len && !text)) return set_error(MBPE_E_INVALID, "null argument");
    int rc = use_device(p->device);
    if (rc) return rc;
    std::vector<uint64_t> bounds;
    if (!plan_segments(p, text, len, p->seg_bytes, bounds))
        return set_error(MBPE_E_UNSUPPORTED, "no safe segment boundary found (malformed UTF-8 or no letters/blanks): use the PCRE2 path");
    const uint32_t n_segs = (uint32_t)bounds.size() - 1;
    if (n_segs > (uint32_t)DD_MAX_SEGS) return set_error(MBPE_E_UNSUPPORTED, "text too large for the resident pre-tokeniser");
    std::vector<uint8_t *> d_text(n_segs, nullptr);
    std::vector<uint32_t *> d_off(n_segs, nullptr);
    std::vector<uint64_t> n_chunks(n_segs, 0), seg_len(n_segs, 0);
    for (uint32_t k = 0; k < n_segs; k++) seg_len[k] = bounds[k + 1] - bounds[k];
    auto cleanup = [&]() {
        if (n_segs == 1) return;
        for (auto q : d_text) cudaFree(q);
        for (auto q : d_off) cudaFree(q);
    };
    const double t_start = pt_now();
    double t_h2d = 0, t_split = 0;
    const bool cached = n_segs == 1; // the usual case: one resident segment, buffers kept in the handle
    if (cached) {
        if ((rc = ensure_segment_buffers(p, len))) return rc;
        d_text[0] = p->d_seg_text;
        d_off[0] = p->d_seg_off;
    }
    for (uint32_t k = 0; k < n_segs && rc == MBPE_OK; k++) {
        const uint64_t b = bounds[k], n = bounds[k + 1] - b;
        if (!cached && (cudaMalloc(&d_text[k], std::max<uint64_t>(n, 1)) != cudaSuccess || cudaMalloc(&d_off[k], (n + 2) * 4) != cudaSuccess)) {
            rc = set_error(MBPE_E_CUDA, "out of device memory");
            break;
        }
        double t0 = pt_now();
        rc = upload_and_split(p, text + b, n, d_text[k], d_off[k], n + 2, &n_chunks[k]);
        t_split += pt_now() - t0;
    }
    double t0 = pt_now();
    if (rc == MBPE_OK)
        rc = mbpe_pretok_dedup_segments(p, d_text.data(), seg_len.data(), d_off.data(), n_chunks.data(), n_segs, out, nullptr);
    const double t_dedup = pt_now() - t0;
    t0 = pt_now();
    cleanup();
    if (pt_debug())
        fprintf(stderr, "[mbpe] pretok_corpus: %.1f MB in %u segment(s): upload + split (overlapped) %.1f ms, dedup %.1f ms, free %.1f ms, total %.1f ms\n",
                len / 1e6, n_segs, (t_h2d + t_split) * 1e3, t_dedup * 1e3, (pt_now() - t0) * 1e3, (pt_now() - t_start) * 1e3);
    return rc;
}

// text in host memory -> ids in host memory (Tokenizer::encode without special tokens, Tokenizer.h:653-717): GPT-4 split
// + merge scan per resident segment, nothing but text going up and ids coming down. Segments (64 MiB, cut at matcher
// cuts) are pipelined over three streams -- upload of segment k+1 and download of segment k-1 run under the kernels
// of segment k -- so with pinned host buffers the call runs at the speed of the slower PCIe direction.
static int ensure_pipe(mbpe_pretok *p, uint64_t max_seg) {
    if (!p->st_c) {
        MB_CUDA(cudaStreamCreateWithFlags(&p->st_in, cudaStreamNonBlocking));
        MB_CUDA(cudaStreamCreateWithFlags(&p->st_c, cudaStreamNonBlocking));
        MB_CUDA(cudaStreamCreateWithFlags(&p->st_out, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            MB_CUDA(cudaEventCreateWithFlags(&p->ev_in[i], cudaEventDisableTiming));
            MB_CUDA(cudaEventCreateWithFlags(&p->ev_out[i], cudaEventDisableTiming));
        }
    }
    if (max_seg <= p->pipe_cap) return MBPE_OK;
    for (int i = 0; i < 2; i++) {
        cudaFree(p->d_pipe_text[i]);
        cudaFree(p->d_pipe_ids[i]);
        p->d_pipe_text[i] = nullptr, p->d_pipe_ids[i] = nullptr;
    }
    cudaFree(p->d_pipe_off);
    p->d_pipe_off = nullptr;
    p->pipe_cap = 0;
    const uint64_t cap = max_seg + max_seg / 16 + 64;
    for (int i = 0; i < 2; i++)
        if (cudaMalloc(&p->d_pipe_text[i], cap) != cudaSuccess || cudaMalloc(&p->d_pipe_ids[i], cap * 4) != cudaSuccess) {
            cudaGetLastError();
            return set_error(MBPE_E_CUDA, "out of device memory");
        }
    if (cudaMalloc(&p->d_pipe_off, (cap + 2) * 4) != cudaSuccess || (!p->d_seg_n && cudaMalloc(&p->d_seg_n, 8) != cudaSuccess)) {
        cudaGetLastError();
        return set_error(MBPE_E_CUDA, "out of device memory");
    }
    p->pipe_cap = cap;
    return MBPE_OK;
}

// Segment boundaries for a text with special-token occurrences: the last occurrence that begins in the usual range
// (its begin ends an ordinary part, so both sides are independent), else an ordinary cut -- which cannot lie inside an
// occurrence, because none begins in the range and occurrences are at most 31 bytes long.
static bool plan_segments_special(const mbpe_pretok *pt, const uint8_t *text, uint64_t len, uint64_t seg_bytes, const uint64_t *sp_b,
                                  uint32_t n_sp, std::vector<uint64_t> &bounds) {
    bounds.assign(1, 0);
    uint32_t err = 0;
    PretokIn<HostText> in{HostText{text}, len, pt->h_table.data(), &err, pt->kind, 0};
    while (len - bounds.back() > seg_bytes + seg_bytes / 4) {
        const uint64_t lo = bounds.back() + seg_bytes / 2, hi = bounds.back() + seg_bytes;
        uint64_t cut = 0;
        const uint64_t *it = std::upper_bound(sp_b, sp_b + n_sp, hi); // first occurrence that begins after hi
        if (it != sp_b && it[-1] > lo) {
            cut = it[-1];
        } else {
            bool bad = false;
            for (uint64_t q = hi; q > lo && !cut && !bad; q--) {
                if ((text[q] & 0xC0) == 0x80) continue;
                const PtCp prev = pt_before(in, q, bad), cur = pt_at(in, q, bad);
                if (!bad && pt_is_cut(prev.cls, cur)) cut = q;
            }
        }
        if (!cut) return false;
        bounds.push_back(cut);
    }
    bounds.push_back(len);
    return true;
}

extern "C" int mbpe_encode_text(mbpe_encoder *enc, mbpe_pretok *p, const uint8_t *text, uint64_t len, uint32_t *out,
                                uint64_t out_cap, uint64_t *n_out) {
    return mbpe_encode_text_special(enc, p, text, len, nullptr, nullptr, 0, out, out_cap, n_out);
}

// ... with special tokens: sp_begin / sp_end = the occurrences in text order (host arrays, Tokenizer.h:605-650); the
// encoder must have been told their ids (mbpe_encoder_seed_special_chunks)
extern "C" int mbpe_encode_text_special(mbpe_encoder *enc, mbpe_pretok *p, const uint8_t *text, uint64_t len,
                                        const uint64_t *sp_begin, const uint64_t *sp_end, uint64_t n_sp, uint32_t *out,
                                        uint64_t out_cap, uint64_t *n_out) {
    if (!enc || !p || !n_out || (len && (!text || !out)) || (n_sp && (!sp_begin || !sp_end)))
        return set_error(MBPE_E_INVALID, "null argument");
    if (n_sp >= (1ull << 31)) return set_error(MBPE_E_INVALID, "too many special-token occurrences");
    *n_out = 0;
    int rc = use_device(p->device);
    if (rc) return rc;
    std::vector<uint64_t> bounds;
    const bool planned = n_sp ? plan_segments_special(p, text, len, p->enc_seg_bytes, sp_begin, (uint32_t)n_sp, bounds)
                              : plan_segments(p, text, len, p->enc_seg_bytes, bounds);
    if (!planned)
        return set_error(MBPE_E_UNSUPPORTED, "no safe segment boundary found (malformed UTF-8 or no letters/blanks): use the PCRE2 path");
    const size_t n_seg = bounds.size() - 1;
    uint64_t max_seg = 0;
    for (size_t k = 0; k < n_seg; k++) max_seg = std::max(max_seg, bounds[k + 1] - bounds[k]);
    const double t_start = pt_now();
    if ((rc = ensure_pipe(p, max_seg))) return rc;
    auto upload = [&](size_t k) -> cudaError_t {
        cudaError_t ce = cudaMemcpyAsync(p->d_pipe_text[k & 1], text + bounds[k], bounds[k + 1] - bounds[k], cudaMemcpyHostToDevice, p->st_in);
        return ce != cudaSuccess ? ce : cudaEventRecord(p->ev_in[k & 1], p->st_in);
    };
    uint64_t produced = 0;
    std::vector<uint32_t> sp_rel_b, sp_rel_e;
    cudaError_t ce = n_seg ? upload(0) : cudaSuccess;
    for (size_t k = 0; k < n_seg && rc == MBPE_OK && ce == cudaSuccess; k++) {
        const uint64_t n = bounds[k + 1] - bounds[k];
        // buffer (k+1)&1 is free: the kernels of segment k-1 have finished (the host waited for their id count)
        if (k + 1 < n_seg && (ce = upload(k + 1)) != cudaSuccess) break;
        if ((ce = cudaStreamWaitEvent(p->st_c, p->ev_in[k & 1], 0)) != cudaSuccess) break;
        if (k >= 2 && (ce = cudaStreamWaitEvent(p->st_c, p->ev_out[k & 1], 0)) != cudaSuccess) break; // ids[k&1] drained
        uint64_t n_chunks = 0, n_ids = 0;
        if (n) {
            // the special-token occurrences that lie in this segment, relative to its start
            const uint64_t sb0 = bounds[k], sb1 = bounds[k + 1];
            const uint64_t *o0 = n_sp ? std::lower_bound(sp_begin, sp_begin + n_sp, sb0) : nullptr;
            const uint64_t *o1 = n_sp ? std::lower_bound(sp_begin, sp_begin + n_sp, sb1) : nullptr;
            const uint32_t n_here = n_sp ? (uint32_t)(o1 - o0) : 0;
            if (n_here) {
                sp_rel_b.resize(n_here);
                sp_rel_e.resize(n_here);
                for (uint32_t q = 0; q < n_here; q++) {
                    sp_rel_b[q] = (uint32_t)(o0[q] - sb0);
                    sp_rel_e[q] = (uint32_t)(sp_end[(o0 - sp_begin) + q] - sb0);
                }
                uint32_t *d_b = (uint32_t *)pt_scratch(p, 4, (uint64_t)n_here * 4), *d_e = (uint32_t *)pt_scratch(p, 5, (uint64_t)n_here * 4);
                if (!d_b || !d_e) {
                    rc = set_error(MBPE_E_CUDA, "out of device memory");
                    break;
                }
                if ((ce = cudaMemcpyAsync(d_b, sp_rel_b.data(), (uint64_t)n_here * 4, cudaMemcpyHostToDevice, p->st_c)) != cudaSuccess) break;
                if ((ce = cudaMemcpyAsync(d_e, sp_rel_e.data(), (uint64_t)n_here * 4, cudaMemcpyHostToDevice, p->st_c)) != cudaSuccess) break;
                if ((rc = mbpe_pretok_split_device_parts(p, p->d_pipe_text[k & 1], n, d_b, d_e, n_here, p->d_pipe_off, n + 2, &n_chunks, p->st_c))) break;
            } else if ((rc = mbpe_pretok_split_device(p, p->d_pipe_text[k & 1], n, p->d_pipe_off, n + 2, &n_chunks, p->st_c))) {
                break;
            }
            if ((rc = mbpe_encode_device(enc, p->d_pipe_text[k & 1], n, p->d_pipe_off, n_chunks, p->d_pipe_ids[k & 1], n, p->d_seg_n, p->st_c))) break;
            if ((ce = cudaMemcpyAsync(&n_ids, p->d_seg_n, 8, cudaMemcpyDeviceToHost, p->st_c)) != cudaSuccess) break;
        }
        if ((ce = cudaStreamSynchronize(p->st_c)) != cudaSuccess) break;
        if (produced + n_ids > out_cap) {
            rc = set_error(MBPE_E_CAPACITY, "output buffer too small");
            break;
        }
        if ((ce = cudaMemcpyAsync(out + produced, p->d_pipe_ids[k & 1], n_ids * 4, cudaMemcpyDeviceToHost, p->st_out)) != cudaSuccess) break;
        if ((ce = cudaEventRecord(p->ev_out[k & 1], p->st_out)) != cudaSuccess) break;
        produced += n_ids;
    }
    // nothing of this call may still be in flight when it returns, whatever happened
    cudaError_t e1 = cudaStreamSynchronize(p->st_in), e2 = cudaStreamSynchronize(p->st_c), e3 = cudaStreamSynchronize(p->st_out);
    if (rc == MBPE_OK) {
        if (ce == cudaSuccess) ce = e1 != cudaSuccess ? e1 : e2 != cudaSuccess ? e2 : e3;
        if (ce != cudaSuccess) rc = cuda_fail(ce, "encode_text pipeline", __FILE__, __LINE__);
    }
    if (pt_debug())
        fprintf(stderr, "[mbpe] encode_text: %.1f MB in %zu segment(s), %.1f ms\n", len / 1e6, n_seg, (pt_now() - t_start) * 1e3);
    if (rc == MBPE_OK) *n_out = produced;
    return rc;
}

// ---------------------------------------------------------------------------------------------------------
// File to file (the reference slurps the text through a stringstream and issues one write() per id,
// examples/minbpe-cc.cpp:36-46, :58-69). Here: a reader thread fills pinned blocks, the calling thread runs the
// upload / split + merge scan / download pipeline of mbpe_encode_text on them, a writer thread drains pinned id
// blocks into the .enc file. Memory use is a handful of blocks whatever the file size; blocks end at matcher cuts,
// the bytes after the last cut of a block are carried into the next one.
// ---------------------------------------------------------------------------------------------------------
namespace {
struct Gate { // hand-over of numbered blocks between two threads
    std::mutex mu;
    std::condition_variable cv;
    long long ready = -1; // highest block index published
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
    bool wait_for(long long k) { // false: the other side failed
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return ready >= k || failed; });
        return ready >= k;
    }
};
} // namespace

extern "C" int mbpe_encode_file(mbpe_encoder *enc, mbpe_pretok *p, const char *in_path, const char *out_path,
                                uint64_t *n_bytes_out, uint64_t *n_ids_out) {
    if (!enc || !p || !in_path || !out_path) return set_error(MBPE_E_INVALID, "null argument");
    int rc = use_device(p->device);
    if (rc) return rc;
    FILE *fin = fopen(in_path, "rb");
    if (!fin) return set_error(MBPE_E_IO, std::string("cannot open ") + in_path);
    FILE *fout = fopen(out_path, "wb");
    if (!fout) {
        fclose(fin);
        return set_error(MBPE_E_IO, std::string("cannot open ") + out_path);
    }
    const double t_begin = pt_now();
    // a block = carried tail (< SEG) + up to SEG fresh bytes; blocks are smaller than the in-memory path's segments so
    // that the pinned staging memory (3 x 10 x SEG) is allocated in ~0.1 s
    const uint64_t SEG = std::min<uint64_t>(p->enc_seg_bytes, 16u << 20), CAP = 2 * SEG + 64;
    constexpr int NB = 3;
    uint8_t *h_in[NB] = {nullptr, nullptr, nullptr};
    uint32_t *h_out[NB] = {nullptr, nullptr, nullptr};
    bool ok_alloc = ensure_pipe(p, CAP) == MBPE_OK;
    for (int i = 0; i < NB && ok_alloc; i++)
        ok_alloc = cudaMallocHost(&h_in[i], CAP) == cudaSuccess && cudaMallocHost(&h_out[i], CAP * 4) == cudaSuccess;
    auto release = [&]() {
        for (int i = 0; i < NB; i++) {
            cudaFreeHost(h_in[i]);
            cudaFreeHost(h_out[i]);
        }
        fclose(fin);
        fclose(fout);
    };
    if (!ok_alloc) {
        release();
        return set_error(MBPE_E_CUDA, "out of (pinned) memory");
    }
    // block k lives in h_in[k % NB]: [0, fill[k]) valid, the segment to encode is [0, cut[k]), the rest is carried over
    std::vector<uint64_t> fill, cut;
    std::vector<char> last;
    std::mutex meta_mu;
    Gate read_gate, in_free, write_gate, out_free;
    std::vector<uint64_t> ids_in_block;
    uint32_t err_flags = 0;
    PretokIn<HostText> hin{HostText{nullptr}, 0, p->h_table.data(), &err_flags, p->kind};
    bool unsupported = false;
    in_free.publish(NB - 1); // blocks 0..NB-1 may be filled right away
    out_free.publish(NB - 1);
    std::thread reader([&]() {
        uint64_t carry = 0;
        const uint8_t *carry_src = nullptr;
        for (long long k = 0;; k++) {
            if (!in_free.wait_for(k)) return;
            uint8_t *buf = h_in[k % NB];
            if (carry) memmove(buf, carry_src, carry); // from the previous block (a different buffer: NB >= 2)
            const size_t got = fread(buf + carry, 1, SEG, fin);
            const uint64_t filled = carry + got;
            const bool eof = got < SEG;
            uint64_t c = filled;
            if (!eof) { // last cut of the block, not too close to the end (a code point may be cut off there)
                c = 0;
                PretokIn<HostText> in = hin;
                in.t.p = buf;
                in.len = filled;
                bool bad = false;
                for (uint64_t q = filled - 8; q > filled / 4 && !c && !bad; q--) {
                    if ((buf[q] & 0xC0) == 0x80) continue;
                    const PtCp prev = pt_before(in, q, bad), cur = pt_at(in, q, bad);
                    if (!bad && pt_is_cut(prev.cls, cur)) c = q;
                }
                if (!c) {
                    unsupported = true;
                    read_gate.fail();
                    return;
                }
            }
            {
                std::lock_guard<std::mutex> lk(meta_mu);
                fill.push_back(filled);
                cut.push_back(c);
                last.push_back(eof);
            }
            read_gate.publish(k);
            if (eof) return;
            carry = filled - c;
            carry_src = buf + c;
            if (carry >= SEG) { // no cut in a whole segment: the carried tail would not fit the next block
                unsupported = true;
                read_gate.fail();
                return;
            }
        }
    });
    bool write_failed = false;
    long long n_blocks_total = -1; // set by the main thread when it has seen the last block
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
                n = ids_in_block[k];
            }
            if (n == ~0ull) return; // sentinel: no more blocks
            if (fwrite(h_out[k % NB], 4, n, fout) != n) {
                write_failed = true;
                out_free.fail();
                return;
            }
            out_free.publish(k + NB);
        }
    });
    uint64_t total_bytes = 0, total_ids = 0;
    cudaError_t ce = cudaSuccess;
    long long k = 0;
    const double t_loop = pt_now();
    double t_wait_read = 0, t_wait_write = 0;
    for (;; k++) {
        double tw = pt_now();
        const bool have_block = read_gate.wait_for(k);
        t_wait_read += pt_now() - tw;
        if (!have_block) {
            rc = unsupported ? set_error(MBPE_E_UNSUPPORTED, "no safe block boundary found in the file (malformed UTF-8 or no letters/blanks): use the whole-file path")
                             : set_error(MBPE_E_IO, "read failed");
            break;
        }
        uint64_t n;
        bool is_last;
        {
            std::lock_guard<std::mutex> lk(meta_mu);
            n = cut[k];
            is_last = last[k];
        }
        tw = pt_now();
        const bool have_room = out_free.wait_for(k);
        t_wait_write += pt_now() - tw;
        if (!have_room) {
            rc = set_error(MBPE_E_IO, std::string("write failed: ") + out_path);
            break;
        }
        uint64_t n_chunks = 0, n_ids = 0;
        if (n) {
            if ((ce = cudaMemcpyAsync(p->d_pipe_text[k & 1], h_in[k % NB], n, cudaMemcpyHostToDevice, p->st_c)) != cudaSuccess) break;
            if ((rc = mbpe_pretok_split_device(p, p->d_pipe_text[k & 1], n, p->d_pipe_off, n + 2, &n_chunks, p->st_c))) break;
            if (k >= 2 && (ce = cudaStreamWaitEvent(p->st_c, p->ev_out[k & 1], 0)) != cudaSuccess) break;
            if ((rc = mbpe_encode_device(enc, p->d_pipe_text[k & 1], n, p->d_pipe_off, n_chunks, p->d_pipe_ids[k & 1], n, p->d_seg_n, p->st_c))) break;
            if ((ce = cudaMemcpyAsync(&n_ids, p->d_seg_n, 8, cudaMemcpyDeviceToHost, p->st_c)) != cudaSuccess) break;
            if ((ce = cudaStreamSynchronize(p->st_c)) != cudaSuccess) break;
        }
        in_free.publish(k + NB); // the block's bytes are on the device (the reader copied its carried tail before this)
        // previous block's ids have landed? then hand them to the writer; this block's download runs under the next block
        if (k >= 1) {
            if ((ce = cudaEventSynchronize(p->ev_out[(k - 1) & 1])) != cudaSuccess) break;
            write_gate.publish(k - 1);
        }
        if ((ce = cudaMemcpyAsync(h_out[k % NB], p->d_pipe_ids[k & 1], n_ids * 4, cudaMemcpyDeviceToHost, p->st_out)) != cudaSuccess) break;
        if ((ce = cudaEventRecord(p->ev_out[k & 1], p->st_out)) != cudaSuccess) break;
        {
            std::lock_guard<std::mutex> lk(meta_mu);
            ids_in_block.push_back(n_ids);
        }
        total_bytes += n;
        total_ids += n_ids;
        if (is_last) {
            if ((ce = cudaEventSynchronize(p->ev_out[k & 1])) != cudaSuccess) break;
            {
                std::lock_guard<std::mutex> lk(total_mu);
                n_blocks_total = k + 1;
            }
            write_gate.publish(k);
            break;
        }
    }
    if (rc != MBPE_OK || ce != cudaSuccess) { // unblock the helpers
        in_free.fail();
        write_gate.fail();
    }
    const double t_joining = pt_now();
    reader.join();
    writer.join();
    if (pt_debug())
        fprintf(stderr, "[mbpe] encode_file: %.1f MB, %lld blocks: setup %.0f ms, pipeline %.0f ms (waiting for the reader %.0f ms, for the writer %.0f ms), "
                        "writer tail %.0f ms\n", total_bytes / 1e6, k + 1, (t_loop - t_begin) * 1e3, (t_joining - t_loop) * 1e3,
                t_wait_read * 1e3, t_wait_write * 1e3, (pt_now() - t_joining) * 1e3);
    cudaStreamSynchronize(p->st_c);
    cudaStreamSynchronize(p->st_out);
    if (rc == MBPE_OK && ce != cudaSuccess) rc = cuda_fail(ce, "encode_file pipeline", __FILE__, __LINE__);
    if (rc == MBPE_OK && (write_failed || fflush(fout) != 0)) rc = set_error(MBPE_E_IO, std::string("write failed: ") + out_path);
    release();
    if (rc == MBPE_OK) {
        if (n_bytes_out) *n_bytes_out = total_bytes;
        if (n_ids_out) *n_ids_out = total_ids;
    }
    return rc;
}
