// lookback.cuh -- single-pass ordered placement of variable-sized tile outputs (decoupled look-back), shared by the
// encode kernel (ids of a tile), the decode kernel (bytes of a tile) and the pre-tokeniser (chunk starts of a tile).
#pragma once
#include <stdint.h>

namespace mbpe {

// ---------------------------------------------------------------------------------------------------------
// decoupled look-back over tiles (status word = flag:2 | value:62)
// ---------------------------------------------------------------------------------------------------------
constexpr uint64_t LB_AGG = 1ull << 62, LB_PREFIX = 2ull << 62, LB_VAL = (1ull << 62) - 1;

// status words are polled with GPU-scope relaxed loads (a `volatile` access compiles to a SYSTEM-scope one)
__device__ __forceinline__ unsigned long long lb_load(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Called by all 32 lanes of one warp. Each round inspects 32 * W predecessors at once (W independent loads per lane,
// one L2 round trip), so a tile that starts while several hundred older tiles are still in flight resolves its base
// in a handful of rounds instead of hundreds of serial loads. W = 1 is the round-1 behaviour; the tile kernels of
// encode and decode use W = 4 (128 predecessors per round trip: with ~900 tiles in flight the walk is <= 8 rounds).
// The two halves of a look-back, for callers that have work to do in between (the longer the gap, the more of the
// predecessors have announced themselves by the time the walk starts). announce: this tile's total becomes visible.
// Returns the base of tile 0 (ids of earlier launches; read ONCE, here: the stream's last tile may overwrite the word as
// soon as tile 0 is announced), 0 for the other tiles.
__device__ __forceinline__ uint64_t lookback_announce(unsigned long long *status, uint32_t tile, uint64_t total,
                                                      const unsigned long long *first_base = nullptr) {
    if (tile == 0) {
        const uint64_t b = first_base ? *((const volatile unsigned long long *)first_base) : 0;
        __syncwarp();
        if ((threadIdx.x & 31) == 0) atomicExch(&status[0], LB_PREFIX | (b + total));
        return b;
    }
    if ((threadIdx.x & 31) == 0) atomicExch(&status[tile], LB_AGG | total);
    return 0;
}
// resolve: the walk over the predecessors; returns the tile's base and publishes its inclusive prefix.
// base0 = what lookback_announce returned.
template <int W = 1>
__device__ __forceinline__ uint64_t lookback_resolve(unsigned long long *status, uint32_t tile, uint64_t total, uint64_t base0) {
    const uint32_t lane = threadIdx.x & 31;
    if (tile == 0) return base0;
    uint64_t acc = 0;
    int64_t j = (int64_t)tile - 1; // lane l, word w looks at tile j - (w * 32 + l): nearest predecessors first
    for (;;) {
        unsigned long long v[W];
#pragma unroll
        for (int w = 0; w < W; w++) {
            const int64_t idx = j - (w * 32 + (int)lane);
            // tiles before 0 do not exist: tile 0 always ends the walk itself
            v[w] = idx >= 0 ? lb_load(&status[idx]) : LB_PREFIX;
        }
        bool done = false;
#pragma unroll
        for (int w = 0; w < W; w++) {
            const int64_t idx = j - (w * 32 + (int)lane);
            if (idx >= 0)
                while ((v[w] >> 62) == 0) v[w] = lb_load(&status[idx]); // not published yet
            const unsigned pm = __ballot_sync(0xffffffffu, (v[w] & LB_PREFIX) != 0);
            uint64_t val = v[w] & LB_VAL;
            if (pm) {
                const int first = __ffs(pm) - 1; // nearest predecessor that already knows its inclusive prefix
                if ((int)lane > first || idx < 0) val = 0;
                done = true;
            }
            for (int d = 16; d > 0; d >>= 1) val += __shfl_xor_sync(0xffffffffu, val, d);
            acc += val;
            if (done) break;
        }
        if (done) break;
        j -= 32 * W;
    }
    if (lane == 0) atomicExch(&status[tile], LB_PREFIX | (acc + total));
    return acc;
}

template <int W = 1>
__device__ __forceinline__ uint64_t lookback_base(unsigned long long *status, uint32_t tile, uint64_t total,
                                                  const unsigned long long *first_base = nullptr) {
    const uint64_t b0 = lookback_announce(status, tile, total, first_base);
    return lookback_resolve<W>(status, tile, total, b0);
}

} // namespace mbpe
