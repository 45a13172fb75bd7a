// tests/emu/pretok_emu.cpp -- TEST INFRASTRUCTURE. Runs the matcher of csrc/pretok_core.cuh on the CPU: the same
// pretok_window() the CUDA kernel calls, one "thread" per window in a plain loop, marking chunk starts in a byte map.
// The CPU tests compare the result with PCRE2 (mbpe_split) on fixtures and fuzzed Unicode. Never linked into
// libminbpe_b200.so and not a fallback.
#include <stdint.h>
#include <string.h>

#include "../../minbpe-cc_b200/csrc/pretok_core.cuh"

using namespace mbpe;

struct HostText {
    const uint8_t *p;
    uint8_t operator[](uint64_t i) const { return p[i]; }
};

// marks[i] = 1 where a chunk starts; returns the error flags. order: 0 ascending windows, 1 descending.
static uint32_t g_kind = PT_GPT4;
extern "C" void emu_pretok_kind(int kind) { g_kind = (uint32_t)kind; }

extern "C" uint32_t emu_pretok(const uint8_t *text, uint64_t len, const uint8_t *table, uint64_t window,
                               uint64_t max_crawl, int order, uint8_t *marks) {
    uint32_t err = 0;
    memset(marks, 0, len);
    PretokIn<HostText> in{HostText{text}, len, table, &err, g_kind};
    const uint64_t n_win = (len + window - 1) / window;
    for (uint64_t k = 0; k < n_win; k++) {
        const uint64_t w = order ? n_win - 1 - k : k;
        pretok_window(in, w * window, (w + 1) * window, max_crawl, [&](uint64_t p) { marks[p] = 1; });
    }
    return err;
}

// the plain sequential loop (one window covering everything), for isolating matcher bugs from cut-rule bugs
extern "C" uint32_t emu_pretok_sequential(const uint8_t *text, uint64_t len, const uint8_t *table, uint8_t *marks) {
    return emu_pretok(text, len, table, len ? len : 1, ~0ull, 0, marks);
}

// the same with the text cut into independent subjects by special-token occurrences [sp_b[i], sp_e[i])
extern "C" uint32_t emu_pretok_parts(const uint8_t *text, uint64_t len, const uint8_t *table, const uint32_t *sp_b,
                                     const uint32_t *sp_e, uint32_t n_sp, uint64_t window, int order, uint8_t *marks) {
    uint32_t err = 0;
    memset(marks, 0, len);
    PretokIn<HostText> in{HostText{text}, len, table, &err, g_kind, 0};
    const uint64_t n_win = (len + window - 1) / window;
    for (uint64_t k = 0; k < n_win; k++) {
        const uint64_t w = order ? n_win - 1 - k : k;
        pretok_window_parts(in, sp_b, sp_e, n_sp, len, w * window, (w + 1) * window, ~0ull, [&](uint64_t p) { marks[p] = 1; });
    }
    return err;
}
