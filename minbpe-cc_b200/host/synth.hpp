// synth.hpp -- the deterministic synthetic Zipfian UTF-8 corpus of the benchmarks (SURVEY 8(d) inputs 3-5). Header-only so
// that the library (mbpe_synth_corpus) and the stand-alone generator the reference arm of bench.py uses
// (oracle/synthgen.cpp -> oracle/_ref/synthgen, no product code in that process) produce the same bytes.
#pragma once
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace mbpe::host::synth {
// ---------------------------------------------------------------------------------------------------------
// synthetic Zipfian UTF-8 corpus (SURVEY 8(d) input 3). Deterministic in (seed, n): the text is generated in
// independent 1 MiB blocks, each seeded by (seed, block index), so any number of threads gives the same bytes.
// ---------------------------------------------------------------------------------------------------------
inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// independent stream per (seed, index): the state is a hash, not an offset into one shared sequence
inline uint64_t stream_seed(uint64_t seed, uint64_t index) { return mix64(mix64(seed) ^ mix64(index * 2 + 1)); }

struct SplitMix {
    uint64_t s;
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double unit() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
    uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32); }
};

constexpr uint32_t kWords = 1u << 20;
constexpr double kZipfS = 1.07;

struct WordList {
    std::vector<uint32_t> off;
    std::vector<uint8_t> bytes;
    std::vector<double> prob;   // alias table
    std::vector<uint32_t> alias;
};

// cumulative English-like letter frequencies (per 1000) for a..z
const uint16_t kLetterCum[26] = {82, 97, 125, 168, 295, 317, 337, 398, 468, 470, 478, 518, 542,
                                 609, 684, 703, 704, 764, 827, 918, 946, 956, 980, 982, 1000, 1001};
inline char letter(SplitMix &r) {
    uint32_t v = r.below(1001);
    for (int i = 0; i < 26; i++)
        if (v < kLetterCum[i]) return (char)('a' + i);
    return 'e';
}
inline void put_utf8(std::vector<uint8_t> &b, uint32_t cp) {
    if (cp < 0x80)
        b.push_back((uint8_t)cp);
    else if (cp < 0x800) {
        b.push_back(0xC0 | (cp >> 6));
        b.push_back(0x80 | (cp & 63));
    } else if (cp < 0x10000) {
        b.push_back(0xE0 | (cp >> 12));
        b.push_back(0x80 | ((cp >> 6) & 63));
        b.push_back(0x80 | (cp & 63));
    } else {
        b.push_back(0xF0 | (cp >> 18));
        b.push_back(0x80 | ((cp >> 12) & 63));
        b.push_back(0x80 | ((cp >> 6) & 63));
        b.push_back(0x80 | (cp & 63));
    }
}
inline uint32_t poisson4(SplitMix &r) { // Knuth, lambda = 4
    const double L = std::exp(-4.0);
    uint32_t k = 0;
    double p = 1.0;
    do {
        k++;
        p *= r.unit();
    } while (p > L);
    return k - 1;
}

// the word list ("language") is fixed; the corpus seed only drives which words are drawn, so a model trained
// on one synthetic corpus is meaningful on another
constexpr uint64_t kLanguageSeed = 0x5EED0000ull;

inline const WordList &word_list() {
    static WordList wl;
    static bool built = false;
    static std::mutex *mu = new std::mutex();
    std::lock_guard<std::mutex> lock(*mu);
    if (built) return wl;
    const uint64_t seed = kLanguageSeed;
    wl.off.reserve(kWords + 1);
    wl.off.push_back(0);
    for (uint32_t i = 0; i < kWords; i++) {
        SplitMix r{stream_seed(seed, i)};
        uint32_t len = std::min<uint32_t>(1 + poisson4(r), 16);
        double kind = r.unit();
        if (kind < 0.92) { // lowercase
            for (uint32_t k = 0; k < len; k++) wl.bytes.push_back((uint8_t)letter(r));
        } else if (kind < 0.95) { // Capitalised
            for (uint32_t k = 0; k < len; k++) {
                char c = letter(r);
                wl.bytes.push_back((uint8_t)(k == 0 ? c - 32 : c));
            }
        } else { // letters from other scripts: Latin-1 supplement, Cyrillic, CJK, emoji
            uint32_t script = r.below(4);
            for (uint32_t k = 0; k < len; k++) {
                uint32_t cp;
                switch (script) {
                case 0: cp = (k & 1) ? 0xE0 + r.below(23) : (uint32_t)letter(r); break; // àáâ... mixed with ASCII
                case 1: cp = 0x430 + r.below(32); break;                                 // а..я
                case 2: cp = 0x4E00 + r.below(2048); break;                              // CJK ideographs
                default: cp = 0x1F600 + r.below(64); break;                              // emoji
                }
                put_utf8(wl.bytes, cp);
                if (script >= 2 && k >= 3) break; // keep CJK / emoji words short
            }
        }
        wl.off.push_back((uint32_t)wl.bytes.size());
    }
    // Zipf(s) over ranks 1..kWords, Walker alias table
    std::vector<double> w(kWords);
    double sum = 0;
    for (uint32_t i = 0; i < kWords; i++) sum += (w[i] = std::pow((double)(i + 1), -kZipfS));
    wl.prob.assign(kWords, 0.0);
    wl.alias.assign(kWords, 0);
    std::vector<uint32_t> small, large;
    for (uint32_t i = 0; i < kWords; i++) {
        w[i] = w[i] / sum * kWords;
        (w[i] < 1.0 ? small : large).push_back(i);
    }
    while (!small.empty() && !large.empty()) {
        uint32_t s = small.back(), l = large.back();
        small.pop_back();
        wl.prob[s] = w[s];
        wl.alias[s] = l;
        w[l] = (w[l] + w[s]) - 1.0;
        if (w[l] < 1.0) {
            large.pop_back();
            small.push_back(l);
        }
    }
    for (uint32_t i : large) wl.prob[i] = 1.0;
    for (uint32_t i : small) wl.prob[i] = 1.0;
    built = true;
    return wl;
}

inline void fill_block(const WordList &wl, uint64_t seed, uint64_t block, uint8_t *out, uint64_t n) {
    SplitMix r{stream_seed(seed ^ 0xB10C5EEDull, block)};
    uint64_t o = 0;
    uint8_t tmp[80];
    while (o < n) {
        uint32_t wlen;
        const uint8_t *wp;
        if (r.unit() < 0.03) { // 1-6 digit number
            wlen = 1 + r.below(6);
            for (uint32_t k = 0; k < wlen; k++) tmp[k] = (uint8_t)('0' + r.below(10));
            wp = tmp;
        } else {
            uint32_t i = r.below(kWords);
            if (r.unit() >= wl.prob[i]) i = wl.alias[i];
            wp = wl.bytes.data() + wl.off[i];
            wlen = wl.off[i + 1] - wl.off[i];
        }
        double sep = r.unit();
        const char *sp = sep < 0.82 ? " " : sep < 0.87 ? ", " : sep < 0.92 ? ". " : sep < 0.98 ? "\n" : "\n\n";
        uint32_t slen = (uint32_t)strlen(sp);
        if (o + wlen + slen > n) { // tail of the block: pad with spaces, end the block on a newline
            while (o + 1 < n) out[o++] = ' ';
            out[o++] = '\n';
            break;
        }
        memcpy(out + o, wp, wlen);
        o += wlen;
        memcpy(out + o, sp, slen);
        o += slen;
    }
}

// blocks [first_block, first_block + ceil(n / 1 MiB)) of the corpus `seed`: bytes [first_block MiB, first_block MiB + n)
inline void generate(uint64_t seed, uint64_t first_block, uint8_t *out, uint64_t n, int n_threads) {
    if (n_threads <= 0) {
        n_threads = (int)std::thread::hardware_concurrency();
        if (n_threads <= 0) n_threads = 1;
    }
    const WordList &wl = word_list();
    const uint64_t B = 1ull << 20, n_blocks = (n + B - 1) / B;
    std::atomic<uint64_t> next{0};
    auto work = [&]() {
        for (uint64_t b; (b = next.fetch_add(1)) < n_blocks;) fill_block(wl, seed, first_block + b, out + b * B, std::min(B, n - b * B));
    };
    std::vector<std::thread> th;
    for (int t = 1; t < n_threads && (uint64_t)t < n_blocks; t++) th.emplace_back(work);
    work();
    for (auto &t : th) t.join();
}
} // namespace mbpe::host::synth
