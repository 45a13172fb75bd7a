// chunker.cpp -- regex pre-tokenisation (PCRE2), chunk dedup and the synthetic corpus generator. Host side.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <queue>
#include <thread>

#include "bpe_host.hpp"
#include "synth.hpp"

namespace mbpe::host {

const char *const kGpt2Pattern = "'(?:[sdmt]|ll|ve|re)| ?\\p{L}+| ?\\p{N}+| ?[^\\s\\p{L}\\p{N}]+|\\s+(?!\\S)|\\s+";
const char *const kGpt4Pattern =
    "'(?i:[sdmt]|ll|ve|re)|[^\\r\\n\\p{L}\\p{N}]?+\\p{L}+|\\p{N}{1,3}| ?[^\\s\\p{L}\\p{N}]++[\\r\\n]*|\\s*[\\r\\n]|\\s+(?!\\S)|\\s+";

int hardware_threads() {
    unsigned n = std::thread::hardware_concurrency();
    return n ? (int)n : 1;
}

Regex::~Regex() {
    if (code_) pcre2_code_free_8(code_);
}

Regex &Regex::operator=(Regex &&o) noexcept {
    if (this != &o) {
        if (code_) pcre2_code_free_8(code_);
        code_ = o.code_;
        o.code_ = nullptr;
    }
    return *this;
}

int Regex::compile(const std::string &pattern, std::string *err) {
    if (code_) {
        pcre2_code_free_8(code_);
        code_ = nullptr;
    }
    if (pattern.empty()) return MBPE_OK; // encoder "basic": no regex (Tokenizer.h:400)
    uint32_t options = MBPE_PCRE2_UTF | MBPE_PCRE2_UCP;                            // Tokenizer.h:407
    if (pattern.find("(?i:") != std::string::npos) options |= MBPE_PCRE2_CASELESS; // Tokenizer.h:413-415
    int errcode = 0;
    size_t erroff = 0;
    code_ = pcre2_compile_8(reinterpret_cast<const uint8_t *>(pattern.data()), pattern.size(), options, &errcode,
                            &erroff, nullptr);
    if (!code_) {
        uint8_t buf[256];
        pcre2_get_error_message_8(errcode, buf, sizeof buf);
        if (err) *err = std::string("PCRE2 pattern compilation failed: ") + reinterpret_cast<char *>(buf);
        return MBPE_E_REGEX;
    }
    pcre2_jit_compile_8(code_, MBPE_PCRE2_JIT_COMPLETE); // failure just means the interpreter runs
    return MBPE_OK;
}

// ---------------------------------------------------------------------------------------------------------
// ASCII fast path for the two built-in patterns. The PCRE2 call costs ~1 us per 5-byte chunk; most text is
// ASCII, where each alternative of Tokenizer.h:59-60 is a few byte-class tests. The scanner below decides a
// match ONLY from bytes < 0x80 (including the byte it must look at past the end of the match); as soon as a
// byte >= 0x80 could influence the decision it reports "unknown" and that one match is taken from PCRE2.
// Equivalence with PCRE2 is tested on every fixture and on random ASCII/Unicode mixes (tests/test_host.py).
// ---------------------------------------------------------------------------------------------------------
namespace {
enum : uint8_t { C_OTHER = 0, C_LETTER = 1, C_DIGIT = 2, C_SPACE = 3, C_NEWLINE = 4, C_HIGH = 5 };
struct ByteClass {
    uint8_t c[256];
    constexpr ByteClass() : c() {
        for (int i = 0; i < 256; i++) {
            uint8_t k = C_OTHER;
            if (i >= 0x80) k = C_HIGH;
            else if ((i >= 'a' && i <= 'z') || (i >= 'A' && i <= 'Z')) k = C_LETTER;
            else if (i >= '0' && i <= '9') k = C_DIGIT;
            else if (i == '\r' || i == '\n') k = C_NEWLINE;
            else if (i == ' ' || i == '\t' || i == 0x0B || i == 0x0C) k = C_SPACE; // \s with UCP, ASCII part
            c[i] = k;
        }
    }
};
constexpr ByteClass kCls{};
inline bool is_ws(uint8_t k) { return k == C_SPACE || k == C_NEWLINE; }

// whitespace alternatives shared by both patterns. gpt4 adds `\s*[\r\n]` in front of `\s+(?!\S)|\s+`.
// Returns 1 = matched (end set), 0 = no match, -1 = unknown (needs PCRE2).
inline int scan_space(const uint8_t *t, uint64_t len, uint64_t pos, bool gpt4, uint64_t *end) {
    uint64_t r = pos, last_nl = 0;
    bool has_nl = false;
    while (r < len && is_ws(kCls.c[t[r]])) {
        if (kCls.c[t[r]] == C_NEWLINE) {
            has_nl = true;
            last_nl = r;
        }
        r++;
    }
    if (r == pos) return 0;
    if (r < len && kCls.c[t[r]] == C_HIGH) return -1; // a multi-byte Unicode space may continue the run
    if (gpt4 && has_nl) {                              // \s*[\r\n]: up to the last CR/LF of the run
        *end = last_nl + 1;
        return 1;
    }
    if (r == len || r - pos == 1) { // \s+(?!\S) at end of text; or a single space char: \s+(?!\S) fails, \s+ takes it
        *end = r;
        return 1;
    }
    *end = r - 1; // \s+(?!\S): leave the last space for the next match
    return 1;
}

inline int scan_gpt4(const uint8_t *t, uint64_t len, uint64_t pos, uint64_t *end) {
    const uint8_t c = t[pos], k = kCls.c[c];
    if (k == C_HIGH) return -1;
    if (c == '\'' && pos + 1 < len) { // '(?i:[sdmt]|ll|ve|re)
        const uint8_t d = t[pos + 1];
        if (d >= 0x80) return -1; // U+017F / U+212A fold to s / k under CASELESS
        const uint8_t dl = d | 0x20;
        if (dl == 's' || dl == 'd' || dl == 'm' || dl == 't') {
            *end = pos + 2;
            return 1;
        }
        if (pos + 2 < len) {
            const uint8_t e = t[pos + 2];
            if (e >= 0x80) {
                if (dl == 'l' || dl == 'v' || dl == 'r') return -1;
            } else {
                const uint8_t el = e | 0x20;
                if ((dl == 'l' && el == 'l') || (dl == 'v' && el == 'e') || (dl == 'r' && el == 'e')) {
                    *end = pos + 3;
                    return 1;
                }
            }
        }
    }
    // [^\r\n\p{L}\p{N}]?+\p{L}+
    if (k != C_NEWLINE && k != C_DIGIT) {
        uint64_t p = (k == C_LETTER) ? pos : pos + 1;
        if (p < len) {
            const uint8_t kp = kCls.c[t[p]];
            if (kp == C_HIGH) return -1;
            if (kp == C_LETTER) {
                uint64_t q = p + 1;
                while (q < len && kCls.c[t[q]] == C_LETTER) q++;
                if (q < len && kCls.c[t[q]] == C_HIGH) return -1; // a non-ASCII letter may continue the word
                *end = q;
                return 1;
            }
        }
    }
    if (k == C_DIGIT) { // \p{N}{1,3}
        uint64_t q = pos + 1;
        while (q < len && q < pos + 3 && kCls.c[t[q]] == C_DIGIT) q++;
        if (q < pos + 3 && q < len && kCls.c[t[q]] == C_HIGH) return -1; // a non-ASCII digit may continue
        *end = q;
        return 1;
    }
    { //  ?[^\s\p{L}\p{N}]++[\r\n]*
        uint64_t p = (c == ' ') ? pos + 1 : pos;
        if (p < len) {
            const uint8_t kp = kCls.c[t[p]];
            if (kp == C_HIGH) return -1;
            if (kp == C_OTHER) {
                uint64_t q = p + 1;
                while (q < len && kCls.c[t[q]] == C_OTHER) q++;
                if (q < len && kCls.c[t[q]] == C_HIGH) return -1; // a non-ASCII symbol may continue the run
                while (q < len && kCls.c[t[q]] == C_NEWLINE) q++;
                *end = q;
                return 1;
            }
        }
    }
    return scan_space(t, len, pos, true, end);
}

inline int scan_gpt2(const uint8_t *t, uint64_t len, uint64_t pos, uint64_t *end) {
    const uint8_t c = t[pos], k = kCls.c[c];
    if (k == C_HIGH) return -1;
    if (c == '\'' && pos + 1 < len) { // '(?:[sdmt]|ll|ve|re), case-sensitive
        const uint8_t d = t[pos + 1];
        if (d == 's' || d == 'd' || d == 'm' || d == 't') {
            *end = pos + 2;
            return 1;
        }
        if (pos + 2 < len) {
            const uint8_t e = t[pos + 2];
            if ((d == 'l' && e == 'l') || (d == 'v' && e == 'e') || (d == 'r' && e == 'e')) {
                *end = pos + 3;
                return 1;
            }
        }
    }
    //  ?\p{L}+ |  ?\p{N}+ |  ?[^\s\p{L}\p{N}]+   (same shape: optional space, then a run of one class)
    {
        uint64_t p = (c == ' ') ? pos + 1 : pos;
        if (p < len) {
            const uint8_t kp = kCls.c[t[p]];
            if (kp == C_HIGH) return -1;
            if (kp == C_LETTER || kp == C_DIGIT || kp == C_OTHER) {
                uint64_t q = p + 1;
                while (q < len && kCls.c[t[q]] == kp) q++;
                if (q < len && kCls.c[t[q]] == C_HIGH) return -1;
                *end = q;
                return 1;
            }
        }
    }
    return scan_space(t, len, pos, false, end);
}
} // namespace

// Matches over text[0, len) starting at `begin`, stopping at the first match that starts at or after `stop`.
// The subject is always the WHOLE text so that look-aheads at the end of a segment see the real next byte.
// fast: 0 = PCRE2 only, 1 = gpt2 scanner, 2 = gpt4 scanner (PCRE2 only for matches the scanner cannot decide).
template <class Emit>
static int match_loop(const Regex &re, const uint8_t *text, uint64_t len, uint64_t begin, uint64_t stop, int fast,
                      std::string *err, Emit emit) {
    pcre2_match_data_8 *md = pcre2_match_data_create_from_pattern_8(re.code(), nullptr);
    if (!md) {
        if (err) *err = "PCRE2 match data creation failed.";
        return MBPE_E_REGEX;
    }
    size_t offset = begin;
    int rc_out = MBPE_OK;
    while (offset < stop || (begin == stop && offset == begin)) {
        if (fast && offset < len) {
            uint64_t e = 0;
            int r = fast == 2 ? scan_gpt4(text, len, offset, &e) : scan_gpt2(text, len, offset, &e);
            if (r == 1) {
                emit((uint64_t)offset, e);
                offset = e;
                continue;
            }
        }
        int rc = pcre2_match_8(re.code(), text, len, offset, MBPE_PCRE2_NO_UTF_CHECK, md, nullptr);
        if (rc < 0) {
            if (rc != MBPE_PCRE2_ERROR_NOMATCH) {
                uint8_t buf[256];
                pcre2_get_error_message_8(rc, buf, sizeof buf);
                if (err) *err = std::string("PCRE2 match error: ") + reinterpret_cast<char *>(buf);
                rc_out = MBPE_E_REGEX;
            }
            break;
        }
        const size_t *ov = pcre2_get_ovector_pointer_8(md);
        size_t s = ov[0], e = ov[1];
        if (s == e) { // empty match: step one byte (Tokenizer.h:529-533)
            if (offset >= len) break;
            offset++;
            continue;
        }
        if (s >= stop) break; // belongs to the next segment
        emit((uint64_t)s, (uint64_t)e);
        offset = e;
    }
    pcre2_match_data_free_8(md);
    return rc_out;
}

int split_range(const Regex &re, const uint8_t *text, uint64_t len, uint64_t begin, uint64_t stop, int fast,
                std::vector<Span> &out, std::string *err) {
    if (re.empty()) { // whole text is one chunk (Tokenizer.h:541-544)
        out.push_back(Span{0, len});
        return MBPE_OK;
    }
    return match_loop(re, text, len, begin, stop, fast, err, [&](uint64_t s, uint64_t e) { out.push_back(Span{s, e}); });
}

// cut points p (0 < p < len) where no match of the GPT-2/GPT-4 patterns can span p: text[p-1] == '\n' and
// text[p] is a printable non-space ASCII byte (a byte >= 0x80 could start a multi-byte Unicode space).
static std::vector<uint64_t> safe_cuts(const uint8_t *text, uint64_t len, int parts) {
    std::vector<uint64_t> cuts{0};
    for (int k = 1; k < parts; k++) {
        uint64_t p = std::max<uint64_t>(len / parts * k, cuts.back() + 1);
        while (p < len && !(text[p - 1] == '\n' && text[p] >= 0x21 && text[p] <= 0x7E)) p++;
        if (p >= len) break;
        cuts.push_back(p);
    }
    cuts.push_back(len);
    return cuts;
}

static bool builtin_pattern(const std::string &p) { return p == kGpt2Pattern || p == kGpt4Pattern; }
static int fast_kind(const std::string &p) {
    const char *off = getenv("MBPE_NO_FAST_SPLIT"); // debugging aid: force every match through PCRE2
    if (off && *off == '1') return 0;
    return p == kGpt4Pattern ? 2 : p == kGpt2Pattern ? 1 : 0;
}

int split_parallel(const Regex &re, const std::string &pattern, const uint8_t *text, uint64_t len, int n_threads,
                   std::vector<Span> &out, std::string *err) {
    if (n_threads <= 0) n_threads = hardware_threads();
    if (re.empty() || !builtin_pattern(pattern) || n_threads == 1 || len < (1u << 16))
        return split_range(re, text, len, 0, len, fast_kind(pattern), out, err);
    std::vector<uint64_t> cuts = safe_cuts(text, len, n_threads * 4);
    const size_t n_seg = cuts.size() - 1;
    std::vector<std::vector<Span>> parts(n_seg);
    std::vector<int> rcs(n_seg, MBPE_OK);
    std::vector<std::string> errs(n_seg);
    std::atomic<size_t> next{0};
    const int fast = fast_kind(pattern);
    auto work = [&]() {
        for (size_t s; (s = next.fetch_add(1)) < n_seg;)
            rcs[s] = split_range(re, text, len, cuts[s], cuts[s + 1], fast, parts[s], &errs[s]);
    };
    std::vector<std::thread> th;
    for (int t = 1; t < n_threads; t++) th.emplace_back(work);
    work();
    for (auto &t : th) t.join();
    size_t total = out.size();
    for (size_t s = 0; s < n_seg; s++) {
        if (rcs[s] != MBPE_OK) {
            if (err) *err = errs[s];
            return rcs[s];
        }
        total += parts[s].size();
    }
    out.reserve(total);
    for (auto &p : parts) out.insert(out.end(), p.begin(), p.end());
    return MBPE_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Code point classes for the GPU matcher of the GPT-4 pattern (csrc/pretok_core.cuh), taken from the linked PCRE2:
// every code point is written into one buffer and \p{L}+, \p{N}+, \s+ are matched over it with the options the
// reference compiles with (Tokenizer.h:407), so the table IS that library's Unicode data. 2 bits per code point:
// 0 other, 1 letter, 2 number, 3 white space.
// ---------------------------------------------------------------------------------------------------------
namespace {
inline uint64_t cp_offset(uint32_t cp) { // byte offset of code point cp in the all-code-points buffer
    if (cp < 0x80) return cp;
    if (cp < 0x800) return 0x80 + 2ull * (cp - 0x80);
    const uint64_t base3 = 0x80 + 2ull * (0x800 - 0x80);
    if (cp < 0xD800) return base3 + 3ull * (cp - 0x800);
    if (cp < 0x10000) return base3 + 3ull * (cp - 0x800 - 0x800); // surrogates D800..DFFF are left out
    return base3 + 3ull * (0x10000 - 0x800 - 0x800) + 4ull * (cp - 0x10000);
}
inline uint32_t cp_at_offset(uint64_t off) {
    if (off < 0x80) return (uint32_t)off;
    const uint64_t base3 = 0x80 + 2ull * (0x800 - 0x80);
    if (off < base3) return (uint32_t)(0x80 + (off - 0x80) / 2);
    const uint64_t base4 = base3 + 3ull * (0x10000 - 0x800 - 0x800);
    if (off < base4) {
        uint32_t cp = (uint32_t)(0x800 + (off - base3) / 3);
        return cp >= 0xD800 ? cp + 0x800 : cp;
    }
    return (uint32_t)(0x10000 + (off - base4) / 4);
}
} // namespace

int pretok_class_table(uint8_t *table, std::string *err) {
    std::vector<uint8_t> all(cp_offset(0x110000));
    for (uint32_t cp = 0; cp < 0x110000; cp++) {
        if (cp >= 0xD800 && cp <= 0xDFFF) continue;
        uint8_t *o = &all[cp_offset(cp)];
        if (cp < 0x80) {
            o[0] = (uint8_t)cp;
        } else if (cp < 0x800) {
            o[0] = (uint8_t)(0xC0 | (cp >> 6));
            o[1] = (uint8_t)(0x80 | (cp & 0x3F));
        } else if (cp < 0x10000) {
            o[0] = (uint8_t)(0xE0 | (cp >> 12));
            o[1] = (uint8_t)(0x80 | ((cp >> 6) & 0x3F));
            o[2] = (uint8_t)(0x80 | (cp & 0x3F));
        } else {
            o[0] = (uint8_t)(0xF0 | (cp >> 18));
            o[1] = (uint8_t)(0x80 | ((cp >> 12) & 0x3F));
            o[2] = (uint8_t)(0x80 | ((cp >> 6) & 0x3F));
            o[3] = (uint8_t)(0x80 | (cp & 0x3F));
        }
    }
    memset(table, 0, 0x110000 / 4);
    // runs of one class; `cls` 0 collects the caseless partners of the letters of alternative 1 instead
    auto runs = [&](const char *pattern, uint32_t cls, std::vector<uint32_t> *hits) -> int {
        Regex re;
        int rc = re.compile(pattern, err);
        if (rc) return rc;
        pcre2_match_data_8 *md = pcre2_match_data_create_from_pattern_8(re.code(), nullptr);
        if (!md) return MBPE_E_REGEX;
        size_t off = 0;
        while (off < all.size()) {
            int m = pcre2_match_8(re.code(), all.data(), all.size(), off, MBPE_PCRE2_NO_UTF_CHECK, md, nullptr);
            if (m < 0) break;
            const size_t *ov = pcre2_get_ovector_pointer_8(md);
            if (ov[1] <= ov[0]) break;
            const uint32_t a = cp_at_offset(ov[0]), b = ov[1] >= all.size() ? 0x110000u : cp_at_offset(ov[1]);
            for (uint32_t cp = a; cp < b; cp++) {
                if (cp >= 0xD800 && cp <= 0xDFFF) continue;
                if (hits)
                    hits->push_back(cp);
                else
                    table[cp >> 2] |= (uint8_t)(cls << ((cp & 3) * 2));
            }
            off = ov[1];
        }
        pcre2_match_data_free_8(md);
        return MBPE_OK;
    };
    int rc;
    if ((rc = runs("\\p{L}+", 1, nullptr))) return rc;
    if ((rc = runs("\\p{N}+", 2, nullptr))) return rc;
    if ((rc = runs("\\s+", 3, nullptr))) return rc;
    // the matcher hard-codes which code points fold onto s d m t l v e r: ASCII both cases, and U+017F -> s
    std::vector<uint32_t> hits;
    if ((rc = runs("(?i:[sdmtlver])+", 0, &hits))) return rc;
    std::vector<uint32_t> want;
    for (const char *q = "DELMRSTVdelmrstv"; *q; q++) want.push_back((uint32_t)*q);
    want.push_back(0x17F);
    Regex long_s;
    if ((rc = long_s.compile("^(?i:s)$", err))) return rc;
    pcre2_match_data_8 *md = pcre2_match_data_create_from_pattern_8(long_s.code(), nullptr);
    const uint8_t ls[2] = {0xC5, 0xBF};
    const bool folds = md && pcre2_match_8(long_s.code(), ls, 2, 0, MBPE_PCRE2_NO_UTF_CHECK, md, nullptr) >= 0;
    if (md) pcre2_match_data_free_8(md);
    if (hits != want || !folds) {
        if (err) *err = "the linked PCRE2 folds case differently from the GPU matcher's assumption (U+017F only)";
        return MBPE_E_REGEX;
    }
    return MBPE_OK;
}

bool marker_token(std::string_view chunk, Token *id) {
    if (chunk.empty() || chunk[0] != '\0') return false;
    try {
        int v = std::stoi(std::string(chunk.substr(1))); // Tokenizer.h:88
        *id = static_cast<Token>(v);
        return true;
    } catch (...) {
        return false; // falls back to plain bytes (Tokenizer.h:90-92)
    }
}

// ---------------------------------------------------------------------------------------------------------
// dedup: unique chunk -> count, first-appearance order
// ---------------------------------------------------------------------------------------------------------
static inline uint64_t hash_bytes(const uint8_t *p, uint64_t n) {
    uint64_t h = 0x9E3779B97F4A7C15ull ^ (n * 0xff51afd7ed558ccdULL);
    while (n >= 8) {
        uint64_t w;
        memcpy(&w, p, 8);
        h = (h ^ w) * 0xc4ceb9fe1a85ec53ULL;
        h ^= h >> 29;
        p += 8;
        n -= 8;
    }
    uint64_t w = 0;
    memcpy(&w, p, n);
    h = (h ^ w) * 0xff51afd7ed558ccdULL;
    h ^= h >> 32;
    return h;
}

// Open-addressed set of chunks. A slot keeps the chunk's bytes inline when it is <= 15 bytes (almost every
// chunk of a regex split), so a repeat costs ONE cache line: no trip to the item list or back into the text.
struct UniqueSet {
    struct Item {
        uint64_t hash, start;
        uint32_t len, count;
    };
    struct Slot {
        uint64_t k0, k1; // len <= 15: the bytes (little endian, zero padded) and len in the top byte of k1
                         // longer: k0 = hash, k1 = len | LONG_TAG; empty: k1 == 0
        uint32_t idx;    // index into items
        uint32_t count;  // occurrences seen through this slot (folded into items at the end)
    };
    static constexpr uint64_t LONG_TAG = 0xFFull << 56;
    std::vector<Item> items; // first-appearance order
    std::vector<Slot> slots;
    uint64_t mask = 0;

    void init(uint64_t cap) {
        uint64_t c = 1024;
        while (c < cap) c <<= 1;
        slots.assign(c, Slot{0, 0, 0, 0});
        mask = c - 1;
    }
    static inline void make_key(const uint8_t *p, uint32_t len, uint64_t hash, uint64_t *k0, uint64_t *k1) {
        if (len <= 15) {
            uint64_t a = 0, b = 0;
            if (len >= 8) {
                memcpy(&a, p, 8);
                memcpy(&b, p + 8, len - 8);
            } else {
                memcpy(&a, p, len);
            }
            *k0 = a;
            *k1 = b | ((uint64_t)len << 56);
        } else {
            *k0 = hash;
            *k1 = (uint64_t)len | LONG_TAG;
        }
    }
    void grow() {
        std::vector<Slot> ns((mask + 1) * 2, Slot{0, 0, 0, 0});
        uint64_t m = ns.size() - 1;
        for (const Slot &sl : slots) {
            if (!sl.k1) continue;
            uint64_t h = items[sl.idx].hash & m;
            while (ns[h].k1) h = (h + 1) & m;
            ns[h] = sl;
        }
        slots.swap(ns);
        mask = m;
    }
    void add(const uint8_t *text, uint64_t hash, uint64_t start, uint32_t len, uint32_t count) {
        uint64_t k0, k1;
        make_key(text + start, len, hash, &k0, &k1);
        uint64_t h = hash & mask;
        for (;;) {
            Slot &sl = slots[h];
            if (sl.k1 == 0) break;
            if (sl.k0 == k0 && sl.k1 == k1 &&
                (len <= 15 || memcmp(text + items[sl.idx].start, text + start, len) == 0)) {
                sl.count += count;
                return;
            }
            h = (h + 1) & mask;
        }
        slots[h] = Slot{k0, k1, (uint32_t)items.size(), count};
        items.push_back(Item{hash, start, len, 0});
        if (items.size() * 2 > mask) grow();
    }
    void finish() { // fold the per-slot counters into the item list
        for (const Slot &sl : slots)
            if (sl.k1) items[sl.idx].count += sl.count;
    }
};

static void corpus_from_set(const uint8_t *text, const UniqueSet &g, uint64_t n_chunks, Corpus &out) {
    out.n_chunks = n_chunks;
    out.off.clear();
    out.weight.clear();
    out.tokens.clear();
    uint64_t total = 0;
    for (const auto &it : g.items) total += it.len;
    out.tokens.reserve(total);
    out.off.reserve(g.items.size() + 1);
    out.weight.reserve(g.items.size());
    out.off.push_back(0);
    for (const auto &it : g.items) {
        Token id;
        std::string_view sv(reinterpret_cast<const char *>(text + it.start), it.len);
        if (it.len && text[it.start] == 0 && marker_token(sv, &id))
            out.tokens.push_back(id);
        else
            for (uint32_t i = 0; i < it.len; i++) out.tokens.push_back(text[it.start + i]);
        out.off.push_back(out.tokens.size());
        out.weight.push_back(it.count);
    }
}

// Merge per-thread sets into local[0], keeping global first-appearance order (thread order = text order).
// Parallel by hash partition: worker p walks every local item list in order and owns the keys whose hash falls
// in partition p, so each partition list comes out sorted by (thread, index); a P-way merge of the partition
// lists on that key restores the global order.
static void merge_sets(const uint8_t *text, std::vector<UniqueSet> &local) {
    const size_t T = local.size();
    if (T == 1) {
        local[0].finish();
        return;
    }
    const int P = (int)std::min<size_t>(T, (size_t)hardware_threads());
    std::vector<UniqueSet> part(P);
    std::vector<std::vector<uint64_t>> order(P); // (thread << 32 | index) of each partition item's first appearance
    std::vector<std::thread> th;
    {
        std::atomic<size_t> next{0};
        auto fin = [&]() {
            for (size_t t; (t = next.fetch_add(1)) < T;) local[t].finish();
        };
        for (int p = 1; p < P; p++) th.emplace_back(fin);
        fin();
        for (auto &t : th) t.join();
        th.clear();
    }
    size_t total = 0;
    for (auto &l : local) total += l.items.size();
    auto work = [&](int p) {
        UniqueSet &g = part[p];
        g.init(total / P / 2 + 1024);
        for (size_t t = 0; t < T; t++) {
            const auto &items = local[t].items;
            for (size_t i = 0; i < items.size(); i++) {
                const auto &it = items[i];
                if ((int)(((it.hash >> 40) * (uint64_t)P) >> 24) != p) continue; // top 24 hash bits pick the partition
                size_t before = g.items.size();
                g.add(text, it.hash, it.start, it.len, it.count);
                if (g.items.size() != before) order[p].push_back(((uint64_t)t << 32) | (uint64_t)i);
            }
        }
        g.finish();
    };
    for (int p = 1; p < P; p++) th.emplace_back(work, p);
    work(0);
    for (auto &t : th) t.join();
    // P-way merge on the first-appearance key
    UniqueSet out;
    out.items.reserve(total);
    std::vector<size_t> cur(P, 0);
    using Head = std::pair<uint64_t, int>;
    std::priority_queue<Head, std::vector<Head>, std::greater<Head>> heap;
    for (int p = 0; p < P; p++)
        if (!order[p].empty()) heap.push({order[p][0], p});
    while (!heap.empty()) {
        auto [key, p] = heap.top();
        heap.pop();
        out.items.push_back(part[p].items[cur[p]]);
        if (++cur[p] < order[p].size()) heap.push({order[p][cur[p]], p});
    }
    local[0].items.swap(out.items);
    local[0].slots.clear(); // counts are final in items; the set is not added to again
    local[0].mask = 0;
}

void dedup_chunks(const uint8_t *text, const std::vector<Span> &chunks, int n_threads, Corpus &out) {
    if (n_threads <= 0) n_threads = hardware_threads();
    const uint64_t n = chunks.size();
    if (n < (1u << 16)) n_threads = 1;
    std::vector<UniqueSet> local(n_threads);
    auto work = [&](int t) {
        uint64_t k0 = n * t / n_threads, k1 = n * (t + 1) / n_threads;
        local[t].init(1 << 16);
        for (uint64_t k = k0; k < k1; k++) {
            uint64_t s = chunks[k].start, l = chunks[k].end - s;
            local[t].add(text, hash_bytes(text + s, l), s, (uint32_t)l, 1);
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < n_threads; t++) th.emplace_back(work, t);
    work(0);
    for (auto &t : th) t.join();
    merge_sets(text, local);
    corpus_from_set(text, local[0], n, out);
}

// split + dedup fused: the chunk list of a 1 GiB text (2 x 10^8 spans) is never materialised. Each thread scans
// one contiguous range of the text (between regex-safe cuts) straight into its own set.
int split_dedup_parallel(const Regex &re, const std::string &pattern, const uint8_t *text, uint64_t len, int n_threads,
                         Corpus &out, std::string *err) {
    if (n_threads <= 0) n_threads = hardware_threads();
    if (re.empty()) { // whole text is one chunk (Tokenizer.h:541-544)
        std::vector<Span> one{Span{0, len}};
        dedup_chunks(text, one, 1, out);
        return MBPE_OK;
    }
    if (!builtin_pattern(pattern) || len < (1u << 16)) n_threads = 1;
    std::vector<uint64_t> cuts = n_threads > 1 ? safe_cuts(text, len, n_threads) : std::vector<uint64_t>{0, len};
    const int n_seg = (int)cuts.size() - 1;
    std::vector<UniqueSet> local(n_seg);
    std::vector<int> rcs(n_seg, MBPE_OK);
    std::vector<std::string> errs(n_seg);
    std::vector<uint64_t> counts(n_seg, 0);
    const int fast = fast_kind(pattern);
    auto work = [&](int s) {
        UniqueSet &set = local[s];
        set.init(1 << 16);
        uint64_t n = 0;
        rcs[s] = match_loop(re, text, len, cuts[s], cuts[s + 1], fast, &errs[s], [&](uint64_t a, uint64_t b) {
            set.add(text, hash_bytes(text + a, b - a), a, (uint32_t)(b - a), 1);
            n++;
        });
        counts[s] = n;
    };
    const bool dbg = getenv("MBPE_DEBUG") != nullptr;
    auto tnow = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = tnow();
    std::vector<std::thread> th;
    for (int s = 1; s < n_seg; s++) th.emplace_back(work, s);
    work(0);
    for (auto &t : th) t.join();
    const double t1 = tnow();
    uint64_t n_chunks = 0;
    for (int s = 0; s < n_seg; s++) {
        if (rcs[s] != MBPE_OK) {
            if (err) *err = errs[s];
            return rcs[s];
        }
        n_chunks += counts[s];
    }
    merge_sets(text, local);
    const double t2 = tnow();
    corpus_from_set(text, local[0], n_chunks, out);
    if (dbg)
        fprintf(stderr, "[mbpe] split+dedup: %d threads, scan %.3f s, merge %.3f s, build %.3f s\n", n_seg, t1 - t0, t2 - t1,
                tnow() - t2);
    return MBPE_OK;
}

// synthetic Zipfian UTF-8 corpus (SURVEY 8(d) input 3): host/synth.hpp
void synth_corpus(uint64_t seed, uint8_t *out, uint64_t n, int n_threads) { synth::generate(seed, 0, out, n, n_threads); }
void synth_corpus_at(uint64_t seed, uint64_t first_block, uint8_t *out, uint64_t n, int n_threads) {
    synth::generate(seed, first_block, out, n, n_threads);
}

} // namespace mbpe::host
