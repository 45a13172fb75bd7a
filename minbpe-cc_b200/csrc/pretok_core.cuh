// pretok_core.cuh -- the GPT-4 split pattern (Tokenizer.h:60) as a hand-written matcher, host + device (the simpler
// GPT-2 pattern of :59 rides along, see pretok_match_gpt2).
//
//   '(?i:[sdmt]|ll|ve|re) | [^\r\n\p{L}\p{N}]?+\p{L}+ | \p{N}{1,3} |  ?[^\s\p{L}\p{N}]++[\r\n]* | \s*[\r\n] | \s+(?!\S) | \s+
//
// The reference runs pcre2_match in a loop, each match starting where the last one ended (Tokenizer.h:506-540), with
// PCRE2_UTF | PCRE2_UCP | PCRE2_CASELESS (:407-415). Here:
//  * code point classes (\p{L}, \p{N}, \s) come from a 2-bit table that the host fills by ASKING the linked PCRE2
//    (host/chunker.cpp, pretok_class_table), so they are that library's Unicode tables by construction;
//  * pretok_match() is one step of the loop: the ordered alternation with its possessive / greedy / look-ahead
//    behaviour worked out by hand (comments at each alternative);
//  * pretok_window() makes the loop parallel. A position p is a CUT -- a place where a match is certain to start,
//    whatever came before -- when
//        (1) the code point before p is a letter and the one at p is not, or
//        (2) the code point at p is white space other than CR/LF and the one before p is not white space.
//    Proof sketch: a letter is only ever consumed by alternative 1 (fixed length, ends with a letter) or by the
//    greedy \p{L}+ of alternative 2, both of which end exactly where the letters end; a non-space code point is
//    consumed by alternatives 1-4, none of which can swallow a following blank (alternative 4 swallows CR/LF only).
//    Every thread owns a fixed window of text, looks for the first cut inside it and runs the sequential loop from
//    there until it reaches a cut at or beyond the end of its window. Threads may overlap; they mark chunk starts in
//    a bitmap, so overlap is idempotent. The only inputs that serialise are long stretches with neither letters nor
//    blanks (digit or punctuation runs, white-space runs), which the pattern itself makes sequential.
// The same code runs on the CPU in tests/emu/pretok_emu.cpp against PCRE2 on fixtures and fuzzed Unicode.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PT_HD __host__ __device__ __forceinline__
#else
#define PT_HD inline
#endif

namespace mbpe {

enum : uint32_t { PT_O = 0, PT_L = 1, PT_N = 2, PT_S = 3 }; // other, \p{L}, \p{N}, \s
constexpr uint32_t PT_TABLE_BYTES = 0x110000 / 4;           // 2 bits per code point
enum : uint32_t { PT_ERR_UTF8 = 1, PT_ERR_LONG = 2 };       // reasons to hand the text back to the host path
enum : uint32_t { PT_GPT4 = 0, PT_GPT2 = 1 };               // which built-in pattern (Tokenizer.h:60 / :59)

// Text accessor concept: uint8_t operator[](uint64_t) const.
template <class Text>
struct PretokIn {
    Text t;
    uint64_t len;         // end of subject (look-aheads see nothing beyond it)
    const uint8_t *table; // PT_TABLE_BYTES
    uint32_t *err;        // or-ed PT_ERR_*
    uint32_t kind;        // PT_GPT4 or PT_GPT2
    uint64_t begin;       // start of subject (0 unless the text is cut into independent subjects, pretok_window_parts)
};

PT_HD void pt_raise(uint32_t *err, uint32_t what) {
#if defined(__CUDA_ARCH__)
    atomicOr(err, what);
#else
    *err |= what;
#endif
}

PT_HD uint32_t pt_class(const uint8_t *table, uint32_t cp) {
    if (cp >= 0x110000) return PT_O;
    return (table[cp >> 2] >> ((cp & 3) * 2)) & 3u;
}

// decode the code point starting at p (p < len). Malformed input sets *bad and yields (byte, 1).
template <class Text>
PT_HD uint32_t pt_decode(const Text &t, uint64_t len, uint64_t p, uint32_t &n, bool &bad) {
    const uint32_t b0 = t[p];
    if (b0 < 0x80) {
        n = 1;
        return b0;
    }
    uint32_t need = b0 >= 0xF0 ? 3 : b0 >= 0xE0 ? 2 : 1;
    if (b0 < 0xC2 || b0 > 0xF4 || p + need >= len) { // lead byte out of range, or the sequence runs past the subject
        bad = true;
        n = 1;
        return b0;
    }
    uint32_t cp = b0 & (0x3Fu >> need);
    for (uint32_t i = 1; i <= need; i++) {
        const uint32_t b = t[p + i];
        if ((b & 0xC0) != 0x80) {
            bad = true;
            n = 1;
            return b0;
        }
        cp = (cp << 6) | (b & 0x3F);
    }
    // overlong forms, surrogates, > U+10FFFF
    if ((need == 2 && (cp < 0x800 || (cp >= 0xD800 && cp <= 0xDFFF))) || (need == 3 && (cp < 0x10000 || cp > 0x10FFFF))) {
        bad = true;
        n = 1;
        return b0;
    }
    n = need + 1;
    return cp;
}

struct PtCp {
    uint32_t cp, n, cls;
};
template <class Text>
PT_HD PtCp pt_at(const PretokIn<Text> &in, uint64_t p, bool &bad) {
    PtCp r;
    r.cp = pt_decode(in.t, in.len, p, r.n, bad);
    r.cls = pt_class(in.table, r.cp);
    return r;
}

// ASCII letter lower-cased, U+017F (long s) -> 's' (the only non-ASCII code point PCRE2's caseless matching folds
// onto one of s d m t l v e r; the host checks this against the linked library), anything else -> 0
PT_HD uint32_t pt_fold(uint32_t cp) {
    if (cp < 0x80) {
        const uint32_t l = cp | 0x20;
        return (l >= 'a' && l <= 'z') ? l : 0;
    }
    return cp == 0x17F ? (uint32_t)'s' : 0;
}
PT_HD bool pt_is_nl(uint32_t cp) { return cp == '\r' || cp == '\n'; }

// The GPT-2 pattern (Tokenizer.h:59), compiled by the reference without CASELESS:
//   '(?:[sdmt]|ll|ve|re) |  ?\p{L}+ |  ?\p{N}+ |  ?[^\s\p{L}\p{N}]+ | \s+(?!\S) | \s+
// The same two cut rules hold: letters are consumed only by alternatives 1-2, which end where the letters end, and
// alternatives 1-4 never swallow a following blank.
template <class Text>
PT_HD uint64_t pretok_match_gpt2(const PretokIn<Text> &in, uint64_t p, uint32_t *last_cls, bool &bad) {
    const uint64_t len = in.len;
    const PtCp c0 = pt_at(in, p, bad);
    if (c0.cp == '\'' && p + 1 < len) { // 1. case-sensitive, ASCII only
        const uint32_t d = in.t[p + 1];
        if (d == 's' || d == 'd' || d == 'm' || d == 't') {
            *last_cls = PT_L;
            return p + 2;
        }
        if (p + 2 < len) {
            const uint32_t e = in.t[p + 2];
            if ((d == 'l' && e == 'l') || (d == 'v' && e == 'e') || (d == 'r' && e == 'e')) {
                *last_cls = PT_L;
                return p + 3;
            }
        }
    }
    { // 2-4. optional blank, then a run of one class (a blank that is not followed by such a run is white space: 5-6)
        const uint64_t q = (c0.cp == ' ') ? p + 1 : p;
        if (q < len) {
            const PtCp cq = (q == p) ? c0 : pt_at(in, q, bad);
            if (cq.cls != PT_S) {
                uint64_t r = q + cq.n;
                while (r < len) {
                    const PtCp c = pt_at(in, r, bad);
                    if (c.cls != cq.cls) break;
                    r += c.n;
                }
                *last_cls = cq.cls;
                return r;
            }
        }
    }
    uint64_t r = p, last_start = p; // 5-6. white space
    uint32_t n_cp = 0;
    while (r < len) {
        const PtCp c = pt_at(in, r, bad);
        if (c.cls != PT_S) break;
        last_start = r;
        r += c.n;
        n_cp++;
    }
    *last_cls = PT_S;
    if (n_cp == 0) {
        bad = true;
        return p + c0.n;
    }
    if (r == len || n_cp < 2) return r; // end of subject, or a single white-space code point (\s+)
    return last_start;                  // \s+(?!\S): give one code point back
}

// One match starting at p (p < len). Returns its end; *last_cls = class of the last code point matched (PT_S if it
// ends with CR/LF), which is what the cut test needs.
template <class Text>
PT_HD uint64_t pretok_match(const PretokIn<Text> &in, uint64_t p, uint32_t *last_cls, bool &bad) {
    if (in.kind == PT_GPT2) return pretok_match_gpt2(in, p, last_cls, bad);
    const uint64_t len = in.len;
    const PtCp c0 = pt_at(in, p, bad);
    // 1. '(?i:[sdmt]|ll|ve|re)
    if (c0.cp == '\'' && p + 1 < len) {
        const PtCp c1 = pt_at(in, p + 1, bad);
        const uint32_t f1 = pt_fold(c1.cp);
        if (f1 == 's' || f1 == 'd' || f1 == 'm' || f1 == 't') {
            *last_cls = PT_L;
            return p + 1 + c1.n;
        }
        if ((f1 == 'l' || f1 == 'v' || f1 == 'r') && p + 1 + c1.n < len) {
            const PtCp c2 = pt_at(in, p + 1 + c1.n, bad);
            const uint32_t f2 = pt_fold(c2.cp);
            if ((f1 == 'l' && f2 == 'l') || (f1 != 'l' && f2 == 'e')) {
                *last_cls = PT_L;
                return p + 1 + c1.n + c2.n;
            }
        }
    }
    // 2. [^\r\n\p{L}\p{N}]?+\p{L}+ : the optional first code point is possessive, but giving it back could not help
    //    (it is not a letter), so "take it if it fits the class" is exact
    {
        uint64_t q = p;
        PtCp cq = c0;
        bool have = true;
        if (c0.cls != PT_L && c0.cls != PT_N && !pt_is_nl(c0.cp)) {
            q = p + c0.n;
            have = q < len;
            if (have) cq = pt_at(in, q, bad);
        }
        if (have && cq.cls == PT_L) {
            uint64_t r = q + cq.n;
            while (r < len) {
                const PtCp c = pt_at(in, r, bad);
                if (c.cls != PT_L) break;
                r += c.n;
            }
            *last_cls = PT_L;
            return r;
        }
    }
    // 3. \p{N}{1,3}
    if (c0.cls == PT_N) {
        uint64_t r = p + c0.n;
        for (int k = 1; k < 3 && r < len; k++) {
            const PtCp c = pt_at(in, r, bad);
            if (c.cls != PT_N) break;
            r += c.n;
        }
        *last_cls = PT_N;
        return r;
    }
    // 4.  ?[^\s\p{L}\p{N}]++[\r\n]* : if the blank is taken and no symbol follows, the retry without the blank fails
    //    too (a blank is \s), so one attempt decides
    {
        const uint64_t q = (c0.cp == ' ') ? p + 1 : p;
        if (q < len) {
            const PtCp cq = (q == p) ? c0 : pt_at(in, q, bad);
            if (cq.cls == PT_O) {
                uint64_t r = q + cq.n;
                while (r < len) {
                    const PtCp c = pt_at(in, r, bad);
                    if (c.cls != PT_O) break;
                    r += c.n;
                }
                *last_cls = PT_O;
                while (r < len && pt_is_nl(in.t[r])) {
                    r++;
                    *last_cls = PT_S;
                }
                return r;
            }
        }
    }
    // 5-7. white space. c0 is \s here (letters, numbers and symbols were taken above).
    uint64_t r = p, last_start = p, after_last_nl = 0;
    uint32_t n_cp = 0;
    while (r < len) {
        const PtCp c = pt_at(in, r, bad);
        if (c.cls != PT_S) break;
        last_start = r;
        r += c.n;
        n_cp++;
        if (pt_is_nl(c.cp)) after_last_nl = r;
    }
    *last_cls = PT_S;
    if (n_cp == 0) { // unreachable for well-formed text; keep the loop moving
        bad = true;
        return p + c0.n;
    }
    if (after_last_nl) return after_last_nl; // 5. \s*[\r\n]: greedy, then back to the last CR/LF of the run
    if (r == len) return r;                  // 6. \s+(?!\S) at the end of the subject
    if (n_cp >= 2) return last_start;        // 6. give one code point back so that white space follows
    return r;                                // 7. \s+ (a single blank before a non-blank)
}

// class of the code point that ENDS at p (p > 0); text is assumed well formed (bad set otherwise)
template <class Text>
PT_HD PtCp pt_before(const PretokIn<Text> &in, uint64_t p, bool &bad) {
    uint64_t q = p - 1;
    while (q > in.begin && p - q < 4 && (in.t[q] & 0xC0) == 0x80) q--;
    PtCp c = pt_at(in, q, bad);
    if (q + c.n != p) bad = true;
    return c;
}

PT_HD bool pt_is_cut(uint32_t prev_cls, const PtCp &cur) {
    if (prev_cls == PT_L) return cur.cls != PT_L;
    return cur.cls == PT_S && !pt_is_nl(cur.cp) && prev_cls != PT_S;
}

// The work of one thread: window [w0, w1) of the subject. mark(p) is called for every chunk start the thread
// establishes (possibly also established by other threads). max_crawl bounds the bytes a thread may walk past its
// window before the text is declared pathological (PT_ERR_LONG) and handed to the host path.
template <class Text, class Mark>
PT_HD void pretok_window(const PretokIn<Text> &in, uint64_t w0, uint64_t w1, uint64_t max_crawl, Mark mark) {
    const uint64_t len = in.len;
    if (w0 < in.begin) w0 = in.begin;
    if (w1 > len) w1 = len;
    if (w0 >= w1) return;
    bool bad = false;
    uint64_t p = w0;
    if (w0 > in.begin) {
        while (p < w1 && (in.t[p] & 0xC0) == 0x80) p++; // the code point straddling w0 belongs to the window before
        if (p >= w1) return;
        uint32_t prev = pt_before(in, p, bad).cls;
        for (;;) {
            const PtCp c = pt_at(in, p, bad);
            if (pt_is_cut(prev, c)) break;
            prev = c.cls;
            p += c.n;
            if (p >= w1) { // no cut starts inside this window
                if (bad) pt_raise(in.err, PT_ERR_UTF8);
                return;
            }
        }
    }
    for (;;) {
        mark(p);
        uint32_t last = PT_O;
        p = pretok_match(in, p, &last, bad);
        if (p >= len) break;
        if (p >= w1) {
            const PtCp c = pt_at(in, p, bad);
            if (pt_is_cut(last, c)) break;
            if (p - w1 > max_crawl) {
                pt_raise(in.err, PT_ERR_LONG);
                break;
            }
        }
    }
    if (bad) pt_raise(in.err, PT_ERR_UTF8);
}

// The text cut into independent subjects by special tokens (Tokenizer.h:605-650, :664-704: every ordinary part is
// split on its own, a special token is a chunk of its own). sp_b / sp_e: begin / end of the special occurrences in
// text order. The thread of window [w0, w1) walks the parts that intersect its window: inside an ordinary part it is
// pretok_window() with that part as the subject (part starts are match starts, look-aheads end at the part's end);
// it marks the start of every special occurrence that begins in its window.
template <class Text, class Mark>
PT_HD void pretok_window_parts(PretokIn<Text> in, const uint32_t *sp_b, const uint32_t *sp_e, uint32_t n_sp, uint64_t total_len,
                               uint64_t w0, uint64_t w1, uint64_t max_crawl, Mark mark) {
    if (w1 > total_len) w1 = total_len;
    uint32_t lo = 0, hi = n_sp; // first special occurrence that ends after w0
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (sp_e[mid] > w0)
            hi = mid;
        else
            lo = mid + 1;
    }
    uint32_t i = lo;
    uint64_t pos = w0;
    while (pos < w1) {
        const uint64_t pb = i ? sp_e[i - 1] : 0, pe = i < n_sp ? sp_b[i] : total_len; // the ordinary part before special i
        if (pos < pe) {
            in.begin = pb;
            in.len = pe;
            pretok_window(in, pos, w1, max_crawl, mark);
            if (pe >= w1) return;
            pos = pe;
        }
        if (i >= n_sp) return;
        if (sp_b[i] >= w0) mark(sp_b[i]);
        pos = sp_e[i];
        i++;
    }
}

} // namespace mbpe
