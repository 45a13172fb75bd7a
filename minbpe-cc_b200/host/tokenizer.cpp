// tokenizer.cpp -- host mirror of MinBpeCC::Tokenizer::Tokenizer (Tokenizer.h:379-927) over the GPU engine,
// and the tokenizer / host-only part of the C ABI.
#include <chrono>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <thread>
#include <algorithm>

#include "bpe_host.hpp"

namespace mbpe::host {

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

Tokenizer::Tokenizer(const std::string &pattern, int device) : pattern_(pattern), device_(device) {
    if (regex_.compile(pattern_, &error_) != MBPE_OK) throw std::runtime_error(error_); // Tokenizer.h:427-431
    rebuild_vocab();
}

Tokenizer::~Tokenizer() {
    if (encoder_) mbpe_encoder_destroy(encoder_);
    if (pretok_) mbpe_pretok_destroy(pretok_);
}

// The GPT-4 and GPT-2 patterns have a device matcher (csrc/pretok_core.cuh). Texts of at least this many bytes go through it;
// MBPE_GPU_SPLIT=0 keeps pre-tokenisation on the host (PCRE2), MBPE_GPU_SPLIT=<n> sets the threshold (1 = always).
// Under these patterns no chunk can be a ready-made id (a chunk "\0<int>"): a chunk that starts with NUL
// continues with letters (alternative 2) or symbols (alternative 4), never with a digit, so std::stoi always throws.
static uint64_t gpu_split_min_bytes() {
    const char *v = getenv("MBPE_GPU_SPLIT");
    if (!v || !*v) return 1u << 16;
    const uint64_t n = strtoull(v, nullptr, 10);
    return n == 0 ? ~0ull : n;
}

bool Tokenizer::use_gpu_split(size_t n_bytes) {
    if ((pattern_ != kGpt4Pattern && pattern_ != kGpt2Pattern) || n_bytes < gpu_split_min_bytes()) return false;
    if (!pretok_ && !pretok_failed_) {
        if (mbpe_pretok_create(device_, &pretok_) != MBPE_OK) {
            if (pretok_) mbpe_pretok_destroy(pretok_);
            pretok_ = nullptr;
            pretok_failed_ = true;
        }
        pretok_pattern_.clear();
    }
    // load() may have replaced the pattern since the matcher was set up (Tokenizer.h:782-814 recompiles the regex):
    // the device matcher must follow, or large texts would split under the old pattern and small ones under the new
    if (pretok_ && pretok_pattern_ != pattern_) {
        if (mbpe_pretok_select(pretok_, pattern_.c_str()) != MBPE_OK) {
            mbpe_pretok_destroy(pretok_);
            pretok_ = nullptr;
            pretok_failed_ = true;
        } else {
            pretok_pattern_ = pattern_;
        }
    }
    return pretok_ != nullptr;
}

void Tokenizer::set_special_tokens_from_file(const std::string &contents) {
    special_tokens_.clear();
    special_reverse_.clear();
    std::istringstream iss(contents);
    std::string key;
    Token value;
    while (iss >> key >> value) { // "token id" per line (Tokenizer.h:482-485)
        special_tokens_[key] = value;
        special_reverse_[value] = key;
    }
    encoder_stale_ = true;
}

void Tokenizer::rebuild_vocab() { // Tokenizer.h:103-111 + :844-861
    vocab_.clear();
    vocab_.reserve(256 + merges_.size());
    for (int i = 0; i < 256; i++) vocab_.push_back(std::string(1, static_cast<char>(i)));
    for (const auto &[a, b] : merges_) vocab_.push_back(vocab_[a] + vocab_[b]);
}

static std::string printable(const std::string &bytes) { // Tokenizer.h:569-575
    std::string s;
    for (unsigned char c : bytes) s += (c >= 32 && c < 127) ? static_cast<char>(c) : ' ';
    return s;
}

int Tokenizer::train(std::string_view text, int vocab_size, CONFLICT_RESOLUTION mode, bool verbose) {
    if (vocab_size < 256) { // assert in the reference (Tokenizer.h:492)
        error_ = "vocab_size must be >= 256";
        return MBPE_E_INVALID;
    }
    merges_.clear();
    rebuild_vocab();
    encoder_stale_ = true;
    if (regex_.empty() && text.empty()) { // the reference dereferences an empty list here
        error_ = "empty training text";
        return MBPE_E_EMPTY;
    }
    const uint8_t *bytes = reinterpret_cast<const uint8_t *>(text.data());
    double t0 = now_s();
    int rc = MBPE_OK;
    mbpe_trainer *tr = nullptr;
    last_split_on_gpu = false;
    if (use_gpu_split(text.size())) { // split + dedup on the device; the corpus never comes back to the host
        mbpe_device_corpus dc;
        rc = mbpe_pretok_corpus(pretok_, bytes, text.size(), &dc);
        if (rc == MBPE_OK) {
            last_n_chunks = dc.n_chunks;
            last_n_unique = dc.n_unique;
            rc = mbpe_trainer_create_device(&dc, &tr);
            mbpe_device_corpus_free(&dc);
            last_split_on_gpu = rc == MBPE_OK;
        }
        if (rc != MBPE_OK && rc != MBPE_E_UNSUPPORTED) {
            error_ = mbpe_last_error();
            return rc;
        }
    }
    if (!tr) {
        Corpus corpus;
        rc = split_dedup_parallel(regex_, pattern_, bytes, text.size(), n_threads_, corpus, &error_);
        if (rc) return rc;
        last_n_chunks = corpus.n_chunks;
        last_n_unique = corpus.weight.size();
        rc = mbpe_trainer_create(corpus.tokens.data(), corpus.tokens.size(), corpus.off.data(), corpus.weight.size(),
                                 corpus.weight.data(), device_, &tr);
        if (rc) {
            error_ = mbpe_last_error();
            return rc;
        }
    }
    double t2 = now_s();
    if (verbose) std::cout << "Split input text into " << last_n_chunks << " chunks\n"; // Tokenizer.h:547
    last_split_s = t2 - t0; // regex split and dedup are one fused pass (host threads, or the device for the GPT-4 pattern)
    last_dedup_s = 0;

    const uint32_t n_target = static_cast<uint32_t>(vocab_size - 256);
    std::vector<uint32_t> m(2ull * std::max<uint32_t>(n_target, 1));
    std::vector<int32_t> counts(std::max<uint32_t>(n_target, 1));
    uint32_t n_merges = 0;
    rc = mbpe_trainer_run(tr, static_cast<uint32_t>(vocab_size), mode, engine_, nullptr, m.data(), counts.data(),
                          &n_merges, &last_stats);
    double t3 = now_s();
    mbpe_trainer_destroy(tr);
    if (getenv("MBPE_DEBUG"))
        fprintf(stderr, "[mbpe] Tokenizer::train: front end %.1f ms (%s), merge loop call %.1f ms (device %.1f ms), destroy %.1f ms\n",
                (t2 - t0) * 1e3, last_split_on_gpu ? "device" : "host", (t3 - t2) * 1e3, last_stats.gpu_ms, (now_s() - t3) * 1e3);
    if (rc) {
        error_ = mbpe_last_error();
        return rc;
    }
    merges_.reserve(n_merges);
    for (uint32_t i = 0; i < n_merges; i++) {
        Token a = m[2 * i], b = m[2 * i + 1];
        vocab_.push_back(vocab_[a] + vocab_[b]); // Tokenizer.h:562-564
        if (verbose) // Tokenizer.h:576
            std::cout << "merge " << i + 1 << "/" << n_target << ": (" << a << ", " << b << ") -> " << 256 + i << " (b'"
                      << printable(vocab_.back()) << "') had " << counts[i] << " occurrences\n";
        merges_.emplace_back(a, b);
    }
    if (verbose) // Tokenizer.h:591-597 prints the per-copy length; the GPU path keeps unique chunks only
        std::cout << "Length of training text " << text.length() << ". Unique chunks " << last_n_unique << " of "
                  << last_n_chunks << ", " << last_stats.n_positions << " resident tokens.\n";
    return MBPE_OK;
}

// Tokenizer.h:605-650 finds, from the current position, the next occurrence of EVERY special token, takes the earliest
// (ties: the first in the map's iteration order, :618-626) and repeats -- O(occurrences x tokens x text). Same result
// in one sweep per token: a token's next occurrence is searched again only once the position has
// passed the one already known, so every token scans the text once overall.
std::vector<SpecialPart> split_special_spans(std::string_view text,
                                             const std::unordered_map<std::string, Token> &specials) {
    std::vector<SpecialPart> parts;
    if (specials.empty()) {
        parts.push_back({0, text.size(), -1});
        return parts;
    }
    // Large texts: every occurrence of every token (overlapping ones too) is collected by all host threads over ranges
    // of the text; the reference's rule then is one pass over that list -- from the current position take the first
    // occurrence at or after it (equal positions: the token the map iterates first), jump behind it, repeat.
    const int n_threads = hardware_threads();
    if (text.size() >= (8u << 20) && n_threads > 1) {
        std::vector<const std::string *> toks;
        std::vector<Token> tids;
        for (const auto &kv : specials) {
            toks.push_back(&kv.first);
            tids.push_back(kv.second);
        }
        const size_t T = (size_t)std::min(n_threads, 32), step = (text.size() + T - 1) / T;
        std::vector<std::vector<std::pair<size_t, uint32_t>>> found(T);
        std::vector<std::thread> th;
        for (size_t t = 0; t < T; t++)
            th.emplace_back([&, t]() {
                const size_t a = t * step, b = std::min(text.size(), a + step);
                for (uint32_t k = 0; k < toks.size() && a < b; k++) {
                    const std::string &tok = *toks[k];
                    if (tok.empty()) continue;
                    const std::string_view sub = text.substr(a, (b - a) + tok.size() - 1); // occurrences that START in [a, b)
                    for (size_t q = sub.find(tok); q != std::string_view::npos; q = sub.find(tok, q + 1)) found[t].emplace_back(a + q, k);
                }
                std::sort(found[t].begin(), found[t].end());
            });
        for (auto &x : th) x.join();
        size_t pos = 0, last = 0;
        for (size_t t = 0; t < T; t++)
            for (const auto &[at, k] : found[t]) {
                if (at < pos) continue; // overlaps the occurrence taken before it
                if (at > last) parts.push_back({last, at, -1});
                parts.push_back({at, at + toks[k]->size(), (int64_t)tids[k]});
                pos = last = at + toks[k]->size();
            }
        if (last < text.size()) parts.push_back({last, text.size(), -1});
        if (parts.empty()) parts.push_back({0, text.size(), -1});
        return parts;
    }
    struct Next {
        const std::string *tok;
        Token id;
        size_t at; // next occurrence at or after the position it was searched from; npos = none left
    };
    std::vector<Next> next;
    next.reserve(specials.size());
    for (const auto &kv : specials) next.push_back({&kv.first, kv.second, 0}); // map iteration order = tie order
    for (auto &n : next) n.at = n.tok->empty() ? std::string::npos : text.find(*n.tok, 0);
    size_t pos = 0, last = 0;
    while (pos < text.size()) {
        const Next *best = nullptr;
        for (auto &n : next) {
            if (n.at != std::string::npos && n.at < pos) n.at = text.find(*n.tok, pos); // stale: search on from pos
            if (n.at != std::string::npos && (!best || n.at < best->at)) best = &n;
        }
        if (!best) break;
        if (best->at > last) parts.push_back({last, best->at, -1});
        parts.push_back({best->at, best->at + best->tok->size(), (int64_t)best->id});
        pos = best->at + best->tok->size();
        last = pos;
    }
    if (last < text.size()) parts.push_back({last, text.size(), -1});
    if (parts.empty()) parts.push_back({0, text.size(), -1});
    return parts;
}

std::vector<std::string> Tokenizer::split_on_special(std::string_view text) {
    std::vector<std::string> result;
    for (const auto &p : split_special_spans(text, special_tokens_)) {
        if (p.id < 0) {
            result.emplace_back(text.substr(p.start, p.end - p.start));
        } else { // marker "\0<id>" (Tokenizer.h:631-634)
            std::string marker(1, '\0');
            marker += std::to_string(p.id);
            result.push_back(std::move(marker));
        }
    }
    return result;
}

int Tokenizer::ensure_encoder() {
    if (encoder_ && !encoder_stale_) return MBPE_OK;
    if (encoder_) {
        mbpe_encoder_destroy(encoder_);
        encoder_ = nullptr;
    }
    std::vector<uint32_t> m;
    m.reserve(2 * merges_.size());
    for (const auto &[a, b] : merges_) {
        m.push_back(a);
        m.push_back(b);
    }
    int rc = mbpe_encoder_create(m.data(), static_cast<uint32_t>(merges_.size()), device_, &encoder_);
    if (rc) {
        error_ = mbpe_last_error();
        return rc;
    }
    std::vector<uint32_t> ids;
    std::vector<uint8_t> bytes;
    std::vector<uint64_t> off{0};
    for (const auto &kv : special_reverse_) {
        ids.push_back(kv.first);
        bytes.insert(bytes.end(), kv.second.begin(), kv.second.end());
        off.push_back(bytes.size());
    }
    if (!ids.empty()) {
        rc = mbpe_encoder_set_specials(encoder_, ids.data(), bytes.data(), off.data(), static_cast<uint32_t>(ids.size()));
        if (rc) {
            error_ = mbpe_last_error();
            return rc;
        }
    }
    // encode side: "a chunk with exactly these bytes is this id" (Tokenizer.h:667-671) for the device front end
    specials_on_device_ = false;
    if (!special_tokens_.empty()) {
        std::vector<uint32_t> sids;
        std::vector<uint8_t> sbytes;
        std::vector<uint64_t> soff{0};
        for (const auto &kv : special_tokens_) {
            sids.push_back(kv.second);
            sbytes.insert(sbytes.end(), kv.first.begin(), kv.first.end());
            soff.push_back(sbytes.size());
        }
        rc = mbpe_encoder_seed_special_chunks(encoder_, sids.data(), sbytes.data(), soff.data(), static_cast<uint32_t>(sids.size()));
        if (rc != MBPE_OK && rc != MBPE_E_UNSUPPORTED) {
            error_ = mbpe_last_error();
            return rc;
        }
        specials_on_device_ = rc == MBPE_OK;
    }
    encoder_stale_ = false;
    return MBPE_OK;
}

// Tokenizer::encode straight into a caller buffer: for a text without special tokens under the GPT-4 pattern the ids
// come down from the device into `out` with no intermediate copy; everything else goes through encode() above.
// text up, ids down with split and merge scan both on the device (built-in patterns; special tokens if the encoder
// knows them as ready-made chunks). MBPE_E_UNSUPPORTED: not applicable to this tokenizer / text -- use the host path.
int Tokenizer::encode_on_device(std::string_view text, Token *out, uint64_t cap, uint64_t *n_out) {
    if (!(special_tokens_.empty() || specials_on_device_) || !use_gpu_split(text.size())) {
        if (getenv("MBPE_DEBUG"))
            fprintf(stderr, "[mbpe] encode: host front end (specials %zu, seeded %d, device matcher %d)\n", special_tokens_.size(),
                    (int)specials_on_device_, (int)(pretok_ != nullptr));
        return MBPE_E_UNSUPPORTED;
    }
    int rc;
    if (special_tokens_.empty()) {
        rc = mbpe_encode_text(encoder_, pretok_, reinterpret_cast<const uint8_t *>(text.data()), text.size(), out, cap, n_out);
    } else { // the occurrences of special tokens are found here (one sweep per token); everything else on the device
        std::vector<uint64_t> sb, se;
        for (const auto &part : split_special_spans(text, special_tokens_))
            if (part.id >= 0) {
                sb.push_back(part.start);
                se.push_back(part.end);
            }
        rc = mbpe_encode_text_special(encoder_, pretok_, reinterpret_cast<const uint8_t *>(text.data()), text.size(), sb.data(),
                                      se.data(), sb.size(), out, cap, n_out);
    }
    if (rc != MBPE_OK && rc != MBPE_E_UNSUPPORTED) error_ = mbpe_last_error();
    if (rc == MBPE_E_UNSUPPORTED && getenv("MBPE_DEBUG")) fprintf(stderr, "[mbpe] encode: device front end declined: %s\n", mbpe_last_error());
    return rc;
}

int Tokenizer::encode_into(std::string_view text, Token *out, uint64_t cap, uint64_t *n_out) {
    *n_out = 0;
    int rc = ensure_encoder();
    if (rc) return rc;
    if (out) {
        rc = encode_on_device(text, out, cap, n_out);
        if (rc != MBPE_E_UNSUPPORTED) return rc;
    }
    std::vector<Token> ids;
    if ((rc = encode_host(text, false, ids))) return rc;
    *n_out = ids.size();
    if (!out) return MBPE_OK;
    if (ids.size() > cap) {
        error_ = "out too small; *n_out holds the needed count";
        return MBPE_E_CAPACITY;
    }
    memcpy(out, ids.data(), ids.size() * sizeof(Token));
    return MBPE_OK;
}

// File to file without holding either in memory (examples/minbpe-cc.cpp:212-242 slurps and writes id by id): only for
// the GPT-4 pattern without special tokens; MBPE_E_UNSUPPORTED tells the caller to take the whole-file path.
int Tokenizer::encode_file(const std::string &in_path, const std::string &out_path, uint64_t *n_ids) {
    int rc = ensure_encoder();
    if (rc) return rc;
    if (!special_tokens_.empty() || !use_gpu_split(~size_t(0) >> 1)) return MBPE_E_UNSUPPORTED;
    rc = mbpe_encode_file(encoder_, pretok_, in_path.c_str(), out_path.c_str(), nullptr, n_ids);
    if (rc && rc != MBPE_E_UNSUPPORTED) error_ = mbpe_last_error();
    return rc;
}

int Tokenizer::encode(std::string_view text, bool verbose, std::vector<Token> &out) {
    out.clear();
    int rc = ensure_encoder();
    if (rc) return rc;
    if (!verbose) { // (the verbose run reports the parts of the host path)
        std::vector<Token> ids(std::max<size_t>(text.size(), 1));
        uint64_t n_ids = 0;
        rc = encode_on_device(text, ids.data(), ids.size(), &n_ids);
        if (rc == MBPE_OK) {
            ids.resize(n_ids);
            out.swap(ids);
            return MBPE_OK;
        }
        if (rc != MBPE_E_UNSUPPORTED) return rc;
    }
    return encode_host(text, verbose, out);
}

// the general path: special tokens and the regex on the host (any pattern), the merge scan on the device
int Tokenizer::encode_host(std::string_view text, bool verbose, std::vector<Token> &out) {
    out.clear();
    int rc = ensure_encoder();
    if (rc) return rc;
    std::vector<std::string> parts;
    bool single = special_tokens_.empty();
    if (!single) parts = split_on_special(text);
    const size_t n_parts = single ? 1 : parts.size();
    if (verbose) std::cout << "Splitting input text into " << n_parts << " parts\n";

    // chunk list over one byte arena; ready-made ids (special markers) are spliced in afterwards
    std::vector<uint8_t> arena_copy;
    const uint8_t *arena = nullptr;
    std::vector<uint64_t> off{0};
    std::vector<std::pair<uint64_t, Token>> ready; // (index of the chunk this id precedes, id)
    std::vector<Span> spans;
    auto add_chunk = [&](std::string_view chunk) {
        Token id;
        if (!chunk.empty() && chunk[0] == '\0' && marker_token(chunk, &id)) {
            ready.emplace_back(off.size() - 1, id);
            return;
        }
        arena_copy.insert(arena_copy.end(), chunk.begin(), chunk.end());
        off.push_back(arena_copy.size());
    };
    if (single && !regex_.empty()) {
        const uint8_t *bytes = reinterpret_cast<const uint8_t *>(text.data());
        rc = split_parallel(regex_, pattern_, bytes, text.size(), n_threads_, spans, &error_);
        if (rc) return rc;
        bool contiguous = !spans.empty() && spans.front().start == 0 && spans.back().end == text.size();
        for (size_t i = 0; contiguous && i < spans.size(); i++) {
            if (i && spans[i].start != spans[i - 1].end) contiguous = false;
            if (bytes[spans[i].start] == 0) contiguous = false; // possible ready-made id: take the general path
        }
        if (contiguous) { // the usual case: the text itself is the arena, no copy
            arena = bytes;
            off.resize(spans.size() + 1);
            for (size_t i = 0; i < spans.size(); i++) off[i + 1] = spans[i].end;
        } else {
            for (const auto &s : spans) add_chunk(text.substr(s.start, s.end - s.start));
        }
    } else {
        for (size_t p = 0; p < n_parts; p++) {
            std::string_view part = single ? text : std::string_view(parts[p]);
            bool special = !part.empty() && part[0] == '\0';
            if (special || regex_.empty()) { // one chunk (Tokenizer.h:667-671, :706-709)
                add_chunk(part);
                continue;
            }
            spans.clear();
            rc = split_parallel(regex_, pattern_, reinterpret_cast<const uint8_t *>(part.data()), part.size(), n_threads_,
                                spans, &error_);
            if (rc) return rc;
            for (const auto &s : spans) add_chunk(part.substr(s.start, s.end - s.start));
        }
    }
    if (!arena) arena = arena_copy.data();
    const uint64_t n_chunks = off.size() - 1, n_bytes = off.back();
    std::vector<Token> ids(std::max<uint64_t>(n_bytes, 1));
    std::vector<uint64_t> out_off;
    if (!ready.empty()) out_off.resize(n_chunks + 1);
    uint64_t n_ids = 0;
    rc = mbpe_encode(encoder_, arena, n_bytes, off.data(), n_chunks, ids.data(), ids.size(), &n_ids,
                     ready.empty() ? nullptr : out_off.data());
    if (rc) {
        error_ = mbpe_last_error();
        return rc;
    }
    if (ready.empty()) {
        ids.resize(n_ids);
        out.swap(ids);
    } else {
        out.reserve(n_ids + ready.size());
        uint64_t done = 0;
        for (const auto &[before_chunk, id] : ready) {
            uint64_t upto = out_off[before_chunk];
            out.insert(out.end(), ids.begin() + done, ids.begin() + upto);
            done = upto;
            out.push_back(id);
        }
        out.insert(out.end(), ids.begin() + done, ids.begin() + n_ids);
    }
    if (verbose) std::cout << "Encoded input text (length " << text.length() << ") to " << out.size() << " tokens\n";
    return MBPE_OK;
}

int Tokenizer::decode(const std::vector<Token> &tokens, bool verbose, std::string &out) {
    out.clear();
    if (verbose) std::cout << "Decoding " << tokens.size() << " tokens\n";
    int rc = ensure_encoder();
    if (rc) return rc;
    for (Token t : tokens) // same diagnostics as the reference (Tokenizer.h:739-742)
        if (t >= vocab_.size() && special_reverse_.find(t) == special_reverse_.end())
            std::cerr << "Warning: Attempted to decode invalid token ID: " << t << "\n";
    // one call when the guess (4 bytes per id; text averages 2-3) is large enough, else a second one at the exact size
    uint64_t n = 0;
    out.resize(tokens.size() * 4 + 64);
    rc = mbpe_decode(encoder_, tokens.data(), tokens.size(), reinterpret_cast<uint8_t *>(out.data()), out.size(), &n);
    if (rc == MBPE_E_CAPACITY) {
        out.resize(n);
        rc = mbpe_decode(encoder_, tokens.data(), tokens.size(), reinterpret_cast<uint8_t *>(out.data()), out.size(), &n);
    }
    if (rc) {
        error_ = mbpe_last_error();
        out.clear();
        return rc;
    }
    out.resize(n);
    return MBPE_OK;
}

// .enc file -> text file in blocks (examples/minbpe-cc.cpp:243-258 slurps both), any model
int Tokenizer::decode_file(const std::string &in_path, const std::string &out_path, uint64_t *n_ids, uint64_t *n_bytes) {
    int rc = ensure_encoder();
    if (rc) return rc;
    rc = mbpe_decode_file(encoder_, in_path.c_str(), out_path.c_str(), n_ids, n_bytes);
    if (rc) error_ = mbpe_last_error();
    return rc;
}

int Tokenizer::load(const std::string &path, bool verbose) {
    std::ifstream in(path, std::ios::in);
    if (!in.is_open()) {
        std::cerr << "Failed to open file for loading: " << path << "\n";
        error_ = "cannot open " + path;
        return MBPE_E_IO;
    }
    std::string version;
    std::getline(in, version);
    if (version != "minbpe v1") {
        std::cerr << "Unexpected version: " << version << "\n";
        error_ = "unexpected version line";
        return MBPE_E_IO;
    }
    // Everything is parsed into locals and committed only when the whole file was good: a failed load leaves the
    // tokenizer exactly as it was (the reference returns false half-way with its members already overwritten).
    std::string new_pattern;
    std::getline(in, new_pattern);
    Regex new_regex;
    if (new_regex.compile(new_pattern, &error_) != MBPE_OK) {
        std::cerr << "PCRE2 compilation failed on load: " << error_ << "\n";
        return MBPE_E_REGEX;
    }
    int num_special = 0;
    in >> num_special;
    std::vector<std::pair<std::string, Token>> new_specials;
    for (int i = 0; i < num_special; i++) {
        std::string token;
        Token id;
        in >> token >> id;
        new_specials.emplace_back(token, id);
        if (verbose) std::cout << "Loaded special token: " << token << " with ID " << id << "\n";
    }
    std::vector<std::pair<Token, Token>> new_merges;
    Token a, b;
    while (in >> a >> b) {
        if (a >= 256 + new_merges.size() || b >= 256 + new_merges.size()) { // the reference would index past vocab
            error_ = "merge line names an id that does not exist yet";
            return MBPE_E_IO;
        }
        new_merges.emplace_back(a, b);
    }
    pattern_ = std::move(new_pattern);
    regex_ = std::move(new_regex);
    for (auto &[token, id] : new_specials) { // existing specials are kept, as in the reference
        special_tokens_[token] = id;
        special_reverse_[id] = token;
    }
    merges_ = std::move(new_merges);
    if (verbose) std::cout << "Read input model from " << path << "\n";
    rebuild_vocab();
    if (verbose)
        std::cout << "Loaded vocab with " << merges_.size() << " merges, vocab size is " << vocab_.size() << "\n";
    encoder_stale_ = true; // device tables, and (use_gpu_split) the device matcher's pattern, follow on next use
    return MBPE_OK;
}

int write_model_files(const std::string &path, const std::string &pattern,
                      const std::unordered_map<std::string, Token> &specials,
                      const std::vector<std::pair<Token, Token>> &merges, const std::vector<std::string> *vocab,
                      std::string *err) {
    std::ofstream out(path, std::ios::out);
    if (!out.is_open()) {
        std::cerr << "Unable to open file for saving: " << path << std::endl;
        if (err) *err = "cannot open " + path;
        return MBPE_E_IO;
    }
    std::cout << "Writing model...\n";
    out << "minbpe v1" << '\n' << pattern << '\n' << specials.size() << '\n';
    for (const auto &st : specials) out << st.first << ' ' << st.second << '\n'; // map iteration order
    for (const auto &[a, b] : merges) out << a << ' ' << b << "\n";
    out.close();
    if (vocab) {
        const std::string vpath = path + ".vocab"; // appended, not an extension swap (Tokenizer.h:896)
        std::ofstream vf(vpath, std::ios::out);
        if (!vf.is_open()) {
            std::cerr << "Failed to open .vocab file for writing: " << vpath << std::endl;
            if (err) *err = "cannot open " + vpath;
            return MBPE_E_IO;
        }
        Token id = 0;
        for (const auto &v : *vocab) { // Tokenizer.h:905-917
            vf << std::setw(6) << std::left << id << ": \"";
            for (unsigned char c : v) {
                if (c >= 32 && c <= 126)
                    vf << static_cast<char>(c);
                else
                    vf << "\xEF\xBF\xBD";
            }
            vf << "\"\n";
            id++;
        }
    }
    std::cout << "Complete.\n";
    return MBPE_OK;
}

// ---------------------------------------------------------------------------------------------------------
// .vocab in karpathy/minbpe's layout (base.py: save + render_token), the format the reference's README compares its
// models with; the reference's own .vocab layout (Tokenizer.h:894-918) stays the default.
//   [<left>][<right>] -> [<token>] <id>      for a merged token
//   [<token>] <id>                             for a byte / special token
// <token> = the bytes decoded as UTF-8 with U+FFFD for every maximal invalid subpart (Python's errors="replace"), and
// every code point of general category C* written as \uXXXX. The categories are the linked PCRE2's (\p{C}).
// ---------------------------------------------------------------------------------------------------------
namespace {
struct ControlClass { // \p{C} of the linked PCRE2, asked once per code point and remembered
    pcre2_code_8 *code = nullptr;
    pcre2_match_data_8 *md = nullptr;
    std::unordered_map<uint32_t, bool> memo;
    ControlClass() {
        int ec = 0;
        size_t eo = 0;
        code = pcre2_compile_8(reinterpret_cast<const uint8_t *>("\\p{C}"), 5, MBPE_PCRE2_UTF | MBPE_PCRE2_UCP, &ec, &eo, nullptr);
        if (code) md = pcre2_match_data_create_from_pattern_8(code, nullptr);
    }
    ~ControlClass() {
        if (md) pcre2_match_data_free_8(md);
        if (code) pcre2_code_free_8(code);
    }
    bool is_c(uint32_t cp, const std::string &utf8) {
        if (cp < 0x20 || (cp >= 0x7F && cp < 0xA0)) return true; // Cc
        if (cp < 0x7F) return false;
        auto it = memo.find(cp);
        if (it != memo.end()) return it->second;
        const bool c = code && md && pcre2_match_8(code, reinterpret_cast<const uint8_t *>(utf8.data()), utf8.size(), 0, 0, md, nullptr) >= 0;
        memo[cp] = c;
        return c;
    }
};

std::string render_token(const std::string &t, ControlClass &cc) {
    std::string out;
    const size_t n = t.size();
    auto cont = [&](size_t i, unsigned lo, unsigned hi) { return i < n && (unsigned char)t[i] >= lo && (unsigned char)t[i] <= hi; };
    auto emit = [&](uint32_t cp, const std::string &utf8) {
        if (cc.is_c(cp, utf8)) {
            char buf[16];
            snprintf(buf, sizeof buf, "\\u%04x", cp);
            out += buf;
        } else {
            out += utf8;
        }
    };
    for (size_t i = 0; i < n;) {
        const unsigned b = (unsigned char)t[i];
        size_t len = 0; // bytes of a well-formed sequence starting at i, 0 = ill-formed
        size_t bad = 1; // ill-formed: length of the maximal subpart to replace by one U+FFFD
        if (b < 0x80) {
            len = 1;
        } else if (b >= 0xC2 && b <= 0xDF) {
            if (cont(i + 1, 0x80, 0xBF)) len = 2;
        } else if (b >= 0xE0 && b <= 0xEF) {
            const unsigned lo = b == 0xE0 ? 0xA0 : 0x80, hi = b == 0xED ? 0x9F : 0xBF;
            if (cont(i + 1, lo, hi)) {
                if (cont(i + 2, 0x80, 0xBF)) len = 3;
                else bad = 2;
            }
        } else if (b >= 0xF0 && b <= 0xF4) {
            const unsigned lo = b == 0xF0 ? 0x90 : 0x80, hi = b == 0xF4 ? 0x8F : 0xBF;
            if (cont(i + 1, lo, hi)) {
                if (cont(i + 2, 0x80, 0xBF)) {
                    if (cont(i + 3, 0x80, 0xBF)) len = 4;
                    else bad = 3;
                } else {
                    bad = 2;
                }
            }
        }
        if (len == 0) {
            out += "\xEF\xBF\xBD";
            i += bad;
            continue;
        }
        uint32_t cp = b;
        if (len == 2) cp = ((b & 0x1F) << 6) | ((unsigned char)t[i + 1] & 0x3F);
        else if (len == 3) cp = ((b & 0x0F) << 12) | (((unsigned char)t[i + 1] & 0x3F) << 6) | ((unsigned char)t[i + 2] & 0x3F);
        else if (len == 4)
            cp = ((b & 0x07) << 18) | (((unsigned char)t[i + 1] & 0x3F) << 12) | (((unsigned char)t[i + 2] & 0x3F) << 6) |
                 ((unsigned char)t[i + 3] & 0x3F);
        emit(cp, t.substr(i, len));
        i += len;
    }
    return out;
}
} // namespace

// specials: (token, id) in the order they should be listed (karpathy lists them in registration order, after the merges)
int write_vocab_karpathy(const std::string &path, const std::vector<std::pair<std::string, Token>> &specials,
                         const std::vector<std::pair<Token, Token>> &merges, const std::vector<std::string> &vocab,
                         std::string *err) {
    std::ofstream vf(path, std::ios::out | std::ios::binary);
    if (!vf.is_open()) {
        std::cerr << "Failed to open .vocab file for writing: " << path << std::endl;
        if (err) *err = "cannot open " + path;
        return MBPE_E_IO;
    }
    ControlClass cc;
    for (size_t id = 0; id < vocab.size(); id++) {
        const std::string s = render_token(vocab[id], cc);
        if (id >= 256 && id - 256 < merges.size()) {
            const auto &[a, b] = merges[id - 256];
            vf << '[' << render_token(vocab[a], cc) << "][" << render_token(vocab[b], cc) << "] -> [" << s << "] " << id << '\n';
        } else {
            vf << '[' << s << "] " << id << '\n';
        }
    }
    for (const auto &[tok, id] : specials) vf << '[' << render_token(tok, cc) << "] " << id << '\n';
    vf.close();
    if (!vf) {
        if (err) *err = "write failed: " + path;
        return MBPE_E_IO;
    }
    return MBPE_OK;
}

int Tokenizer::save_vocab_karpathy(const std::string &path) {
    if (merges_.empty()) {
        error_ = "no merges to save";
        return MBPE_E_EMPTY;
    }
    std::vector<std::pair<std::string, Token>> sp(special_tokens_.begin(), special_tokens_.end());
    std::sort(sp.begin(), sp.end(), [](const auto &x, const auto &y) { return x.second < y.second; }); // by id: deterministic
    return write_vocab_karpathy(path, sp, merges_, vocab_, &error_);
}

int Tokenizer::save(const std::string &path, bool write_vocab) {
    if (merges_.empty()) { // assert in the reference (Tokenizer.h:876)
        error_ = "no merges to save";
        return MBPE_E_EMPTY;
    }
    return write_model_files(path, pattern_, special_tokens_, merges_, write_vocab ? &vocab_ : nullptr, &error_);
}

} // namespace mbpe::host

// ---------------------------------------------------------------------------------------------------------
// C ABI: tokenizer mirror + host-only helpers
// ---------------------------------------------------------------------------------------------------------
using namespace mbpe::host;

namespace mbpe {
std::string &last_error_ref();
}
static int fail(int code, const std::string &msg) {
    mbpe::last_error_ref() = msg;
    return code;
}

struct mbpe_tokenizer {
    Tokenizer tk;
    mbpe_tokenizer(const char *p, int device) : tk(p ? p : "", device) {}
};

extern "C" const char *mbpe_gpt2_split_pattern(void) { return kGpt2Pattern; }
extern "C" const char *mbpe_gpt4_split_pattern(void) { return kGpt4Pattern; }

extern "C" int mbpe_tokenizer_create(const char *pattern, int device, mbpe_tokenizer **out) {
    if (!out) return fail(MBPE_E_INVALID, "null argument");
    *out = nullptr;
    try {
        *out = new mbpe_tokenizer(pattern, device);
    } catch (const std::exception &e) {
        return fail(MBPE_E_REGEX, e.what());
    }
    return MBPE_OK;
}
extern "C" void mbpe_tokenizer_destroy(mbpe_tokenizer *t) { delete t; }
extern "C" int mbpe_tokenizer_set_special_tokens(mbpe_tokenizer *t, const char *contents, uint64_t len) {
    if (!t || (!contents && len)) return fail(MBPE_E_INVALID, "null argument");
    t->tk.set_special_tokens_from_file(std::string(contents ? contents : "", len));
    return MBPE_OK;
}
extern "C" int mbpe_tokenizer_train(mbpe_tokenizer *t, const uint8_t *text, uint64_t len, int vocab_size, int mode,
                                    int verbose) {
    if (!t || (!text && len)) return fail(MBPE_E_INVALID, "null argument");
    if (mode != MBPE_MODE_FIRST && mode != MBPE_MODE_LEXICAL) return fail(MBPE_E_INVALID, "bad mode");
    int rc = t->tk.train(std::string_view(reinterpret_cast<const char *>(text), len), vocab_size,
                         static_cast<Tokenizer::CONFLICT_RESOLUTION>(mode), verbose != 0);
    return rc ? fail(rc, t->tk.error()) : MBPE_OK;
}
extern "C" int mbpe_tokenizer_save(mbpe_tokenizer *t, const char *path, int write_vocab) {
    if (!t || !path) return fail(MBPE_E_INVALID, "null argument");
    int rc = t->tk.save(path, write_vocab != 0);
    return rc ? fail(rc, t->tk.error()) : MBPE_OK;
}
extern "C" int mbpe_tokenizer_load(mbpe_tokenizer *t, const char *path, int verbose) {
    if (!t || !path) return fail(MBPE_E_INVALID, "null argument");
    int rc = t->tk.load(path, verbose != 0);
    return rc ? fail(rc, t->tk.error()) : MBPE_OK;
}
extern "C" int mbpe_tokenizer_encode(mbpe_tokenizer *t, const uint8_t *text, uint64_t len, uint32_t *out,
                                     uint64_t out_cap, uint64_t *n_out) {
    if (!t || !n_out || (!text && len)) return fail(MBPE_E_INVALID, "null argument");
    int rc = t->tk.encode_into(std::string_view(reinterpret_cast<const char *>(text), len), out, out_cap, n_out);
    return rc ? fail(rc, t->tk.error()) : MBPE_OK;
}
extern "C" int mbpe_tokenizer_encode_file(mbpe_tokenizer *t, const char *in_path, const char *out_path, uint64_t *n_ids) {
    if (!t || !in_path || !out_path) return fail(MBPE_E_INVALID, "null argument");
    int rc = t->tk.encode_file(in_path, out_path, n_ids);
    if (rc == MBPE_E_UNSUPPORTED) return fail(rc, "streaming encode needs the GPT-4 pattern, no special tokens and well-formed UTF-8");
    return rc ? fail(rc, t->tk.error()) : MBPE_OK;
}
extern "C" int mbpe_tokenizer_decode_file(mbpe_tokenizer *t, const char *in_path, const char *out_path, uint64_t *n_ids,
                                          uint64_t *n_bytes) {
    if (!t || !in_path || !out_path) return fail(MBPE_E_INVALID, "null argument");
    int rc = t->tk.decode_file(in_path, out_path, n_ids, n_bytes);
    return rc ? fail(rc, t->tk.error()) : MBPE_OK;
}
extern "C" int mbpe_tokenizer_decode(mbpe_tokenizer *t, const uint32_t *ids, uint64_t n, uint8_t *out,
                                     uint64_t out_cap, uint64_t *n_out) {
    if (!t || !n_out || (!ids && n)) return fail(MBPE_E_INVALID, "null argument");
    std::string s;
    int rc = t->tk.decode(std::vector<Token>(ids, ids + n), false, s);
    if (rc) return fail(rc, t->tk.error());
    *n_out = s.size();
    if (!out) return MBPE_OK;
    if (s.size() > out_cap) return fail(MBPE_E_CAPACITY, "out too small; *n_out holds the needed size");
    memcpy(out, s.data(), s.size());
    return MBPE_OK;
}
extern "C" int mbpe_tokenizer_get_merges(mbpe_tokenizer *t, uint32_t *merges_out, uint32_t cap_pairs,
                                         uint32_t *n_merges) {
    if (!t || !n_merges) return fail(MBPE_E_INVALID, "null argument");
    const auto &m = t->tk.merges();
    *n_merges = static_cast<uint32_t>(m.size());
    if (!merges_out) return MBPE_OK;
    if (m.size() > cap_pairs) return fail(MBPE_E_CAPACITY, "merges_out too small");
    for (size_t i = 0; i < m.size(); i++) {
        merges_out[2 * i] = m[i].first;
        merges_out[2 * i + 1] = m[i].second;
    }
    return MBPE_OK;
}
extern "C" int mbpe_tokenizer_last_train_stats(mbpe_tokenizer *t, mbpe_train_stats *stats, double *split_s,
                                               double *dedup_s, uint64_t *n_chunks, uint64_t *n_unique) {
    if (!t) return fail(MBPE_E_INVALID, "null argument");
    if (stats) *stats = t->tk.last_stats;
    if (split_s) *split_s = t->tk.last_split_s;
    if (dedup_s) *dedup_s = t->tk.last_dedup_s;
    if (n_chunks) *n_chunks = t->tk.last_n_chunks;
    if (n_unique) *n_unique = t->tk.last_n_unique;
    return MBPE_OK;
}
extern "C" int mbpe_tokenizer_last_split_on_gpu(mbpe_tokenizer *t) { return t && t->tk.last_split_on_gpu ? 1 : 0; }
extern "C" void mbpe_tokenizer_set_engine(mbpe_tokenizer *t, int engine) {
    if (t) t->tk.set_engine(engine);
}
extern "C" void mbpe_tokenizer_set_threads(mbpe_tokenizer *t, int n_threads) {
    if (t) t->tk.set_threads(n_threads);
}

extern "C" int mbpe_split(const char *pattern, const uint8_t *text, uint64_t len, int n_threads, uint64_t *starts,
                          uint64_t *ends, uint64_t cap, uint64_t *n_chunks) {
    if (!pattern || !n_chunks || (!text && len)) return fail(MBPE_E_INVALID, "null argument");
    Regex re;
    std::string err;
    int rc = re.compile(pattern, &err);
    if (rc) return fail(rc, err);
    std::vector<Span> spans;
    rc = split_parallel(re, pattern, text, len, n_threads, spans, &err);
    if (rc) return fail(rc, err);
    *n_chunks = spans.size();
    if (!starts || !ends) return MBPE_OK;
    if (spans.size() > cap) return fail(MBPE_E_CAPACITY, "starts/ends too small");
    for (size_t i = 0; i < spans.size(); i++) {
        starts[i] = spans[i].start;
        ends[i] = spans[i].end;
    }
    return MBPE_OK;
}

extern "C" int mbpe_special_split(const char *special_contents, uint64_t special_len, const uint8_t *text, uint64_t len,
                                  uint64_t *starts, uint64_t *ends, int64_t *ids, uint64_t cap, uint64_t *n_parts) {
    if (!n_parts || (!text && len) || (!special_contents && special_len)) return fail(MBPE_E_INVALID, "null argument");
    std::unordered_map<std::string, Token> specials; // filled exactly as set_special_tokens_from_file does (:482-485)
    std::istringstream iss(std::string(special_contents ? special_contents : "", special_len));
    std::string key;
    Token value;
    while (iss >> key >> value) specials[key] = value;
    const auto parts = split_special_spans(std::string_view(reinterpret_cast<const char *>(text), len), specials);
    *n_parts = parts.size();
    if (!starts || !ends || !ids) return MBPE_OK;
    if (parts.size() > cap) return fail(MBPE_E_CAPACITY, "part arrays too small");
    for (size_t i = 0; i < parts.size(); i++) {
        starts[i] = parts[i].start;
        ends[i] = parts[i].end;
        ids[i] = parts[i].id;
    }
    return MBPE_OK;
}

extern "C" int mbpe_pretok_class_table(uint8_t *table_out) {
    if (!table_out) return fail(MBPE_E_INVALID, "null argument");
    std::string err;
    int rc = pretok_class_table(table_out, &err);
    return rc ? fail(rc, err) : MBPE_OK;
}

extern "C" int mbpe_dedup(const uint8_t *text, const uint64_t *starts, const uint64_t *ends, uint64_t n_chunks,
                          uint32_t *tokens_out, uint64_t *n_tokens, uint64_t *off_out, uint32_t *weight_out,
                          uint64_t *n_unique) {
    if (!n_tokens || !n_unique || (n_chunks && (!text || !starts || !ends))) return fail(MBPE_E_INVALID, "null argument");
    std::vector<Span> spans(n_chunks);
    for (uint64_t i = 0; i < n_chunks; i++) spans[i] = Span{starts[i], ends[i]};
    Corpus c;
    dedup_chunks(text, spans, 0, c);
    *n_tokens = c.tokens.size();
    *n_unique = c.weight.size();
    if (tokens_out) memcpy(tokens_out, c.tokens.data(), c.tokens.size() * 4);
    if (off_out) memcpy(off_out, c.off.data(), c.off.size() * 8);
    if (weight_out) memcpy(weight_out, c.weight.data(), c.weight.size() * 4);
    return MBPE_OK;
}

extern "C" int mbpe_split_dedup(const char *pattern, const uint8_t *text, uint64_t len, int n_threads,
                                uint32_t *tokens_out, uint64_t tokens_cap, uint64_t *n_tokens, uint64_t *off_out,
                                uint32_t *weight_out, uint64_t unique_cap, uint64_t *n_unique, uint64_t *n_chunks) {
    if (!pattern || !n_tokens || !n_unique || (!text && len)) return fail(MBPE_E_INVALID, "null argument");
    static thread_local Corpus cached; // sizing call followed by the copying call: do the work once
    static thread_local const uint8_t *cached_text = nullptr;
    static thread_local uint64_t cached_len = 0;
    static thread_local std::string cached_pattern;
    if (!(tokens_out && cached_text == text && cached_len == len && cached_pattern == pattern)) {
        Regex re;
        std::string err;
        int rc = re.compile(pattern, &err);
        if (rc) return fail(rc, err);
        rc = split_dedup_parallel(re, pattern, text, len, n_threads, cached, &err);
        if (rc) return fail(rc, err);
        cached_text = text;
        cached_len = len;
        cached_pattern = pattern;
    }
    *n_tokens = cached.tokens.size();
    *n_unique = cached.weight.size();
    if (n_chunks) *n_chunks = cached.n_chunks;
    if (!tokens_out) return MBPE_OK;
    if (cached.tokens.size() > tokens_cap || cached.weight.size() > unique_cap)
        return fail(MBPE_E_CAPACITY, "output buffers too small");
    memcpy(tokens_out, cached.tokens.data(), cached.tokens.size() * 4);
    if (off_out) memcpy(off_out, cached.off.data(), cached.off.size() * 8);
    if (weight_out) memcpy(weight_out, cached.weight.data(), cached.weight.size() * 4);
    cached = Corpus();
    cached_text = nullptr;
    return MBPE_OK;
}

extern "C" int mbpe_write_model(const char *path, const char *pattern, const char *special_contents,
                                uint64_t special_len, const uint32_t *merges, uint32_t n_merges, int write_vocab) {
    if (!path || !pattern || (n_merges && !merges)) return fail(MBPE_E_INVALID, "null argument");
    if (n_merges == 0) return fail(MBPE_E_EMPTY, "no merges to save");
    std::unordered_map<std::string, Token> specials; // filled exactly as set_special_tokens_from_file does
    if (special_contents && special_len) {
        std::istringstream iss(std::string(special_contents, special_len));
        std::string key;
        Token value;
        while (iss >> key >> value) specials[key] = value;
    }
    std::vector<std::pair<Token, Token>> m;
    std::vector<std::string> vocab;
    for (int i = 0; i < 256; i++) vocab.push_back(std::string(1, static_cast<char>(i)));
    for (uint32_t i = 0; i < n_merges; i++) {
        Token a = merges[2 * i], b = merges[2 * i + 1];
        if (a >= vocab.size() || b >= vocab.size()) return fail(MBPE_E_INVALID, "merge names an id that does not exist yet");
        m.emplace_back(a, b);
        vocab.push_back(vocab[a] + vocab[b]);
    }
    std::string err;
    int rc = write_model_files(path, pattern, specials, m, write_vocab ? &vocab : nullptr, &err);
    return rc ? fail(rc, err) : MBPE_OK;
}

extern "C" int mbpe_plan_shards(const uint64_t *chunk_off, uint64_t n_chunks, uint32_t n_parts,
                                uint64_t *first_chunk_out) {
    if (!chunk_off || !first_chunk_out || n_parts == 0) return fail(MBPE_E_INVALID, "null argument");
    const uint64_t base = chunk_off[0], total = chunk_off[n_chunks] - base;
    first_chunk_out[0] = 0;
    uint64_t c = 0;
    for (uint32_t p = 1; p < n_parts; p++) {
        const uint64_t target = base + total / n_parts * p; // first chunk that starts at or after the byte target
        uint64_t lo = c, hi = n_chunks;
        while (lo < hi) {
            uint64_t mid = (lo + hi) / 2;
            if (chunk_off[mid] < target)
                lo = mid + 1;
            else
                hi = mid;
        }
        c = lo;
        first_chunk_out[p] = c;
    }
    first_chunk_out[n_parts] = n_chunks;
    return MBPE_OK;
}

extern "C" int mbpe_synth_corpus(uint64_t seed, uint8_t *out, uint64_t n, int n_threads) {
    if (!out && n) return fail(MBPE_E_INVALID, "null argument");
    synth_corpus(seed, out, n, n_threads);
    return MBPE_OK;
}

extern "C" int mbpe_synth_corpus_at(uint64_t seed, uint64_t first_block, uint8_t *out, uint64_t n, int n_threads) {
    if (!out && n) return fail(MBPE_E_INVALID, "null argument");
    synth_corpus_at(seed, first_block, out, n, n_threads);
    return MBPE_OK;
}

extern "C" int mbpe_write_vocab_karpathy(const char *path, const char *special_contents, uint64_t special_len,
                                         const uint32_t *merges, uint32_t n_merges) {
    if (!path || (n_merges && !merges)) return fail(MBPE_E_INVALID, "null argument");
    std::vector<std::pair<std::string, Token>> specials; // file order = registration order
    if (special_contents && special_len) {
        std::istringstream iss(std::string(special_contents, special_len));
        std::string key;
        Token value;
        while (iss >> key >> value) specials.emplace_back(key, value);
    }
    std::vector<std::pair<Token, Token>> m;
    std::vector<std::string> vocab;
    for (int i = 0; i < 256; i++) vocab.push_back(std::string(1, static_cast<char>(i)));
    for (uint32_t i = 0; i < n_merges; i++) {
        Token a = merges[2 * i], b = merges[2 * i + 1];
        if (a >= vocab.size() || b >= vocab.size()) return fail(MBPE_E_INVALID, "merge names an id that does not exist yet");
        m.emplace_back(a, b);
        vocab.push_back(vocab[a] + vocab[b]);
    }
    std::string err;
    int rc = write_vocab_karpathy(path, specials, m, vocab, &err);
    return rc ? fail(rc, err) : MBPE_OK;
}

extern "C" int mbpe_tokenizer_save_vocab_karpathy(mbpe_tokenizer *t, const char *path) {
    if (!t || !path) return fail(MBPE_E_INVALID, "null argument");
    int rc = t->tk.save_vocab_karpathy(path);
    return rc ? fail(rc, t->tk.error()) : MBPE_OK;
}
