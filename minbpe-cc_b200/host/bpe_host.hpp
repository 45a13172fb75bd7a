// bpe_host.hpp -- host (C++23) front end: regex pre-tokenisation, chunk dedup, special tokens, model files.
// The merge loop, the merge scan and the decode gather are GPU calls through include/minbpe_b200.h; pre-tokenisation
// and chunk dedup run here on the host (PCRE2, any pattern) or, for the GPT-4 pattern, on the device (section 6 of the ABI).
#pragma once
#include <cstdint>
#include <string>
#include <string_view>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../../include/minbpe_b200.h"
#include "pcre2_api.h"

namespace mbpe::host {

using Token = uint32_t; // Tokenizer.h:38

extern const char *const kGpt2Pattern; // Tokenizer.h:59
extern const char *const kGpt4Pattern; // Tokenizer.h:60

// One compiled pattern, shareable across threads (PCRE2 codes are read-only at match time; each thread
// brings its own match data -- the reference shares one match_data and is single-threaded, Tokenizer.h:68).
class Regex {
  public:
    Regex() = default;
    ~Regex();
    Regex(const Regex &) = delete;
    Regex &operator=(const Regex &) = delete;
    Regex &operator=(Regex &&o) noexcept;
    // options as the reference picks them (Tokenizer.h:407-415); JIT always (results do not depend on it)
    int compile(const std::string &pattern, std::string *err);
    bool empty() const { return code_ == nullptr; }
    const pcre2_code_8 *code() const { return code_; }

  private:
    pcre2_code_8 *code_ = nullptr;
};

struct Span {
    uint64_t start, end;
};

// Tokenizer.h:506-540: the sequential match loop over text[0, len), restricted to matches that start in
// [begin, stop). The subject is always the whole text, so look-aheads see past `stop`.
int split_range(const Regex &re, const uint8_t *text, uint64_t len, uint64_t begin, uint64_t stop, int fast,
                std::vector<Span> &out, std::string *err);
// Same chunk list as split_range over the whole text, computed on n_threads threads by cutting the text at
// regex-safe points: after '\n' and before a printable non-space ASCII byte. Only used for the
// two built-in patterns; any other pattern runs sequentially.
int split_parallel(const Regex &re, const std::string &pattern, const uint8_t *text, uint64_t len, int n_threads,
                   std::vector<Span> &out, std::string *err);

// 2-bit class per code point (0 other, 1 \p{L}, 2 \p{N}, 3 \s) as the linked PCRE2 sees them, for the GPU matcher of
// the GPT-4 pattern; table = 0x110000 / 4 bytes. Fails if PCRE2's caseless folding differs from the matcher's.
int pretok_class_table(uint8_t *table, std::string *err);

// Tokenizer.h:85-100: a chunk that starts with NUL and parses as an int becomes that single id.
bool marker_token(std::string_view chunk, Token *id);

struct Corpus {
    std::vector<uint32_t> tokens;   // unique chunks, bytes widened to u32, concatenated
    std::vector<uint64_t> off;      // n_unique + 1
    std::vector<uint32_t> weight;   // multiplicity
    uint64_t n_chunks = 0;          // before dedup
};
// unique chunk -> count, unique chunks in first-appearance order
void dedup_chunks(const uint8_t *text, const std::vector<Span> &chunks, int n_threads, Corpus &out);

// split + dedup in one parallel pass (the train path): same Corpus as split_parallel followed by dedup_chunks
int split_dedup_parallel(const Regex &re, const std::string &pattern, const uint8_t *text, uint64_t len, int n_threads,
                         Corpus &out, std::string *err);

// deterministic synthetic Zipfian UTF-8 corpus (SURVEY 8(d) input 3)
void synth_corpus(uint64_t seed, uint8_t *out, uint64_t n, int n_threads);
void synth_corpus_at(uint64_t seed, uint64_t first_block, uint8_t *out, uint64_t n, int n_threads);

int hardware_threads();

// Tokenizer.h:605-650: cut the text at special tokens (earliest occurrence first; equal positions go to the token the
// map iterates first). id < 0: ordinary text [start, end); otherwise the special token with that id.
struct SpecialPart {
    size_t start, end;
    int64_t id;
};
std::vector<SpecialPart> split_special_spans(std::string_view text, const std::unordered_map<std::string, Token> &specials);

// The reference's Tokenizer, method for method (Tokenizer.h:379-927).
class Tokenizer {
  public:
    enum CONFLICT_RESOLUTION { FIRST = MBPE_MODE_FIRST, LEXICAL = MBPE_MODE_LEXICAL }; // Tokenizer.h:54-57

    explicit Tokenizer(const std::string &pattern, int device = 0); // Tokenizer.h:391
    ~Tokenizer();
    Tokenizer(const Tokenizer &) = delete;
    Tokenizer &operator=(const Tokenizer &) = delete;

    void set_special_tokens_from_file(const std::string &contents);                                    // :476
    int train(std::string_view text, int vocab_size, CONFLICT_RESOLUTION mode, bool verbose);          // :489
    int encode(std::string_view text, bool verbose, std::vector<Token> &out);                          // :653
    int encode_into(std::string_view text, Token *out, uint64_t cap, uint64_t *n_out); // same ids, caller's buffer
    int encode_file(const std::string &in_path, const std::string &out_path, uint64_t *n_ids); // streaming, .enc layout
    int decode_file(const std::string &in_path, const std::string &out_path, uint64_t *n_ids, uint64_t *n_bytes);
    int decode(const std::vector<Token> &tokens, bool verbose, std::string &out);                      // :725
    int load(const std::string &path, bool verbose);                                                   // :754
    int save(const std::string &path, bool write_vocab);                                               // :875
    int save_vocab_karpathy(const std::string &path); // .vocab in karpathy/minbpe's layout

    const std::vector<std::pair<Token, Token>> &merges() const { return merges_; }
    const std::string &pattern() const { return pattern_; }
    const std::string &error() const { return error_; }
    void set_engine(int e) { engine_ = e; }
    void set_threads(int n) { n_threads_ = n; }

    // timings / sizes of the last train() for the benchmark
    mbpe_train_stats last_stats{};
    double last_split_s = 0, last_dedup_s = 0;
    uint64_t last_n_chunks = 0, last_n_unique = 0;
    bool last_split_on_gpu = false;

  private:
    std::vector<std::string> split_on_special(std::string_view text); // :605-650
    int ensure_encoder();
    int encode_on_device(std::string_view text, Token *out, uint64_t cap, uint64_t *n_out);
    int encode_host(std::string_view text, bool verbose, std::vector<Token> &out);
    bool use_gpu_split(size_t n_bytes);
    void rebuild_vocab();

    std::string pattern_;
    Regex regex_;
    std::unordered_map<std::string, Token> special_tokens_;          // same container as the reference: the
    std::unordered_map<Token, std::string> special_reverse_;         // .model lists them in ITS iteration order (F7)
    std::vector<std::pair<Token, Token>> merges_;
    std::vector<std::string> vocab_;
    mbpe_encoder *encoder_ = nullptr;
    mbpe_pretok *pretok_ = nullptr; // device matcher of the GPT-4 pattern, created on first use
    bool pretok_failed_ = false;
    std::string pretok_pattern_;     // the pattern the device matcher was last told to match
    bool specials_on_device_ = false; // the encoder knows the special tokens as ready-made chunks
    bool encoder_stale_ = true;
    int device_ = 0, engine_ = MBPE_ENGINE_PERSISTENT, n_threads_ = 0;
    std::string error_;
};

int write_vocab_karpathy(const std::string &path, const std::vector<std::pair<std::string, Token>> &specials,
                         const std::vector<std::pair<Token, Token>> &merges, const std::vector<std::string> &vocab,
                         std::string *err);
// .model / .vocab bytes (Tokenizer.h:878-918)
int write_model_files(const std::string &path, const std::string &pattern,
                      const std::unordered_map<std::string, Token> &specials,
                      const std::vector<std::pair<Token, Token>> &merges, const std::vector<std::string> *vocab,
                      std::string *err);

} // namespace mbpe::host
