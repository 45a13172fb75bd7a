/* minbpe_b200.h -- C ABI of the B200-native BPE engine (libminbpe_b200.so).
 *
 * Drop-in boundary for the train and encode hot paths of justinhj/minbpe-cc. The reference is a
 * header-only C++ library with no FFI of its own; the seams this ABI sits behind are:
 *   1. PairCount<Token>          code/include/PairCount.h:27-47   (both conflict-resolution implementations)
 *   2. Tokenizer public methods  code/include/Tokenizer.h:489 train, :653 encode, :725 decode, :754 load, :875 save
 *   3. on-disk formats           .model Tokenizer.h:881-891, .vocab :905-917, .enc examples/minbpe-cc.cpp:58-89
 *
 * Conventions: plain pointers and sizes, caller-owned buffers, int status (0 = ok, negative = error, see
 * MBPE_E_*), nothing throws across the boundary, no torch types. Every compute entry point runs on a B200
 * through hand-written sm_100a kernels; there is NO CPU fallback: without a CUDA device the compute calls
 * return MBPE_E_NO_DEVICE.
 *
 * "stream" arguments are a cudaStream_t passed as void* (NULL = the legacy default stream), so a caller
 * can time the kernels with its own events on its own stream.
 */
#ifndef MINBPE_B200_H
#define MINBPE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MBPE_OK 0
#define MBPE_E_INVALID (-1)    /* bad argument (null pointer, vocab_size < 256, token >= 256 inside a chunk, ...) */
#define MBPE_E_NO_DEVICE (-2)  /* no usable CUDA device / extension built without kernels */
#define MBPE_E_CUDA (-3)       /* a CUDA call failed; mbpe_last_error() has the text */
#define MBPE_E_CAPACITY (-4)   /* caller buffer too small; required size is reported through the out parameter */
#define MBPE_E_IO (-5)         /* file could not be opened / parsed */
#define MBPE_E_REGEX (-6)      /* PCRE2 compile or match error */
#define MBPE_E_UNSUPPORTED (-8) /* the GPU pre-tokeniser declines this text (malformed UTF-8, pathological runs): use the PCRE2 path */
#define MBPE_E_EMPTY (-7)      /* nothing to do where the reference would assert/abort */

/* Tokenizer::CONFLICT_RESOLUTION (Tokenizer.h:54-57) */
#define MBPE_MODE_FIRST 0
#define MBPE_MODE_LEXICAL 1

/* how the merge loop is driven on the device (results are identical) */
#define MBPE_ENGINE_STEPWISE 0   /* every phase of every merge is its own full-grid kernel launch */
#define MBPE_ENGINE_PERSISTENT 1 /* one resident CTA runs small merges back to back; big merges, table growth and
                                    candidate rebuilds go to full-grid kernels */

const char *mbpe_version(void);
const char *mbpe_last_error(void); /* thread-local text of the last error */
int mbpe_device_count(void);       /* number of CUDA devices visible (0 on a CPU-only box) */

/* ------------------------------------------------------------------------------------------------------------
 * 1. PairCount seam  (PairCount.h:27-47; PairCountInsertOrder :101-181, PairCountLexicalOrder :227-279)
 *    Device-resident pair table + arg-max with both tie-breaks. Batched: one call = n create_or_modify_pair
 *    calls applied in array order (insertion order = array order, PairCount.h:149).
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct mbpe_paircount mbpe_paircount;
int mbpe_paircount_create(int mode, int device, mbpe_paircount **out);
void mbpe_paircount_destroy(mbpe_paircount *pc);
/* create_or_modify_pair x n (PairCount.h:141 / :249) */
int mbpe_paircount_add(mbpe_paircount *pc, const uint32_t *a, const uint32_t *b, const int32_t *delta, uint64_t n);
/* get_top_pair_count (PairCount.h:159 / :262): *found = 0 when the table is empty */
int mbpe_paircount_top(mbpe_paircount *pc, uint32_t *a, uint32_t *b, int32_t *count, int *found);
/* get_pair (PairCount.h:123 / :239) */
int mbpe_paircount_get(mbpe_paircount *pc, uint32_t a, uint32_t b, int32_t *count, int *found);
/* get_count (PairCount.h:114 / :235) */
int mbpe_paircount_size(mbpe_paircount *pc, uint64_t *n_pairs);

/* ------------------------------------------------------------------------------------------------------------
 * 2. Train merge loop  (Tokenizer.h:551-589: create_lists, calculate_freqs, loop of get_top_pair_count +
 *    merge_chunks [+ recount in FIRST mode])
 *
 *    Input is the chunk list after pre-tokenisation: tokens = every chunk's bytes widened to u32 and
 *    concatenated, chunk_off = n_chunks+1 offsets into tokens, chunk_weight = multiplicity of each chunk
 *    (NULL = all 1). Chunks may be deduplicated provided unique chunks are in first-appearance order
 *    (exact in both modes). Tokens inside chunks longer than one token must be < 256.
 *
 *    Output: merges_out[2*i], merges_out[2*i+1] = pair merged into id 256+i; counts_out (optional) = that
 *    pair's count when chosen (the "had C occurrences" figure, Tokenizer.h:576); *n_merges_out <=
 *    vocab_size-256 (FIRST stops early when no pair is left, Tokenizer.h:586-588; LEXICAL repeats the
 *    smallest zero-count pair).
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct mbpe_trainer mbpe_trainer;

typedef struct mbpe_train_stats {
    double gpu_ms;           /* CUDA-event time of the whole merge loop incl. index build */
    double build_ms;         /* of which: initial histogram + occurrence index build */
    uint64_t n_positions;    /* tokens uploaded */
    uint64_t n_pairs;        /* distinct pairs ever inserted */
    uint64_t table_slots;    /* final pair-table capacity */
    uint64_t n_launches;     /* kernels launched by the run */
    uint64_t n_big_merges;   /* merges done by full-grid kernels (persistent engine) */
    uint64_t n_rebuilds;     /* candidate-list rebuilds */
    uint64_t n_grows;        /* pair-table rehashes */
    uint64_t rescan_bytes;   /* sum over merges of 12*T_m + 16*P_m: SURVEY 8(d) full-rescan algorithmic volume */
    uint64_t resident_cycles[8]; /* SM cycles inside the resident CTA: select, hits, mutate+seg_alloc, seg_fill, fin,
                                    steps done there, total */
} mbpe_train_stats;

/* uploads the corpus to `device` (H2D) and keeps a pristine copy so run() can be repeated */
int mbpe_trainer_create(const uint32_t *tokens, uint64_t n_tokens, const uint64_t *chunk_off, uint64_t n_chunks,
                        const uint32_t *chunk_weight, int device, mbpe_trainer **out);
/* runs the whole merge loop from the resident pristine copy; only the merge list crosses back (D2H) */
int mbpe_trainer_run(mbpe_trainer *t, uint32_t vocab_size, int mode, int engine, void *stream, uint32_t *merges_out,
                     int32_t *counts_out, uint32_t *n_merges_out, mbpe_train_stats *stats /* optional */);
void mbpe_trainer_destroy(mbpe_trainer *t);

/* one-shot with host buffers: create + run + destroy on device 0 */
int mbpe_train(const uint32_t *tokens, uint64_t n_tokens, const uint64_t *chunk_off, uint64_t n_chunks,
               const uint32_t *chunk_weight, uint32_t vocab_size, int mode, uint32_t *merges_out, int32_t *counts_out,
               uint32_t *n_merges_out);

/* Sharded training, one process per GPU. Rank r owns a contiguous, token-balanced share of the unique
 * chunks and a full replica of the pair table with GLOBAL counts; every merge step the ranks exchange their count
 * deltas over NVLink, so every rank picks the same pair with no broadcast and ends with the same merge list.
 * MBPE_ENGINE_PERSISTENT: one resident CTA per rank runs the merges back to back and does the exchange itself -- it
 * stores its records straight into the peers' memory (mapped with cudaIpc by mbpe_comm_create) and polls the flag words
 * the peers store into its own; only rebuilds and the few early merges with very large counts are driven from the host.
 * MBPE_ENGINE_STEPWISE (or no peer mapping: mbpe_comm_resident() == 0): grid kernels + one NCCL all-gather per merge.
 * Every rank passes the SAME deduplicated corpus. The 128-byte id comes from rank 0 (mbpe_comm_unique_id) and
 * reaches the other ranks by any means (torch.distributed broadcast in bench.py). NCCL is loaded at run time. */
typedef struct mbpe_comm mbpe_comm;
int mbpe_comm_unique_id(uint8_t *id_out /* 128 bytes */);
int mbpe_comm_create(const uint8_t *id /* 128 bytes */, int rank, int world, int device, mbpe_comm **out);
int mbpe_comm_resident(const mbpe_comm *c); /* 1: every rank mapped every peer: the resident exchange is available */
void mbpe_comm_destroy(mbpe_comm *c);
/* this rank's share of the corpus uploaded once; run() can be repeated (what bench.py times at N > 1) */
typedef struct mbpe_sharded_trainer mbpe_sharded_trainer;
int mbpe_sharded_trainer_create(mbpe_comm *comm, const uint32_t *tokens, uint64_t n_tokens, const uint64_t *chunk_off,
                                uint64_t n_chunks, const uint32_t *chunk_weight, mbpe_sharded_trainer **out);
int mbpe_sharded_trainer_run(mbpe_sharded_trainer *t, uint32_t vocab_size, int mode, int engine, void *stream,
                             uint32_t *merges_out, int32_t *counts_out, uint32_t *n_merges_out, mbpe_train_stats *stats);
void mbpe_sharded_trainer_destroy(mbpe_sharded_trainer *t);
/* Tokenizer::train over a text that is sharded over the ranks ("shard what shards naturally"): rank r passes
 * ITS contiguous part of the text (parts in rank order, cut at chunk boundaries); every rank splits and deduplicates its
 * part on its GPU, one all-gather moves the unique chunks (bytes + offsets + weights), every rank merges them into the
 * same corpus (first-appearance order of the whole text) and runs the merge loop on it -- identical merge list on every
 * rank, no per-merge communication. GPT-2 / GPT-4 patterns (the device front end); MBPE_E_UNSUPPORTED as mbpe_pretok_corpus. */
typedef struct mbpe_pretok mbpe_pretok;
int mbpe_train_text_sharded(mbpe_comm *comm, mbpe_pretok *p, const uint8_t *text_part, uint64_t len, uint32_t vocab_size, int mode,
                            uint32_t *merges_out, int32_t *counts_out, uint32_t *n_merges_out, mbpe_train_stats *stats,
                            double *front_end_s /* optional: split + dedup + gather + merge */);
/* one-shot with host buffers: create + run (resident engine when available) + destroy */
int mbpe_train_sharded(mbpe_comm *comm, const uint32_t *tokens, uint64_t n_tokens, const uint64_t *chunk_off,
                       uint64_t n_chunks, const uint32_t *chunk_weight, uint32_t vocab_size, int mode, void *stream,
                       uint32_t *merges_out, int32_t *counts_out, uint32_t *n_merges_out, mbpe_train_stats *stats);

/* ------------------------------------------------------------------------------------------------------------
 * 3. Encode merge scan + decode  (Tokenizer.h:325-377 internal_internal_encode / internal_encode, :714-717
 *    flatten; :725-751 decode; lookup table built as load() does, :833-837, later duplicate pairs overwrite)
 *
 *    Encode semantics are the reference's: per chunk, scan left to right replacing ANY known pair, repeat
 *    until a pass merges nothing -- not rank-ordered BPE.
 *    Chunks are byte ranges of `bytes`: chunk c = [chunk_off[c], chunk_off[c+1]).
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct mbpe_encoder mbpe_encoder;
int mbpe_encoder_create(const uint32_t *merges, uint32_t n_merges, int device, mbpe_encoder **out);
void mbpe_encoder_destroy(mbpe_encoder *e);
/* special tokens for decode (Tokenizer.h:733-736): ids + concatenated byte strings, off has n+1 entries */
int mbpe_encoder_set_specials(mbpe_encoder *e, const uint32_t *ids, const uint8_t *bytes, const uint64_t *off,
                              uint32_t n);

/* host buffers in, host buffers out (H2D + kernels + D2H). out_tokens must hold n_bytes entries in the worst
 * case; *n_out receives the count. out_off (optional) receives n_chunks+1 token offsets. */
int mbpe_encode(mbpe_encoder *e, const uint8_t *bytes, uint64_t n_bytes, const uint64_t *chunk_off,
                uint64_t n_chunks, uint32_t *out_tokens, uint64_t out_cap, uint64_t *n_out, uint64_t *out_off);
/* everything already resident on the encoder's device (d_* are device pointers); *d_n_out is a device u64.
 * d_chunk_off32 holds n_chunks+1 u32 offsets (a batch is < 4 GiB). */
int mbpe_encode_device(mbpe_encoder *e, const uint8_t *d_bytes, uint64_t n_bytes, const uint32_t *d_chunk_off32,
                       uint64_t n_chunks, uint32_t *d_out_tokens, uint64_t out_cap, uint64_t *d_n_out,
                       void *stream);
uint64_t mbpe_encoder_launches(const mbpe_encoder *e); /* kernels launched through this handle so far */
/* bytes of device scratch mbpe_encode_device needs for a batch of that shape (kept inside the handle) */
int mbpe_encode_reserve(mbpe_encoder *e, uint64_t n_bytes, uint64_t n_chunks);

/* ids -> bytes. Call with out == NULL to size. Invalid ids are skipped (Tokenizer.h:739-742). */
/* special tokens on the encode side: "a chunk with exactly these bytes is this one id" (Tokenizer.h:667-671), kept in
 * the chunk cache (which is emptied first). ids[i] <-> bytes[off[i] .. off[i+1]). MBPE_E_UNSUPPORTED: cache disabled or
 * a token longer than 31 bytes -- then special tokens have to be resolved on the host as before. */
int mbpe_encoder_seed_special_chunks(mbpe_encoder *e, const uint32_t *ids, const uint8_t *bytes, const uint64_t *off,
                                     uint32_t n);
int mbpe_decode(mbpe_encoder *e, const uint32_t *ids, uint64_t n_ids, uint8_t *out, uint64_t out_cap,
                uint64_t *n_out);
/* resident ids (16-byte aligned) -> resident bytes, one pass, on `stream`. *d_n_out (device) = decoded size; bytes past
 * out_cap are dropped. d_out == NULL: size only. */
int mbpe_decode_device(mbpe_encoder *e, const uint32_t *d_ids, uint64_t n_ids, uint8_t *d_out, uint64_t out_cap,
                       uint64_t *d_n_out, void *stream);
/* .enc file (raw little-endian u32 ids; a trailing partial word is dropped, examples/minbpe-cc.cpp:79) -> text file,
 * in blocks: reader thread, device gather, writer thread; memory use independent of the file size */
int mbpe_decode_file(mbpe_encoder *e, const char *in_path, const char *out_path, uint64_t *n_ids, uint64_t *n_bytes);

/* ------------------------------------------------------------------------------------------------------------
 * 4. Tokenizer mirror (host C++23 front end over 2. and 3.): same method set, argument meaning and error
 *    behaviour as MinBpeCC::Tokenizer::Tokenizer. Regex pre-tokenisation (PCRE2), chunk dedup, special-token
 *    splitting and file formats run on the host; the merge loop / merge scan / gather run on the GPU.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct mbpe_tokenizer mbpe_tokenizer;
const char *mbpe_gpt2_split_pattern(void); /* Tokenizer.h:59 */
const char *mbpe_gpt4_split_pattern(void); /* Tokenizer.h:60 */

/* Tokenizer(pattern) (Tokenizer.h:391); pattern "" = no regex (encoder "basic") */
int mbpe_tokenizer_create(const char *pattern, int device, mbpe_tokenizer **out);
void mbpe_tokenizer_destroy(mbpe_tokenizer *t);
/* set_special_tokens_from_file (Tokenizer.h:476): the file CONTENTS, "token id" per line */
int mbpe_tokenizer_set_special_tokens(mbpe_tokenizer *t, const char *contents, uint64_t len);
/* train (Tokenizer.h:489) */
int mbpe_tokenizer_train(mbpe_tokenizer *t, const uint8_t *text, uint64_t len, int vocab_size, int mode, int verbose);
/* save (Tokenizer.h:875): writes path and, if write_vocab, path + ".vocab" */
int mbpe_tokenizer_save(mbpe_tokenizer *t, const char *path, int write_vocab);
/* load (Tokenizer.h:754) */
int mbpe_tokenizer_load(mbpe_tokenizer *t, const char *path, int verbose);
/* encode (Tokenizer.h:653). Call with out == NULL to get the count. */
int mbpe_tokenizer_encode(mbpe_tokenizer *t, const uint8_t *text, uint64_t len, uint32_t *out, uint64_t out_cap,
                          uint64_t *n_out);
/* decode (Tokenizer.h:725). Call with out == NULL to get the size. */
int mbpe_tokenizer_decode(mbpe_tokenizer *t, const uint32_t *ids, uint64_t n, uint8_t *out, uint64_t out_cap,
                          uint64_t *n_out);
/* introspection used by the parity tests */
int mbpe_tokenizer_get_merges(mbpe_tokenizer *t, uint32_t *merges_out, uint32_t cap_pairs, uint32_t *n_merges);
int mbpe_tokenizer_last_train_stats(mbpe_tokenizer *t, mbpe_train_stats *stats, double *split_s, double *dedup_s,
                                    uint64_t *n_chunks, uint64_t *n_unique);
/* streaming --decode: .enc file -> text file in blocks (any model) */
int mbpe_tokenizer_decode_file(mbpe_tokenizer *t, const char *in_path, const char *out_path, uint64_t *n_ids,
                               uint64_t *n_bytes);
/* streaming --encode: MBPE_E_UNSUPPORTED unless GPT-4 pattern, no special tokens, well-formed UTF-8 */
int mbpe_tokenizer_encode_file(mbpe_tokenizer *t, const char *in_path, const char *out_path, uint64_t *n_ids);
int mbpe_tokenizer_last_split_on_gpu(mbpe_tokenizer *t); /* 1: the last train() pre-tokenised on the device (section 6) */
void mbpe_tokenizer_set_engine(mbpe_tokenizer *t, int engine);
void mbpe_tokenizer_set_threads(mbpe_tokenizer *t, int n_threads); /* host pre-tokenisation threads, 0 = all */

/* ------------------------------------------------------------------------------------------------------------
 * 5. Host-only pieces exposed for CPU-side tests (no GPU needed)
 * ---------------------------------------------------------------------------------------------------------- */
/* regex pre-tokenisation (Tokenizer.h:500-544), multi-threaded with regex-safe cut points; writes chunk
 * [start,end) pairs. Call with starts == NULL to count. */
int mbpe_split(const char *pattern, const uint8_t *text, uint64_t len, int n_threads, uint64_t *starts,
               uint64_t *ends, uint64_t cap, uint64_t *n_chunks);
/* special-token splitter (Tokenizer.h:605-650) in one sweep per token: parts [starts[i], ends[i]) in text order,
 * ids[i] < 0 for ordinary text, else the special token's id. special_contents = "token id" lines (:482-485).
 * Call with starts == NULL to count. */
int mbpe_special_split(const char *special_contents, uint64_t special_len, const uint8_t *text, uint64_t len,
                       uint64_t *starts, uint64_t *ends, int64_t *ids, uint64_t cap, uint64_t *n_parts);
/* 2-bit class per code point (0 other, 1 \p{L}, 2 \p{N}, 3 \s), 0x110000/4 bytes, read out of the linked PCRE2 with
 * the reference's compile options (Tokenizer.h:407): what the GPU matcher of the GPT-4 pattern classifies with.
 * MBPE_E_REGEX if that PCRE2's caseless folding is not the one the matcher assumes. */
#define MBPE_PRETOK_TABLE_BYTES (0x110000 / 4)
int mbpe_pretok_class_table(uint8_t *table_out);
/* chunk dedup in first-appearance order + byte->token widening (Tokenizer.h:85-100). Sizing: n_unique <=
 * n_chunks, n_tokens <= sum of lengths. */
int mbpe_dedup(const uint8_t *text, const uint64_t *starts, const uint64_t *ends, uint64_t n_chunks,
               uint32_t *tokens_out, uint64_t *n_tokens, uint64_t *off_out, uint32_t *weight_out, uint64_t *n_unique);
/* the train front end in one call: regex split + dedup fused and multi-threaded (the chunk list is never
 * materialised). Call with tokens_out == NULL to size (the result is kept for the following call). */
int mbpe_split_dedup(const char *pattern, const uint8_t *text, uint64_t len, int n_threads, uint32_t *tokens_out,
                     uint64_t tokens_cap, uint64_t *n_tokens, uint64_t *off_out, uint32_t *weight_out,
                     uint64_t unique_cap, uint64_t *n_unique, uint64_t *n_chunks);
/* .model / .vocab writer given a merge list (Tokenizer.h:875-926) */
int mbpe_write_model(const char *path, const char *pattern, const char *special_contents, uint64_t special_len,
                     const uint32_t *merges, uint32_t n_merges, int write_vocab);
/* .vocab in karpathy/minbpe's layout (base.py save / render_token) instead of the reference's */
int mbpe_write_vocab_karpathy(const char *path, const char *special_contents, uint64_t special_len, const uint32_t *merges,
                              uint32_t n_merges);
int mbpe_tokenizer_save_vocab_karpathy(mbpe_tokenizer *t, const char *path);
/* multi-GPU encode: split n_chunks chunks into n_parts contiguous ranges of (nearly) equal BYTES; the ranges
 * are whole chunks and in order, so the concatenation of the parts' id streams is the id stream of the whole
 * (no communication). first_chunk_out gets n_parts+1 chunk indices. */
int mbpe_plan_shards(const uint64_t *chunk_off, uint64_t n_chunks, uint32_t n_parts, uint64_t *first_chunk_out);
/* deterministic synthetic Zipfian UTF-8 corpus (SURVEY 8(d) input 3): fills out[0..n) */
int mbpe_synth_corpus(uint64_t seed, uint8_t *out, uint64_t n, int n_threads);
/* the same corpus from 1 MiB block `first_block` on: out[0..n) = bytes [first_block MiB, first_block MiB + n) of it (a rank
 * of a sharded run generates only its own part) */
int mbpe_synth_corpus_at(uint64_t seed, uint64_t first_block, uint8_t *out, uint64_t n, int n_threads);

/* ------------------------------------------------------------------------------------------------------------
 * 6. GPU pre-tokeniser and chunk dedup for the GPT-4 split pattern  (regex loop Tokenizer.h:500-544
 *    with the pattern of :60, chunk -> count SURVEY F2)
 *    Same chunk list as the reference's pcre2_match loop, computed on the device by a hand-written matcher whose
 *    code point classes are read out of the linked PCRE2. Text it cannot take (malformed UTF-8, pathological runs)
 *    is refused with MBPE_E_UNSUPPORTED; the caller then uses mbpe_split (PCRE2 itself).
 * ---------------------------------------------------------------------------------------------------------- */
int mbpe_pretok_create(int device, mbpe_pretok **out); /* matches the GPT-4 pattern until told otherwise */
/* pattern = mbpe_gpt4_split_pattern() or mbpe_gpt2_split_pattern(); MBPE_E_UNSUPPORTED for anything else */
int mbpe_pretok_select(mbpe_pretok *p, const char *pattern);
void mbpe_pretok_destroy(mbpe_pretok *p);
/* resident text (< 4 GiB) -> chunk offsets: d_off_out[0 .. n_chunks] (u32, last = len), the layout
 * mbpe_encode_device takes. off_cap counts u32 entries; len + 2 always suffices. */
int mbpe_pretok_split_device(mbpe_pretok *p, const uint8_t *d_text, uint64_t len, uint32_t *d_off_out, uint64_t off_cap,
                             uint64_t *n_chunks, void *stream);
/* the same for a text with special tokens (Tokenizer.h:605-650): d_sp_begin / d_sp_end = the occurrences in text
 * order (device arrays of n_sp offsets into d_text); every ordinary part in between is split as a subject of its own,
 * every occurrence is one chunk */
int mbpe_pretok_split_device_parts(mbpe_pretok *p, const uint8_t *d_text, uint64_t len, const uint32_t *d_sp_begin,
                                   const uint32_t *d_sp_end, uint32_t n_sp, uint32_t *d_off_out, uint64_t off_cap,
                                   uint64_t *n_chunks, void *stream);
/* host text -> host offsets off_out[0 .. n_chunks] (chunk c = [off[c], off[c+1])): GPU twin of mbpe_split */
int mbpe_pretok_split(mbpe_pretok *p, const uint8_t *text, uint64_t len, uint64_t *off_out, uint64_t off_cap,
                      uint64_t *n_chunks);
/* unique chunks resident on the device, in first-appearance order, in the trainer's input layout */
typedef struct mbpe_device_corpus {
    uint32_t *d_tokens;  /* bytes of the unique chunks widened to u32, concatenated */
    uint64_t *d_off;     /* n_unique + 1 */
    uint32_t *d_weight;  /* multiplicity of each unique chunk */
    uint64_t n_tokens, n_unique, n_chunks;
    int device;
} mbpe_device_corpus;
int mbpe_pretok_dedup_device(mbpe_pretok *p, const uint8_t *d_text, uint64_t len, const uint32_t *d_off,
                             uint64_t n_chunks, mbpe_device_corpus *out, void *stream);
/* the same over several resident text segments (each < 4 GiB; a corpus larger than one segment): chunk order =
 * segment order, so first-appearance order is that of the whole text */
int mbpe_pretok_dedup_segments(mbpe_pretok *p, const uint8_t *const *d_texts, const uint64_t *seg_bytes,
                               const uint32_t *const *d_offs, const uint64_t *seg_chunks, uint32_t n_segs,
                               mbpe_device_corpus *out, void *stream);
/* several ALREADY deduplicated chunk lists (bytes + u32 offsets + weights per list, resident; lists in text order) into
 * one device corpus: the second stage of a sharded train front end (every rank deduplicates its part of the text) */
int mbpe_pretok_merge_corpora(mbpe_pretok *p, const uint8_t *const *d_texts, const uint64_t *seg_bytes,
                              const uint32_t *const *d_offs, const uint32_t *const *d_weights, const uint64_t *seg_chunks,
                              uint32_t n_segs, mbpe_device_corpus *out, void *stream);
/* host text -> device corpus (H2D + split + dedup): the front end of Tokenizer::train (:500-556) */
int mbpe_pretok_corpus(mbpe_pretok *p, const uint8_t *text, uint64_t len, mbpe_device_corpus *out);
int mbpe_device_corpus_download(const mbpe_device_corpus *c, uint32_t *tokens, uint64_t *off, uint32_t *weight);
void mbpe_device_corpus_free(mbpe_device_corpus *c);
/* host text -> host ids: GPT-4 split + merge scan on the device, segment by segment (Tokenizer::encode for a text
 * without special tokens, :653-717). out_cap counts ids; len always suffices. */
int mbpe_encode_text(mbpe_encoder *enc, mbpe_pretok *p, const uint8_t *text, uint64_t len, uint32_t *out,
                     uint64_t out_cap, uint64_t *n_out);
/* ... with special tokens: sp_begin / sp_end = their occurrences in text order (host arrays, as mbpe_special_split
 * reports them); needs mbpe_encoder_seed_special_chunks on the encoder */
int mbpe_encode_text_special(mbpe_encoder *enc, mbpe_pretok *p, const uint8_t *text, uint64_t len, const uint64_t *sp_begin,
                             const uint64_t *sp_end, uint64_t n_sp, uint32_t *out, uint64_t out_cap, uint64_t *n_out);
/* file -> .enc file (raw little-endian u32 ids, examples/minbpe-cc.cpp:58-69) in blocks: reader thread, device
 * pipeline, writer thread; memory use independent of the file size. MBPE_E_UNSUPPORTED if no block
 * boundary can be found (then read the file and call mbpe_encode_text / mbpe_encode). */
int mbpe_encode_file(mbpe_encoder *enc, mbpe_pretok *p, const char *in_path, const char *out_path, uint64_t *n_bytes,
                     uint64_t *n_ids);
/* a trainer over a device corpus: takes the corpus buffers over (the struct is cleared; freeing it afterwards is a
 * no-op), nothing is copied */
int mbpe_trainer_create_device(mbpe_device_corpus *c, mbpe_trainer **out);

#ifdef __cplusplus
}
#endif
#endif
