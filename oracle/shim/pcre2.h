/* oracle/shim/pcre2.h -- TEST INFRASTRUCTURE, not product code.
 *
 * Minimal prototype header for the 8-bit PCRE2 runtime that is installed in this
 * image as /usr/lib/x86_64-linux-gnu/libpcre2-8.so.0 (PCRE2 10.42, JIT, Unicode 14).
 * The image has the shared object but no development header, so the subset of the
 * published PCRE2 API that the reference calls (Tokenizer.h:385-470, :507-540,
 * :677-702, :793-814) is declared here by hand. Constant values are the published
 * upstream ones. Link with  -l:libpcre2-8.so.0 .
 */
#ifndef ORACLE_SHIM_PCRE2_H
#define ORACLE_SHIM_PCRE2_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef uint8_t PCRE2_UCHAR8;
typedef const PCRE2_UCHAR8 *PCRE2_SPTR8;
typedef size_t PCRE2_SIZE;

typedef struct pcre2_real_general_context_8 pcre2_general_context_8;
typedef struct pcre2_real_compile_context_8 pcre2_compile_context_8;
typedef struct pcre2_real_match_context_8 pcre2_match_context_8;
typedef struct pcre2_real_code_8 pcre2_code_8;
typedef struct pcre2_real_match_data_8 pcre2_match_data_8;

#define PCRE2_CASELESS      0x00000008u
#define PCRE2_UCP           0x00020000u
#define PCRE2_UTF           0x00080000u
#define PCRE2_NO_UTF_CHECK  0x40000000u
#define PCRE2_JIT_COMPLETE  0x00000001u
#define PCRE2_ERROR_NOMATCH (-1)

pcre2_general_context_8 *pcre2_general_context_create_8(void *(*)(size_t, void *), void (*)(void *, void *), void *);
void pcre2_general_context_free_8(pcre2_general_context_8 *);
pcre2_compile_context_8 *pcre2_compile_context_create_8(pcre2_general_context_8 *);
void pcre2_compile_context_free_8(pcre2_compile_context_8 *);
pcre2_match_context_8 *pcre2_match_context_create_8(pcre2_general_context_8 *);
void pcre2_match_context_free_8(pcre2_match_context_8 *);

pcre2_code_8 *pcre2_compile_8(PCRE2_SPTR8, PCRE2_SIZE, uint32_t, int *, PCRE2_SIZE *, pcre2_compile_context_8 *);
void pcre2_code_free_8(pcre2_code_8 *);
int pcre2_jit_compile_8(pcre2_code_8 *, uint32_t);

pcre2_match_data_8 *pcre2_match_data_create_from_pattern_8(const pcre2_code_8 *, pcre2_general_context_8 *);
void pcre2_match_data_free_8(pcre2_match_data_8 *);
int pcre2_match_8(const pcre2_code_8 *, PCRE2_SPTR8, PCRE2_SIZE, PCRE2_SIZE, uint32_t, pcre2_match_data_8 *,
                  pcre2_match_context_8 *);
PCRE2_SIZE *pcre2_get_ovector_pointer_8(pcre2_match_data_8 *);
int pcre2_get_error_message_8(int, PCRE2_UCHAR8 *, PCRE2_SIZE);

#ifdef __cplusplus
}
#endif

/* generic (un-suffixed) names for PCRE2_CODE_UNIT_WIDTH == 8 */
#define PCRE2_UCHAR PCRE2_UCHAR8
#define PCRE2_SPTR PCRE2_SPTR8
#define pcre2_match pcre2_match_8
#define pcre2_get_error_message pcre2_get_error_message_8
#define pcre2_get_ovector_pointer pcre2_get_ovector_pointer_8

#endif
