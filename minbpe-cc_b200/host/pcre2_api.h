/* pcre2_api.h -- prototypes of the PCRE2 8-bit functions the host front end calls.
 *
 * The image ships the PCRE2 runtime (libpcre2-8.so.0, 10.42, JIT, Unicode 14) without its development header,
 * so the published prototypes and option values are declared here. It is the same library the reference links
 * (Tokenizer.h:20-21, build.zig links pcre2-8), which is what makes regex pre-tokenisation identical.
 * Link with -l:libpcre2-8.so.0.
 */
#ifndef MBPE_PCRE2_API_H
#define MBPE_PCRE2_API_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct pcre2_real_general_context_8 pcre2_general_context_8;
typedef struct pcre2_real_compile_context_8 pcre2_compile_context_8;
typedef struct pcre2_real_match_context_8 pcre2_match_context_8;
typedef struct pcre2_real_code_8 pcre2_code_8;
typedef struct pcre2_real_match_data_8 pcre2_match_data_8;
typedef struct pcre2_real_jit_stack_8 pcre2_jit_stack_8;

enum {
    MBPE_PCRE2_CASELESS = 0x00000008u,
    MBPE_PCRE2_UCP = 0x00020000u,
    MBPE_PCRE2_UTF = 0x00080000u,
    MBPE_PCRE2_NO_UTF_CHECK = 0x40000000u,
    MBPE_PCRE2_JIT_COMPLETE = 0x00000001u,
    MBPE_PCRE2_ERROR_NOMATCH = -1
};

pcre2_code_8 *pcre2_compile_8(const uint8_t *pattern, size_t length, uint32_t options, int *errorcode,
                              size_t *erroroffset, pcre2_compile_context_8 *ccontext);
void pcre2_code_free_8(pcre2_code_8 *code);
int pcre2_jit_compile_8(pcre2_code_8 *code, uint32_t options);
pcre2_match_data_8 *pcre2_match_data_create_from_pattern_8(const pcre2_code_8 *code, pcre2_general_context_8 *gcontext);
void pcre2_match_data_free_8(pcre2_match_data_8 *match_data);
int pcre2_match_8(const pcre2_code_8 *code, const uint8_t *subject, size_t length, size_t startoffset,
                  uint32_t options, pcre2_match_data_8 *match_data, pcre2_match_context_8 *mcontext);
size_t *pcre2_get_ovector_pointer_8(pcre2_match_data_8 *match_data);
int pcre2_get_error_message_8(int errorcode, uint8_t *buffer, size_t bufflen);
#ifdef __cplusplus
}
#endif
#endif
