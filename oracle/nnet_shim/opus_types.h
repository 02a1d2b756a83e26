/* ORACLE / TEST INFRASTRUCTURE ONLY — never linked into the product library.
 *
 * Minimal restatement of the integer typedefs that the reference's generated
 * weight tables need from xiph/opus `include/opus_types.h` (opus is NOT vendored
 * in /root/reference; pinned by cmake/BuildOpus.cmake:10 to commit
 * 940d4e5af64351ca8ba8390df3f555484c567fbb).  Evidence for each typedef:
 *   opus_int8  : `static const opus_int8 enc_gru1_input_weights_int8[12288]`
 *                (src/rade_enc_data.c:9385) — signed 8 bit.
 *   opus_int16 : src/lpcnet_demo.c:167.
 */
#ifndef ORACLE_OPUS_TYPES_H
#define ORACLE_OPUS_TYPES_H
#include <stdint.h>
typedef int8_t   opus_int8;
typedef uint8_t  opus_uint8;
typedef int16_t  opus_int16;
typedef uint16_t opus_uint16;
typedef int32_t  opus_int32;
typedef uint32_t opus_uint32;
#endif
