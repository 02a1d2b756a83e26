/* rade_api.h — C ABI of libradae_b200, drop-in for the reference's librade.
 *
 * Every declaration below replaces the same-named symbol of the reference interface
 * /root/reference/src/rade_api.h:71-129 (implementation src/rade_api.c), so that the reference C hosts
 * src/radae_tx.c:14-55 and src/radae_rx.c:14-58 compile and link against this library unchanged.
 * Differences in behaviour are limited to what is behind the boundary: no embedded CPython, all DSP and both
 * core codecs run as sm_100a CUDA kernels (the arithmetic of the reference's *C* codec path, i.e. what the
 * reference does with RADE_USE_C_ENCODER | RADE_USE_C_DECODER), and there is no CPU fallback: rade_open()
 * prints a message and exit(1)s when no CUDA device / no kernel image is available, the same way the
 * reference treats an unusable Python environment (src/rade_api.c:93-102).
 */
#ifndef __RADE_API__
#define __RADE_API__

#include <sys/types.h>

#if IS_BUILDING_RADE_API
#define RADE_EXPORT __attribute__((visibility("default")))
#else
#define RADE_EXPORT
#endif

#ifndef __RADE_COMP__
#define __RADE_COMP__
typedef struct {            /* src/rade_api.h:60-63 */
  float real;
  float imag;
} RADE_COMP;
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define RADE_MODEM_SAMPLE_RATE 8000     /* src/rade_api.h:70 */
#define RADE_SPEECH_SAMPLE_RATE 16000   /* src/rade_api.h:71 */

/* rade_open() flags, src/rade_api.h:74-77.  The codec arithmetic is always the C (int8) path; the two USE_C
 * flags only select the reference behaviours that differ between its two paths: with RADE_USE_C_DECODER the
 * decoder state is NOT reset on (re)sync (src/rade_api.c:494-506), without it it is (radae_rxe.py:263). */
#define RADE_USE_C_ENCODER 0x1
#define RADE_USE_C_DECODER 0x2
#define RADE_FOFF_TEST     0x4          /* +10 Hz frequency error on first sync (src/rade_api.c:263-264) */
#define RADE_VERBOSE_0     0x8

RADE_EXPORT void rade_initialize(void);                                   /* src/rade_api.h:81 */
RADE_EXPORT void rade_finalize(void);                                     /* src/rade_api.h:84 */
/* model_file: unlike the reference (which ignores it, src/rade_api.c:351-352) a readable RDW or DNNw weight
 * file is honoured; anything else falls back to the weights embedded in the library (model19_check3). */
RADE_EXPORT struct rade *rade_open(char model_file[], int flags);         /* src/rade_api.h:87 */
RADE_EXPORT void rade_close(struct rade *r);                              /* src/rade_api.h:88 */
RADE_EXPORT int rade_version(void);                                       /* src/rade_api.h:91 */
RADE_EXPORT int rade_n_tx_out(struct rade *r);                            /* 960,  src/rade_api.h:94 */
RADE_EXPORT int rade_n_tx_eoo_out(struct rade *r);                        /* 1152, src/rade_api.h:95 */
RADE_EXPORT int rade_nin_max(struct rade *r);                             /* 1120, src/rade_api.h:96 */
RADE_EXPORT int rade_n_features_in_out(struct rade *r);                   /* 432,  src/rade_api.h:97 */
RADE_EXPORT int rade_n_eoo_bits(struct rade *r);                          /* 180,  src/rade_api.h:98 */
RADE_EXPORT int rade_tx(struct rade *r, RADE_COMP tx_out[], float features_in[]);            /* src/rade_api.h:102 */
RADE_EXPORT void rade_tx_set_eoo_bits(struct rade *r, float eoo_bits[]);                     /* src/rade_api.h:106 */
RADE_EXPORT int rade_tx_eoo(struct rade *r, RADE_COMP tx_eoo_out[]);                         /* src/rade_api.h:110 */
RADE_EXPORT int rade_nin(struct rade *r);                                                    /* src/rade_api.h:113 */
RADE_EXPORT int rade_rx(struct rade *r, float features_out[], int *has_eoo_out, float eoo_out[], RADE_COMP rx_in[]); /* :119 */
RADE_EXPORT int rade_sync(struct rade *r);                                                   /* src/rade_api.h:122 */
RADE_EXPORT float rade_freq_offset(struct rade *r);                                          /* src/rade_api.h:125 */
RADE_EXPORT int rade_snrdB_3k_est(struct rade *r);                                           /* src/rade_api.h:128 */

#ifdef __cplusplus
}
#endif
#endif
