/* ORACLE / TEST INFRASTRUCTURE ONLY.
 * Restatement of the two element-count memory helpers the reference C sources
 * use from opus `celt/os_support.h` (src/rade_enc.c:49 OPUS_CLEAR,
 * src/rade_enc.c:73 OPUS_COPY).  Semantics: n ELEMENTS, not bytes. */
#ifndef ORACLE_OS_SUPPORT_H
#define ORACLE_OS_SUPPORT_H
#include <string.h>
#include <stdlib.h>
#define OPUS_COPY(dst, src, n)  (memcpy((dst), (src), (n)*sizeof(*(dst))))
#define OPUS_MOVE(dst, src, n)  (memmove((dst), (src), (n)*sizeof(*(dst))))
#define OPUS_CLEAR(dst, n)      (memset((dst), 0, (n)*sizeof(*(dst))))
#endif
