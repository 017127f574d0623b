/* hfg_inflate.h -- streaming gzip decoder used by the `.cov.gz` reader (hfg_inflate.c).  Internal to libhfg. */
#ifndef HFG_INFLATE_H
#define HFG_INFLATE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct hfg_inflate hfg_inflate;

/* NULL unless `path` can be read and starts with the gzip magic.  piece_bytes: nominal size of the pieces handed out. */
hfg_inflate *hfg_inflate_open(const char *path, size_t piece_bytes);
void hfg_inflate_close(hfg_inflate *z);
/* bytes a destination buffer of hfg_inflate_next must hold */
size_t hfg_inflate_piece_capacity(const hfg_inflate *z);
/* next piece of the uncompressed data -> dst; returns its length, 0 at the end, -1 on corrupt input (hfg_inflate_error).
 * Every member's CRC-32 and length are verified. */
long hfg_inflate_next(hfg_inflate *z, uint8_t *dst);
const char *hfg_inflate_error(const hfg_inflate *z);

#ifdef __cplusplus
}
#endif
#endif
