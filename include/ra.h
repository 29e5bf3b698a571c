/*
 * ra.h -- RawArray (.ra) container, C ABI of libtron_b200.
 *
 * Binary-compatible replacement for the reference interface
 *   /root/reference/src/ra.h:38-48   (ra_t)
 *   /root/reference/src/ra.h:101-111 (ra_read, ra_write, ra_free, ra_query,
 *                                     ra_reshape, ra_convert, ra_squash, ra_diff)
 *   /root/reference/src/ra.cu:87-174 (the three functions the reference defines)
 *
 * On-disk layout (little endian):
 *   u64 magic = 0x7961727261776172   ("rawarray" read as bytes)
 *   u64 flags, u64 eltype, u64 elbyte, u64 size (payload bytes), u64 ndims,
 *   u64 dims[ndims], then the payload, dimension 0 fastest.
 *   A 5-D complex64 file therefore has an 88-byte header.
 *
 * Differences from the reference, all on purpose:
 *   - errors are returned (negative errno-style codes) instead of exit();
 *     ra_read keeps returning 0 on success so existing callers are unaffected;
 *   - ra_write handles payloads whose size is not a multiple of 2 GiB
 *     (the reference never shrinks its chunk size, ra.cu:153-158);
 *   - ra_free releases both dims and data (the reference leaks dims when
 *     built with USE_CUDA, ra.cu:165-174);
 *   - ra_query, ra_reshape, ra_convert, ra_squash and ra_diff, which the
 *     reference header declares but never defines, are implemented;
 *   - half-precision payloads (eltype 3 / elbyte 2 and eltype 4 / elbyte 4) are
 *     understood by ra_convert (the fp16 storage path).
 */
#ifndef TRON_B200_RA_H
#define TRON_B200_RA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    uint64_t flags;     /* RA_FLAG_* bits */
    uint64_t eltype;    /* ra_type */
    uint64_t elbyte;    /* bytes per element (a complex element counts both parts) */
    uint64_t size;      /* payload size in bytes */
    uint64_t ndims;
    uint64_t *dims;     /* malloc'd, ndims entries */
    uint8_t *data;      /* malloc'd (or pinned, see ra_read_pinned), size bytes */
} ra_t;

#define RA_MAGIC_NUMBER      0x7961727261776172ULL
#define RA_FLAG_BIG_ENDIAN   (1ULL << 0)
#define RA_FLAG_COMPRESSED   (1ULL << 1)
#define RA_FLAG_PINNED_DATA  (1ULL << 62)   /* in-memory only: data came from ra_read_pinned */

typedef enum {
    RA_TYPE_USER = 0,
    RA_TYPE_INT,
    RA_TYPE_UINT,
    RA_TYPE_FLOAT,
    RA_TYPE_COMPLEX
} ra_type;

/* the reference's trio */
int  ra_read(ra_t *a, const char *path);
int  ra_write(ra_t *a, const char *path);
void ra_free(ra_t *a);

/* declared by the reference header, defined only here */
void ra_query(const char *path);
int  ra_reshape(ra_t *r, const uint64_t newdims[], const uint64_t ndimsnew);
void ra_convert(ra_t *r, const uint64_t eltype, const uint64_t elbyte);
int  ra_squash(ra_t *r);
int  ra_diff(const ra_t *a, const ra_t *b);

/* additions */
int  ra_read_header(ra_t *a, const char *path);          /* dims only, data = NULL */
int  ra_read_pinned(ra_t *a, const char *path);          /* payload in cudaMallocHost memory */
uint64_t ra_header_bytes(const ra_t *a);                 /* 48 + 8*ndims */

#ifdef __cplusplus
}
#endif
#endif /* TRON_B200_RA_H */
