/*
 * ra.c -- RawArray container I/O for libtron_b200 (see include/ra.h).
 *
 * Format and behaviour follow /root/reference/src/ra.cu:87-174 and ra.h:38-72;
 * the MATLAB reader/writer (raread.m:44-55, rawrite.m:47-61) documents the same
 * byte layout.  Written as a streaming reader/writer with short-read/short-write
 * loops instead of the reference's fixed 2 GiB chunk arithmetic.
 */
#define _GNU_SOURCE
#include "../../include/ra.h"
#include "../../include/float16.h"

#include <errno.h>
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

/* provided by the CUDA side of the library (plan.cu links cudart) */
extern int tron_pinned_alloc(void **p, size_t bytes);
extern void tron_pinned_free(void *p);

#define RA_IO_CHUNK ((size_t)1 << 30)

static int read_fully(int fd, void *buf, uint64_t n)
{
    uint8_t *p = (uint8_t *)buf;
    while (n > 0) {
        size_t want = n < RA_IO_CHUNK ? (size_t)n : RA_IO_CHUNK;
        ssize_t got = read(fd, p, want);
        if (got < 0) { if (errno == EINTR) continue; return -errno; }
        if (got == 0) return -EIO;                    /* truncated file */
        p += got; n -= (uint64_t)got;
    }
    return 0;
}

static int write_fully(int fd, const void *buf, uint64_t n)
{
    const uint8_t *p = (const uint8_t *)buf;
    while (n > 0) {
        size_t want = n < RA_IO_CHUNK ? (size_t)n : RA_IO_CHUNK;
        ssize_t put = write(fd, p, want);
        if (put < 0) { if (errno == EINTR) continue; return -errno; }
        p += put; n -= (uint64_t)put;
    }
    return 0;
}

uint64_t ra_header_bytes(const ra_t *a) { return 48 + 8 * a->ndims; }

static int read_header_fd(int fd, ra_t *a)
{
    uint64_t h[6];
    int rc = read_fully(fd, h, sizeof h);
    if (rc) return rc;
    if (h[0] != RA_MAGIC_NUMBER) { fprintf(stderr, "Invalid RA file.\n"); return -EINVAL; }
    a->flags = h[1] & ~(uint64_t)RA_FLAG_PINNED_DATA;      /* in-memory only: a file cannot claim pinned storage */
    a->eltype = h[2]; a->elbyte = h[3]; a->size = h[4]; a->ndims = h[5];
    if (a->flags & ~(uint64_t)(RA_FLAG_BIG_ENDIAN | RA_FLAG_COMPRESSED))
        fprintf(stderr, "Warning: RA file carries flags this reader does not know (0x%llx).\n",
                (unsigned long long)a->flags);
    if (a->ndims == 0 || a->ndims > 64) return -EINVAL;
    a->dims = (uint64_t *)malloc(sizeof(uint64_t) * a->ndims);
    if (!a->dims) return -ENOMEM;
    rc = read_fully(fd, a->dims, sizeof(uint64_t) * a->ndims);
    if (!rc && !(a->flags & RA_FLAG_COMPRESSED)) {
        /* size must be prod(dims) * elbyte (ra.cu:141-147 writes it that way); the product is overflow checked */
        uint64_t ne = 1;
        int bad = a->elbyte == 0;
        for (uint64_t i = 0; i < a->ndims && !bad; ++i) {
            if (a->dims[i] != 0 && ne > UINT64_MAX / a->dims[i]) bad = 1; else ne *= a->dims[i];
        }
        if (!bad && ne != 0 && a->elbyte > UINT64_MAX / ne) bad = 1;
        if (bad || ne * a->elbyte != a->size) {
            fprintf(stderr, "Invalid RA file: size field %llu does not match its dims and element size.\n", (unsigned long long)a->size);
            rc = -EINVAL;
        }
    }
    if (rc) { free(a->dims); a->dims = NULL; }
    return rc;
}

static int ra_read_impl(ra_t *a, const char *path, int mode /* 0 malloc, 1 pinned, 2 header only */)
{
    memset(a, 0, sizeof *a);
    int fd = open(path, O_RDONLY);
    if (fd < 0) { int e = errno; fprintf(stderr, "ra_read: cannot open %s: %s\n", path, strerror(e)); return -e; }
    int rc = read_header_fd(fd, a);
    if (rc || mode == 2) { close(fd); return rc; }
    if (mode == 1) {
        void *p = NULL;
        if (tron_pinned_alloc(&p, a->size ? a->size : 1)) rc = -ENOMEM;
        a->data = (uint8_t *)p;
        a->flags |= RA_FLAG_PINNED_DATA;
    } else {
        a->data = (uint8_t *)malloc(a->size ? a->size : 1);
        if (!a->data) rc = -ENOMEM;
    }
    if (!rc) rc = read_fully(fd, a->data, a->size);
    close(fd);
    if (rc) { fprintf(stderr, "ra_read: %s: %s\n", path, strerror(-rc)); ra_free(a); }
    return rc;
}

int ra_read(ra_t *a, const char *path) { return ra_read_impl(a, path, 0); }
int ra_read_pinned(ra_t *a, const char *path) { return ra_read_impl(a, path, 1); }
int ra_read_header(ra_t *a, const char *path) { return ra_read_impl(a, path, 2); }

int ra_write(ra_t *a, const char *path)
{
    int fd = open(path, O_WRONLY | O_TRUNC | O_CREAT, 0644);
    if (fd < 0) { int e = errno; fprintf(stderr, "ra_write: cannot open %s: %s\n", path, strerror(e)); return -e; }
    uint64_t h[6] = { RA_MAGIC_NUMBER, a->flags & ~RA_FLAG_PINNED_DATA, a->eltype, a->elbyte, a->size, a->ndims };
    int rc = write_fully(fd, h, sizeof h);
    if (!rc) rc = write_fully(fd, a->dims, sizeof(uint64_t) * a->ndims);
    if (!rc) rc = write_fully(fd, a->data, a->size);
    if (close(fd) && !rc) rc = -errno;
    if (rc) fprintf(stderr, "ra_write: %s: %s\n", path, strerror(-rc));
    return rc;
}

void ra_free(ra_t *a)
{
    if (!a) return;
    free(a->dims);
    if (a->flags & RA_FLAG_PINNED_DATA) tron_pinned_free(a->data);
    else free(a->data);
    a->dims = NULL; a->data = NULL; a->flags &= ~RA_FLAG_PINNED_DATA;
}

static const char *type_name(uint64_t t)
{
    static const char *names[] = { "user", "int", "uint", "float", "complex" };
    return t < 5 ? names[t] : "unknown";
}

void ra_query(const char *path)
{
    ra_t a;
    if (ra_read_header(&a, path)) return;
    printf("---\nname: %s\nendian: %s\ncompressed: %s\ntype: %s%llu\neltype: %llu\nelbyte: %llu\nsize: %llu\ndimension: %llu\nshape:\n",
           path, (a.flags & RA_FLAG_BIG_ENDIAN) ? "big" : "little", (a.flags & RA_FLAG_COMPRESSED) ? "yes" : "no",
           type_name(a.eltype), (unsigned long long)(a.elbyte * 8), (unsigned long long)a.eltype,
           (unsigned long long)a.elbyte, (unsigned long long)a.size, (unsigned long long)a.ndims);
    for (uint64_t i = 0; i < a.ndims; ++i) printf("  - %llu\n", (unsigned long long)a.dims[i]);
    printf("...\n");
    ra_free(&a);
}

static uint64_t nelem(const ra_t *a)
{
    uint64_t n = 1;
    for (uint64_t i = 0; i < a->ndims; ++i) n *= a->dims[i];
    return n;
}

int ra_reshape(ra_t *r, const uint64_t newdims[], const uint64_t ndimsnew)
{
    uint64_t n = 1;
    for (uint64_t i = 0; i < ndimsnew; ++i) n *= newdims[i];
    if (n != nelem(r)) return -EINVAL;
    uint64_t *d = (uint64_t *)malloc(sizeof(uint64_t) * (ndimsnew ? ndimsnew : 1));
    if (!d) return -ENOMEM;
    memcpy(d, newdims, sizeof(uint64_t) * ndimsnew);
    free(r->dims);
    r->dims = d; r->ndims = ndimsnew;
    return 0;
}

/* drop singleton dimensions (keeps at least one) */
int ra_squash(ra_t *r)
{
    uint64_t j = 0;
    for (uint64_t i = 0; i < r->ndims; ++i)
        if (r->dims[i] != 1) r->dims[j++] = r->dims[i];
    if (j == 0) { r->dims[0] = 1; j = 1; }
    r->ndims = j;
    return 0;
}

/* 0 if identical in type, shape and payload; 1 otherwise */
int ra_diff(const ra_t *a, const ra_t *b)
{
    if ((a->flags & ~RA_FLAG_PINNED_DATA) != (b->flags & ~RA_FLAG_PINNED_DATA)) return 1;
    if (a->eltype != b->eltype || a->elbyte != b->elbyte || a->size != b->size || a->ndims != b->ndims) return 1;
    if (memcmp(a->dims, b->dims, sizeof(uint64_t) * a->ndims)) return 1;
    return memcmp(a->data, b->data, a->size) ? 1 : 0;
}

/* Element-type conversion between the floating layouts the NUFFT path uses:
 * float <-> half (eltype 3, elbyte 4 <-> 2) and complex64 <-> complex-half
 * (eltype 4, elbyte 8 <-> 4).  Other requests leave the array untouched. */
void ra_convert(ra_t *r, const uint64_t eltype, const uint64_t elbyte)
{
    if (eltype != r->eltype || elbyte == r->elbyte) return;
    if (eltype != RA_TYPE_FLOAT && eltype != RA_TYPE_COMPLEX) return;
    uint64_t scalars = nelem(r) * (eltype == RA_TYPE_COMPLEX ? 2 : 1);
    uint64_t from = r->elbyte / (eltype == RA_TYPE_COMPLEX ? 2 : 1);
    uint64_t to = elbyte / (eltype == RA_TYPE_COMPLEX ? 2 : 1);
    if (!((from == 4 && to == 2) || (from == 2 && to == 4))) return;
    uint8_t *nd = (uint8_t *)malloc(scalars * to ? scalars * to : 1);
    if (!nd) return;
    if (to == 2) tron_float_to_half_array((uint16_t *)nd, (const float *)r->data, scalars);
    else tron_half_to_float_array((float *)nd, (const uint16_t *)r->data, scalars);
    if (r->flags & RA_FLAG_PINNED_DATA) { tron_pinned_free(r->data); r->flags &= ~RA_FLAG_PINNED_DATA; }
    else free(r->data);
    r->data = nd; r->elbyte = elbyte; r->size = scalars * to;
}
