/*
 * tron_oracle.h -- CPU restatement of the TRON radial NUFFT hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, load or call it, and there only as the
 * checker or the reported CPU baseline.  The product (tron_b200/) never links
 * or imports this file.
 *
 * Parity status: the reference (davidssmith/TRON, src/tron.cu) has no CPU path
 * and ships no golden vectors for the NUFFT path.  This
 * restatement is pinned against outputs of the reference itself, produced by
 * oracle/_ref (the unmodified reference sources compiled in place) on a B200
 * and committed under tests/golden/ together with the generating script.
 * float16 and RA are pinned against oracle/_ref's host objects directly.
 *
 * Every function cites the reference file:line it follows.  Arithmetic that the
 * reference runs through GPU special-function units (sin.approx, sqrt.approx,
 * rcp.approx) is restated with libm; results agree with the reference to
 * ~1e-6 relative L2, not bit-for-bit.
 */
#ifndef TRON_ORACLE_H
#define TRON_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float x, y; } ocplx;        /* == CUDA float2 */

/* geometry + flags, mirrors the file-static configuration of tron.cu:54-87 */
typedef struct {
    /* inputs: RA header dims and CLI flags */
    uint64_t dims[5];        /* ra_in.dims */
    int adjoint;             /* -a */
    int golden_angle;        /* -G */
    float gridos;            /* -o, default 2 */
    float kernwidth;         /* -k, default 2 */
    float data_undersamp;    /* -u, default 1 */
    int prof_slide;          /* -d, default 0 -> npe1work */
    int skip_angles;         /* -s, default 0 */
    /* derived (oracle_geometry fills these; tron.cu:905-961) */
    int nc, nt, nro, npe1, npe2, npe1work;
    int nx, ny, nz, nxos, nyos;
    uint64_t out_dims[5];
    uint64_t out_elems;      /* number of complex elements of the output */
    /* extensions, 0 = the reference's executed path */
    int niter;               /* -i: CGNR iterations (tron.cu:665-720, restated so that it converges) */
    int coil_combine;        /* 0 root sum of squares (tron.cu:764), 1 Walsh adaptive combine (tron.cu:270-302, 766) */
    int walsh_npatch;        /* patch half-width of the Walsh combine (tron.cu:766 passes 1) */
} oracle_cfg;

void oracle_cfg_defaults(oracle_cfg *c);
int  oracle_geometry(oracle_cfg *c);

/* scalar kernels (tron.cu:304-378) */
float oracle_besseli0(float x);
float oracle_gridkernel(float x, float kernwidth);
float oracle_gridkernelhat(float u, float kernwidth);
float oracle_modang(float x);
float oracle_spoke_angle_grid(int pe, int npe, int skip_angles, int golden);
float oracle_spoke_angle_degrid(int pe, int npe, int skip_angles, int golden);

/* install (or clear with NULL) SFU sin/cos values of the spokes, see tron_oracle.c */
void oracle_set_trig_table(const float *ct, const float *st, int n);

/* array kernels */
void oracle_precompensate(ocplx *nudata, int nchan, int nro, int npe1work);
void oracle_gridradial2d(ocplx *udata, const ocplx *nudata, int nxos, int nchan,
                         int nro, int npe, float kernwidth, int skip_angles, int golden);
void oracle_degridradial2d(ocplx *nudata, const ocplx *udata, int n, int nrep,
                           int nro, int npe, float W, int skip_angles, int golden);
void oracle_fftshift(ocplx *dst, const ocplx *src, int n, int nchan, int inverse_dir);
void oracle_fft2(ocplx *data, int n, int nchan, int sign);
void oracle_crop(ocplx *dst, int ndst, const ocplx *src, int nsrc, int nchan);
void oracle_pad(ocplx *dst, int ndst, const ocplx *src, int nsrc, int nchan);
void oracle_deapod(ocplx *a, int n, int nrep, float m, float sigma);
void oracle_coilcombinesos(ocplx *img, const ocplx *coilimg, int nimg, int nchan);
void oracle_coilcombinewalsh(ocplx *img, const ocplx *coilimg, int nimg, int nchan, int npatch);

/* index-map dumps (bit-level comparison targets, see tests) */
long oracle_grid_hits(int32_t *hits, long maxhits, int nxos, int nro, int npe,
                      float kernwidth, int skip_angles, int golden);

/* one-slice pipelines (tron.cu:623-649 + 764) and the slice loop (tron.cu:726-786) */
void oracle_nufft_adj_slice(const oracle_cfg *c, ocplx *img_out, const ocplx *samples, int peoffset);
void oracle_nufft_adj_coils(const oracle_cfg *c, ocplx *coilimg_out, const ocplx *samples, int peoffset);
void oracle_cgnr_coils(const oracle_cfg *c, ocplx *x, const ocplx *samples, int peoffset, int niter);
void oracle_nufft_fwd_slice(const oracle_cfg *c, ocplx *samples_out, const ocplx *img);
int  oracle_recon_radial2d(const oracle_cfg *c, ocplx *h_out, const ocplx *h_in);

/* OpenMP thread count actually used */
int oracle_num_threads(void);

/* RA file format (ra.h:38-48, ra.cu:87-174) */
typedef struct {
    uint64_t flags, eltype, elbyte, size, ndims;
    uint64_t *dims;
    uint8_t *data;
} oracle_ra;
int  oracle_ra_read(oracle_ra *a, const char *path);
int  oracle_ra_write(const oracle_ra *a, const char *path);
void oracle_ra_free(oracle_ra *a);

/* IEEE binary16 conversions (float16.cu:76-166, 261-291) */
uint16_t oracle_floatbits_to_halfbits(uint32_t f);
uint32_t oracle_halfbits_to_floatbits(uint16_t h);

#ifdef __cplusplus
}
#endif
#endif
