/*
 * tron_oracle.c -- CPU restatement of the TRON radial NUFFT hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see tron_oracle.h).  Plain C99 + optional OpenMP.
 * Build: gcc -O3 -fopenmp -march=native -fno-fast-math -shared -fPIC
 *
 * All file:line citations are into /root/reference/src (davidssmith/TRON).
 * Data layout everywhere is the reference's: complex64, channel fastest,
 *   nudata[nchan*(nro*pe + ro) + ch],  udata[nchan*(row*n + col) + ch].
 */
#include "tron_oracle.h"

#include <errno.h>
#include <fcntl.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* golden-angle increment as a float constant, tron.cu:90 */
static const float ORACLE_PHI = 1.9416089796736116f;

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------ */
/* scalar kernels                                                      */
/* ------------------------------------------------------------------ */

/* tron.cu:304-321.  The coefficients are double literals, so the Horner
 * evaluation runs in double on z = x*x (formed in float); numerator and
 * denominator are rounded to float, then divided in float. */
float oracle_besseli0(float x)
{
    static const double P[15] = {
        0.210580722890567e-22, 0.380715242345326e-19, 0.479440257548300e-16,
        0.435125971262668e-13, 0.300931127112960e-10, 0.160224679395361e-7,
        0.654858370096785e-5,  0.202591084143397e-2,  0.463076284721000e0,
        0.754337328948189e2,   0.830792541809429e4,   0.571661130563785e6,
        0.216415572361227e8,   0.356644482244025e9,   0.144048298227235e10 };
    if (x == 0.f) return 1.f;
    float zf = x * x;
    double z = (double)zf;
    double acc = P[0];
    for (int i = 1; i < 15; ++i) acc = acc * z + P[i];
    float num = (float)acc;
    float den = (float)(z * (z * (z - 0.307646912682801e4) + 0.347626332405882e7)
                        - 0.144048298227235e10);
    return -num / den;
}

/* tron.cu:323-349 with the default (non-BEATTY) shape: beta = 2.34f*2.0f*W. */
float oracle_gridkernel(float x, float kernwidth)
{
    float beta = 2.34f * 2.0f * kernwidth;
    if (fabsf(x) < kernwidth) {
        float r = x / kernwidth;
        float f = sqrtf(1.0f - r * r);
        return 0.5f * oracle_besseli0(beta * f) / kernwidth;
    }
    return 0.0f;
}

/* tron.cu:351-370: Fourier transform of the Kaiser-Bessel window. */
float oracle_gridkernelhat(float u, float kernwidth)
{
    float J = 2.0f * kernwidth;
    float beta = 2.34f * 2.0f * kernwidth;
    float r = (float)(M_PI * (double)J * (double)u);   /* double product, tron.cu:357 */
    float q = r * r - beta * beta;
    float y, z;
    if (q > 0) { z = sqrtf(q);  y = sinf(z) / z; }
    else if (q < 0) { z = sqrtf(-q); y = sinhf(z) / z; }
    else y = 1;
    return y;
}

/* tron.cu:372-378 */
float oracle_modang(float x)
{
    const float TWOPI = (float)(2.f * M_PI);
    float y = fmodf(x, TWOPI);
    return y < 0.f ? y + TWOPI : y;
}

/* tron.cu:509: golden angle in f32 from the absolute spoke index; linear angle
 * in double (pe*2.0f is float, M_PI promotes the rest), plus pi/2. */
float oracle_spoke_angle_grid(int pe, int npe, int skip_angles, int golden)
{
    if (golden) return oracle_modang(ORACLE_PHI * (float)(pe + skip_angles));
    return (float)((double)(pe * 2.0f) * M_PI / (double)(float)npe + M_PI * 0.5f);
}

/* tron.cu:555: the forward direction uses pi*pe/npe (half circle, no offset). */
float oracle_spoke_angle_degrid(int pe, int npe, int skip_angles, int golden)
{
    if (golden) return oracle_modang(ORACLE_PHI * (float)(pe + skip_angles));
    return (float)((double)pe * M_PI / (double)(float)npe);
}

/* ------------------------------------------------------------------ */
/* array kernels                                                       */
/* ------------------------------------------------------------------ */

/* tron.cu:405-416: analytic ramp density compensation, in place. */
void oracle_precompensate(ocplx *nudata, int nchan, int nro, int npe1work)
{
    float a = (2.f - 2.f / (float)npe1work) / (float)nro;
    float b = 1.f / (float)npe1work;
#pragma omp parallel for schedule(static)
    for (int pe = 0; pe < npe1work; ++pe)
        for (int r = 0; r < nro; ++r) {
            float sdc = a * fabsf(r - (float)(nro / 2)) + b;
            ocplx *p = nudata + (size_t)nro * nchan * pe + (size_t)nchan * r;
            for (int c = 0; c < nchan; ++c) { p[c].x *= sdc; p[c].y *= sdc; }
        }
}

/* Enumerate the taps of one Cartesian cell in reference order (tron.cu:498-529):
 * for every spoke, radii Rlo..Rhi (aligned) then -Rhi..-Rlo (anti-aligned);
 * r == 0 is therefore visited twice when Rlo == 0.  The in-support decision
 * uses the fused form fma(ct, r, -X) the reference's SASS evaluates. */
typedef void (*tap_fn)(void *ctx, int pe, int r, int ridx, float wgt);

static void oracle_cell_taps(int X, int Y, int nxos, int nro, int npe, float W,
                             int skip_angles, int golden, const float *ct, const float *st,
                             tap_fn fn, void *ctx)
{
    float R = hypotf((float)X, (float)Y);
    int Rhi = (int)fminf(floorf(R + W), (float)(nxos / 2 - 1));
    int Rlo = (int)fmaxf(ceilf(R - W), 0.f);
    (void)skip_angles; (void)golden;
    for (int pe = 0; pe < npe; ++pe) {
        for (int pass = 0; pass < 2; ++pass) {
            int r0 = pass ? -Rhi : Rlo, r1 = pass ? -Rlo : Rhi;
            for (int r = r0; r <= r1; ++r) {
                float dx = fmaf(ct[pe], (float)r, -(float)X);
                float dy = fmaf(st[pe], (float)r, -(float)Y);
                float wgt = oracle_gridkernel(dx, W) * oracle_gridkernel(dy, W);
                if (wgt > 0.f) fn(ctx, pe, r, (r * nro) / nxos, wgt);
            }
        }
    }
}

typedef struct { ocplx *acc; const ocplx *nudata; int nchan, nro; } grid_ctx;

static void grid_tap(void *vctx, int pe, int r, int ridx, float wgt)
{
    grid_ctx *g = (grid_ctx *)vctx; (void)r;
    const ocplx *s = g->nudata + (size_t)g->nchan * ((size_t)g->nro * pe + ridx + g->nro / 2);
    for (int ch = 0; ch < g->nchan; ++ch) {
        g->acc[ch].x += wgt * s[ch].x;
        g->acc[ch].y += wgt * s[ch].y;
    }
}

/* Optional trig override.  The reference takes sin/cos from the GPU's special
 * function unit (__sincosf, tron.cu:511,559), which libm cannot reproduce bit
 * for bit; for axis-aligned linear spokes that decides whole rows of edge taps.
 * The golden fixtures therefore carry the SFU values of every spoke (dumped by
 * oracle/_ref on the GPU), indexed by pe + skip for golden angles and by pe for
 * linear ones, and the tests install them here. */
static const float *g_trig_ct = NULL, *g_trig_st = NULL;
static int g_trig_n = 0;

void oracle_set_trig_table(const float *ct, const float *st, int n)
{
    g_trig_ct = ct; g_trig_st = st; g_trig_n = (ct && st) ? n : 0;
}

static void spoke_tables(float *ct, float *st, int npe, int skip_angles, int golden, int degrid)
{
    for (int pe = 0; pe < npe; ++pe) {
        int k = golden ? pe + skip_angles : pe;
        if (k >= 0 && k < g_trig_n) { ct[pe] = g_trig_ct[k]; st[pe] = g_trig_st[k]; continue; }
        float t = degrid ? oracle_spoke_angle_degrid(pe, npe, skip_angles, golden)
                         : oracle_spoke_angle_grid(pe, npe, skip_angles, golden);
        st[pe] = sinf(t);
        ct[pe] = cosf(t);
    }
}

/* tron.cu:465-536.  The 4x4 thread remap of tron.cu:488-494 only permutes which
 * thread owns which cell (it needs nxos % 4 == 0 to be a bijection); the value
 * written to udata[nchan*id + ch] is what is restated here. */
void oracle_gridradial2d(ocplx *udata, const ocplx *nudata, int nxos, int nchan,
                         int nro, int npe, float W, int skip_angles, int golden)
{
    float *ct = (float *)malloc(sizeof(float) * (size_t)npe * 2), *st = ct + npe;
    spoke_tables(ct, st, npe, skip_angles, golden, 0);
    float scale = 1.f / (float)nxos / (float)npe;          /* tron.cu:532 */
#pragma omp parallel
    {
        ocplx *acc = (ocplx *)malloc(sizeof(ocplx) * (size_t)nchan);
#pragma omp for schedule(dynamic, 64)
        for (int id = 0; id < nxos * nxos; ++id) {
            int Y = id / nxos - nxos / 2, X = id % nxos - nxos / 2;
            memset(acc, 0, sizeof(ocplx) * (size_t)nchan);
            grid_ctx g = { acc, nudata, nchan, nro };
            oracle_cell_taps(X, Y, nxos, nro, npe, W, skip_angles, golden, ct, st, grid_tap, &g);
            for (int ch = 0; ch < nchan; ++ch) {
                udata[(size_t)nchan * id + ch].x = acc[ch].x * scale;
                udata[(size_t)nchan * id + ch].y = acc[ch].y * scale;
            }
        }
        free(acc);
    }
    free(ct);
}

typedef struct { int32_t *hits; long n, max; int id; } hit_ctx;

static void hit_tap(void *vctx, int pe, int r, int ridx, float wgt)
{
    hit_ctx *h = (hit_ctx *)vctx; (void)wgt;
    if (h->n < h->max) {
        int32_t *o = h->hits + 4 * h->n;
        o[0] = h->id; o[1] = pe; o[2] = r; o[3] = ridx;
    }
    h->n++;
}

/* Sample -> cell index map of the gridding operator, in reference order:
 * quadruples (cell id, pe, r, ridx).  Returns the number of taps. */
long oracle_grid_hits(int32_t *hits, long maxhits, int nxos, int nro, int npe,
                      float W, int skip_angles, int golden)
{
    float *ct = (float *)malloc(sizeof(float) * (size_t)npe * 2), *st = ct + npe;
    spoke_tables(ct, st, npe, skip_angles, golden, 0);
    hit_ctx h = { hits, 0, maxhits, 0 };
    for (int id = 0; id < nxos * nxos; ++id) {
        h.id = id;
        oracle_cell_taps(id % nxos - nxos / 2, id / nxos - nxos / 2, nxos, nro, npe, W,
                         skip_angles, golden, ct, st, hit_tap, &h);
    }
    free(ct);
    return h.n;
}

/* tron.cu:540-577.  X walks rows, Y walks columns; (n+1)/2 is an integer
 * division; taps wrap periodically; no output scaling. */
static void degrid_with_trig(ocplx *nudata, const ocplx *udata, int n, int nrep,
                             int nro, int npe, float W, const float *ct, const float *st);

void oracle_degridradial2d(ocplx *nudata, const ocplx *udata, int n, int nrep,
                           int nro, int npe, float W, int skip_angles, int golden)
{
    float *ct = (float *)malloc(sizeof(float) * (size_t)npe * 2), *st = ct + npe;
    spoke_tables(ct, st, npe, skip_angles, golden, 1);
    degrid_with_trig(nudata, udata, n, nrep, nro, npe, W, ct, st);
    free(ct);
}

static void degrid_with_trig(ocplx *nudata, const ocplx *udata, int n, int nrep,
                             int nro, int npe, float W, const float *ct, const float *st)
{
    const float c0 = (float)((n + 1) / 2);
#pragma omp parallel for schedule(static)
    for (int id = 0; id < nro * npe; ++id) {
        ocplx *out = nudata + (size_t)nrep * id;
        for (int c = 0; c < nrep; ++c) out[c].x = out[c].y = 0.f;
        int pe = id / nro, ro = id % nro;
        float R = (float)ro / (float)nro - 0.5f;
        float nR = (float)n * R;
        float X = fmaf(nR, st[pe], c0);
        float Y = fmaf(nR, ct[pe], c0);
        for (int xu = (int)ceilf(X - W); (float)xu <= X + W; ++xu) {
            float wx = oracle_gridkernel((float)xu - X, W);
            for (int yu = (int)ceilf(Y - W); (float)yu <= Y + W; ++yu) {
                float wgt = wx * oracle_gridkernel((float)yu - Y, W);
                int i = (xu + n) % n, j = (yu + n) % n;
                const ocplx *src = udata + (size_t)nrep * ((size_t)i * n + j);
                for (int c = 0; c < nrep; ++c) {
                    out[c].x += wgt * src[c].x;
                    out[c].y += wgt * src[c].y;
                }
            }
        }
    }
}

/* tron.cu:161-178: out-of-place circular shift; FORWARD moves by n/2,
 * INVERSE by n - n/2. */
void oracle_fftshift(ocplx *dst, const ocplx *src, int n, int nchan, int inverse_dir)
{
    int offset = inverse_dir ? n - n / 2 : n / 2;
#pragma omp parallel for schedule(static)
    for (int x = 0; x < n; ++x)
        for (int y = 0; y < n; ++y) {
            size_t s = (size_t)x * n + y;
            size_t d = (size_t)((x + offset) % n) * n + (size_t)((y + offset) % n);
            memcpy(dst + d * nchan, src + s * nchan, sizeof(ocplx) * (size_t)nchan);
        }
}

/* 1-D DFT of arbitrary length in double precision: recursive decimation in
 * time over the smallest prime factor, naive DFT for prime lengths.  This is
 * the stand-in for cufftExecC2C (tron.cu:632,645): unnormalised,
 * sign = -1 -> CUFFT_FORWARD, sign = +1 -> CUFFT_INVERSE. */
static void dft_rec(double *ore, double *oim, const double *ire, const double *iim,
                    int n, int istride, const double *wre, const double *wim, int wstride)
{
    if (n == 1) { ore[0] = ire[0]; oim[0] = iim[0]; return; }
    int p = 0;
    for (int f = 2; f * f <= n; ++f) if (n % f == 0) { p = f; break; }
    if (!p) {                                    /* prime: O(n^2) */
        for (int k = 0; k < n; ++k) {
            double sr = 0, si = 0;
            for (int j = 0; j < n; ++j) {
                int t = (int)(((long)j * k) % n) * wstride;
                sr += ire[(size_t)j * istride] * wre[t] - iim[(size_t)j * istride] * wim[t];
                si += ire[(size_t)j * istride] * wim[t] + iim[(size_t)j * istride] * wre[t];
            }
            ore[k] = sr; oim[k] = si;
        }
        return;
    }
    int m = n / p;
    for (int q = 0; q < p; ++q)                  /* p sub-transforms of length m */
        dft_rec(ore + (size_t)q * m, oim + (size_t)q * m, ire + (size_t)q * istride,
                iim + (size_t)q * istride, m, istride * p, wre, wim, wstride * p);
    double tr[64], ti[64];
    for (int k = 0; k < m; ++k) {
        for (int q = 0; q < p; ++q) {            /* twiddle w_n^{qk} */
            int t = (int)(((long)q * k) % n) * wstride;
            double ar = ore[(size_t)q * m + k], ai = oim[(size_t)q * m + k];
            tr[q] = ar * wre[t] - ai * wim[t];
            ti[q] = ar * wim[t] + ai * wre[t];
        }
        for (int s = 0; s < p; ++s) {            /* p-point DFT across the sub-transforms */
            double sr = 0, si = 0;
            for (int q = 0; q < p; ++q) {
                int t = (int)(((long)q * s * m) % n) * wstride;
                sr += tr[q] * wre[t] - ti[q] * wim[t];
                si += tr[q] * wim[t] + ti[q] * wre[t];
            }
            ore[(size_t)s * m + k] = sr; oim[(size_t)s * m + k] = si;
        }
    }
}

/* batched 2-D transform of nchan channel-interleaved n x n arrays, in place */
void oracle_fft2(ocplx *data, int n, int nchan, int sign)
{
    double *wre = (double *)malloc(sizeof(double) * 2 * (size_t)n), *wim = wre + n;
    for (int k = 0; k < n; ++k) {
        wre[k] = cos(2.0 * M_PI * k / n);
        wim[k] = sign * sin(2.0 * M_PI * k / n);
    }
#pragma omp parallel
    {
        double *buf = (double *)malloc(sizeof(double) * 4 * (size_t)n);
        double *ire = buf, *iim = buf + n, *ore = buf + 2 * n, *oim = buf + 3 * n;
        for (int pass = 0; pass < 2; ++pass) {
            /* pass 0: along columns index (contiguous), pass 1: along rows */
            size_t estride = pass ? (size_t)n * nchan : (size_t)nchan;
            size_t lstride = pass ? (size_t)nchan : (size_t)n * nchan;
#pragma omp for schedule(static) collapse(2)
            for (int line = 0; line < n; ++line)
                for (int ch = 0; ch < nchan; ++ch) {
                    ocplx *base = data + lstride * line + ch;
                    for (int j = 0; j < n; ++j) { ire[j] = base[estride * j].x; iim[j] = base[estride * j].y; }
                    dft_rec(ore, oim, ire, iim, n, 1, wre, wim, 1);
                    for (int j = 0; j < n; ++j) { base[estride * j].x = (float)ore[j]; base[estride * j].y = (float)oim[j]; }
                }
        }
        free(buf);
    }
    free(wre);
}

/* tron.cu:418-431: central ndst x ndst window. */
void oracle_crop(ocplx *dst, int ndst, const ocplx *src, int nsrc, int nchan)
{
    int w = (nsrc - ndst) / 2;
    for (int x = 0; x < ndst; ++x)
        for (int y = 0; y < ndst; ++y)
            memcpy(dst + ((size_t)x * ndst + y) * nchan,
                   src + ((size_t)(x + w) * nsrc + y + w) * nchan, sizeof(ocplx) * (size_t)nchan);
}

/* tron.cu:435-457: centred zero padding.  The strict "> 0" tests drop source
 * row 0 and source column 0. */
void oracle_pad(ocplx *dst, int ndst, const ocplx *src, int nsrc, int nchan)
{
    int w = ndst > nsrc ? (ndst - nsrc) / 2 : 0;
    memset(dst, 0, sizeof(ocplx) * (size_t)ndst * ndst * nchan);
    for (int x = 0; x < ndst; ++x)
        for (int y = 0; y < ndst; ++y)
            if (x - w > 0 && x - w < nsrc && y - w > 0 && y - w < nsrc)
                memcpy(dst + ((size_t)x * ndst + y) * nchan,
                       src + ((size_t)(x - w) * nsrc + (y - w)) * nchan, sizeof(ocplx) * (size_t)nchan);
}

/* tron.cu:390-402.  x = id/float(n) - (n+1)/2 is a float division of the
 * linear index, so it carries the column fraction; division by the weight is a
 * multiplication by its float reciprocal (float2math.h:23). */
void oracle_deapod(ocplx *a, int n, int nrep, float m, float sigma)
{
#pragma omp parallel for schedule(static)
    for (long id = 0; id < (long)n * n; ++id) {
        float x = (float)id / (float)n - (float)((n + 1) / 2);
        float y = (float)(id % n) - (float)((n + 1) / 2);
        float scale = 1.f / (float)n / sigma;
        float wgt = oracle_gridkernelhat(x * scale, m) * oracle_gridkernelhat(y * scale, m);
        float inv = 1.0f / (wgt > 0.f ? wgt : 1.f);
        for (int c = 0; c < nrep; ++c) { a[nrep * id + c].x *= inv; a[nrep * id + c].y *= inv; }
    }
}

/* tron.cu:255-268: root sum of squares, sequential over channels; a single
 * channel passes through as complex. */
void oracle_coilcombinesos(ocplx *img, const ocplx *coilimg, int nimg, int nchan)
{
    for (long id = 0; id < (long)nimg * nimg; ++id) {
        if (nchan > 1) {
            float val = 0.f;
            for (int c = 0; c < nchan; ++c) {
                ocplx z = coilimg[(size_t)nchan * id + c];
                val += z.x * z.x + z.y * z.y;
            }
            img[id].x = sqrtf(val); img[id].y = 0.f;
        } else img[id] = coilimg[id];
    }
}

/* tron.cu:222-253 (powit) and 270-302 (coilcombinewalsh): adaptive coil combine.
 * Per pixel: A[c1][c2] = sum over the (2*npatch+1)^2 patch (clipped at the image
 * border) of z_c1 * conj(z_c2), accumulated px-outer / py-inner; five power
 * iterations from x = (1,..,1) with y/|y| as a multiplication by the float
 * reciprocal (float2math.h:24-28); img = sum_c conj(x_c) * z_c.  The reference
 * zeroes NCHAN*NCHAN = 36 entries whatever nchan is (tron.cu:282, tron.h:50), so
 * it is only defined for nchan <= 6; this restatement zeroes nchan^2.  A patch of
 * all zeros gives 0 * (1/0) = NaN in the reference; here (and in the product) it gives 0. */
void oracle_coilcombinewalsh(ocplx *img, const ocplx *coilimg, int nimg, int nchan, int npatch)
{
    if (nchan == 1) { memcpy(img, coilimg, sizeof(ocplx) * (size_t)nimg * nimg); return; }
#pragma omp parallel
    {
        ocplx *A = (ocplx *)malloc(sizeof(ocplx) * (size_t)nchan * nchan);
        ocplx *x = (ocplx *)malloc(sizeof(ocplx) * nchan), *y = (ocplx *)malloc(sizeof(ocplx) * nchan);
#pragma omp for schedule(static)
        for (long id = 0; id < (long)nimg * nimg; ++id) {
            int px0 = (int)(id / nimg), py0 = (int)(id % nimg);
            for (int k = 0; k < nchan * nchan; ++k) A[k].x = A[k].y = 0.f;
            int xlo = px0 - npatch < 0 ? 0 : px0 - npatch, xhi = px0 + npatch > nimg - 1 ? nimg - 1 : px0 + npatch;
            int ylo = py0 - npatch < 0 ? 0 : py0 - npatch, yhi = py0 + npatch > nimg - 1 ? nimg - 1 : py0 + npatch;
            for (int px = xlo; px <= xhi; ++px)
                for (int py = ylo; py <= yhi; ++py) {
                    const ocplx *z = coilimg + (size_t)nchan * ((size_t)px * nimg + py);
                    for (int c2 = 0; c2 < nchan; ++c2)
                        for (int c1 = 0; c1 < nchan; ++c1) {     /* a * conj(b), float2math.h:38-42 */
                            float ax = z[c1].x, ay = z[c1].y, bx = z[c2].x, by = -z[c2].y;
                            A[c1 * nchan + c2].x += ax * bx - ay * by;
                            A[c1 * nchan + c2].y += ax * by + ay * bx;
                        }
                }
            for (int k = 0; k < nchan; ++k) { x[k].x = 1.f; x[k].y = 0.f; }
            for (int t = 0; t < 5; ++t) {                        /* tron.cu:291: powit(A, nchan, 5) */
                float nsq = 0.f;
                for (int j = 0; j < nchan; ++j) {
                    y[j].x = y[j].y = 0.f;
                    for (int k = 0; k < nchan; ++k) {
                        ocplx a = A[j * nchan + k];
                        y[j].x += a.x * x[k].x - a.y * x[k].y;
                        y[j].y += a.x * x[k].y + a.y * x[k].x;
                    }
                }
                for (int k = 0; k < nchan; ++k) nsq += y[k].x * y[k].x + y[k].y * y[k].y;
                float inv = nsq > 0.f ? 1.0f / sqrtf(nsq) : 0.f;   /* the reference: NaN for a patch of zeros */
                for (int k = 0; k < nchan; ++k) { x[k].x = y[k].x * inv; x[k].y = y[k].y * inv; }
            }
            ocplx o = {0.f, 0.f};
            const ocplx *z = coilimg + (size_t)nchan * id;
            for (int c = 0; c < nchan; ++c) {                    /* conj(x_c) * z_c, tron.cu:294 */
                o.x += x[c].x * z[c].x - (-x[c].y) * z[c].y;
                o.y += x[c].x * z[c].y + (-x[c].y) * z[c].x;
            }
            img[id] = o;
        }
        free(A); free(x); free(y);
    }
}

/* ------------------------------------------------------------------ */
/* geometry and pipelines                                              */
/* ------------------------------------------------------------------ */

void oracle_cfg_defaults(oracle_cfg *c)
{
    memset(c, 0, sizeof *c);
    c->gridos = 2.f; c->kernwidth = 2.f; c->data_undersamp = 1.f;   /* tron.cu:67-74 */
}

/* tron.cu:905-961, int truncations included. */
int oracle_geometry(oracle_cfg *c)
{
    c->nc = (int)c->dims[0]; c->nt = (int)c->dims[1];
    c->out_dims[0] = 1;                                   /* tron.cu:899 (sic, also forward) */
    if (c->adjoint) {
        c->nro = (int)c->dims[2]; c->npe1 = (int)c->dims[3]; c->npe2 = (int)c->dims[4];
        c->nx = c->ny = c->nro / 2;
        c->nxos = (int)(c->nx * c->gridos); c->nyos = (int)(c->ny * c->gridos);
        if ((float)c->npe1 <= (float)c->nro * c->data_undersamp) c->npe1work = c->npe1;
        else c->npe1work = (int)((float)c->nro * c->data_undersamp);
        if (c->prof_slide == 0) c->prof_slide = c->npe1work;
        c->nz = 1 + (c->npe1 - c->npe1work) / c->prof_slide;
        c->out_dims[1] = (uint64_t)c->nt; c->out_dims[2] = (uint64_t)c->nx;
        c->out_dims[3] = (uint64_t)c->ny; c->out_dims[4] = (uint64_t)c->nz;
        c->out_elems = (uint64_t)c->nt * c->nx * c->ny * c->nz;
    } else {
        c->nx = (int)c->dims[2]; c->ny = (int)c->dims[3]; c->nz = (int)c->dims[4];
        c->nxos = (int)(c->nx * c->gridos); c->nyos = (int)(c->ny * c->gridos);
        c->nro = (int)(c->gridos * c->nx);
        c->npe1work = (int)(c->data_undersamp * (float)c->nro);
        c->npe1 = c->npe1work; c->npe2 = 1;
        c->out_dims[1] = (uint64_t)c->nt; c->out_dims[2] = (uint64_t)c->nro;
        c->out_dims[3] = (uint64_t)c->npe1; c->out_dims[4] = (uint64_t)c->npe2;
        c->out_elems = (uint64_t)c->nc * c->nt * c->nro * c->npe1 * c->npe2;
    }
    if (!(c->nc % 2 == 0 || c->nc == 1)) return -1;       /* tron.cu:963 */
    if (c->nt != 1) return -2;                            /* reference is broken for nt>1 */
    return 0;
}

/* tron.cu:623-637: the per-coil images of one adjoint slice (what tron_nufft_adj_radial2d
 * returns).  samples points at the first spoke of the window; peoffset enters only the
 * golden-angle index. */
static void adj_coils_impl(const oracle_cfg *c, ocplx *coilimg_out, const ocplx *samples, int peoffset, int matched)
{
    int nchan = c->nc * c->nt, n = c->nxos;
    size_t ns = (size_t)nchan * c->nro * c->npe1work, ng = (size_t)nchan * n * n;
    ocplx *u = (ocplx *)malloc(sizeof(ocplx) * (ns > ng ? ns : ng));
    ocplx *v = (ocplx *)malloc(sizeof(ocplx) * (ns > ng ? ns : ng));
    memcpy(u, samples, sizeof(ocplx) * ns);
    oracle_precompensate(u, nchan, c->nro, c->npe1work);
    oracle_gridradial2d(v, u, n, nchan, c->nro, c->npe1work, c->kernwidth,
                        c->skip_angles + peoffset, c->golden_angle);
    oracle_fftshift(u, v, n, nchan, 1);
    oracle_fft2(u, n, nchan, +1);                          /* CUFFT_INVERSE */
    oracle_fftshift(v, u, n, nchan, 0);
    if (!matched) {
        oracle_crop(u, c->nx, v, n, nchan);
        oracle_deapod(u, c->nx, nchan, c->kernwidth, c->gridos);
    } else {                                               /* the forward model's weights (tron.cu:643) */
        oracle_deapod(v, n, nchan, c->kernwidth, 1.f);
        oracle_crop(u, c->nx, v, n, nchan);
    }
    memcpy(coilimg_out, u, sizeof(ocplx) * (size_t)nchan * c->nx * c->nx);
    free(u); free(v);
}

void oracle_nufft_adj_coils(const oracle_cfg *c, ocplx *coilimg_out, const ocplx *samples, int peoffset)
{
    adj_coils_impl(c, coilimg_out, samples, peoffset, 0);
}

/* tron.cu:764 (root sum of squares) or the disabled call at tron.cu:766 (Walsh) */
static void combine_coils(const oracle_cfg *c, ocplx *img_out, const ocplx *coilimg)
{
    if (c->coil_combine == 1) oracle_coilcombinewalsh(img_out, coilimg, c->nx, c->nc, c->walsh_npatch);
    else oracle_coilcombinesos(img_out, coilimg, c->nx, c->nc);
}

/* tron.cu:639-649 with the spoke angles of the GRIDDING operator (tron.cu:509, absolute index
 * skip + peoffset): the forward model that the adjoint of window `peoffset` is the adjoint of.
 * The reference's CGNR calls tron_nufft_radial2d, whose angles ignore peoffset and, for linear
 * angles, follow a different formula (tron.cu:555) -- one of the reasons it is marked
 * "NOT WORKING CORRECTLY YET" (tron.cu:670). */
static void fwd_coils_grid_angles(const oracle_cfg *c, ocplx *samples_out, const ocplx *coilimg, int peoffset)
{
    int nchan = c->nc * c->nt, n = c->nxos, npe = c->npe1work;
    size_t ng = (size_t)nchan * n * n, ns = (size_t)nchan * c->nro * npe;
    ocplx *u = (ocplx *)malloc(sizeof(ocplx) * (ns > ng ? ns : ng));
    ocplx *v = (ocplx *)malloc(sizeof(ocplx) * (ns > ng ? ns : ng));
    float *ct = (float *)malloc(sizeof(float) * (size_t)npe * 2), *st = ct + npe;
    oracle_pad(v, n, coilimg, c->nx, nchan);
    oracle_deapod(v, n, nchan, c->kernwidth, 1.f);
    oracle_fftshift(u, v, n, nchan, 0);
    oracle_fft2(u, n, nchan, -1);
    oracle_fftshift(v, u, n, nchan, 1);
    spoke_tables(ct, st, npe, c->skip_angles + peoffset, c->golden_angle, 0);
    degrid_with_trig(samples_out, v, n, nchan, c->nro, npe, c->kernwidth, ct, st);
    free(u); free(v); free(ct);
}

/* Conjugate gradients on the weighted normal equations (-i niter): the algorithm tron.cu:665-720
 * names (Knopp et al. 2007, Algorithm 1), restated so that it converges.  The reference's own
 * version is self-declared broken (tron.cu:670): norms where squared norms belong, image vectors
 * sized nxos^2 although the adjoint returns nx^2, a byte-count memset, forward angles without
 * peoffset, and an operator pair that is not an adjoint pair.  What is restated here keeps the
 * reference's two operators and repairs exactly what breaks the symmetry of B A:
 *   A  = forward model above (pad, deapod(nxos, 1), FFT, degrid with the gridding angles);
 *   B  = tron_nufft_adj_radial2d with the FORWARD model's deapodisation weights (tron.cu:643
 *        instead of 635: the two tables differ by the F9 coordinate quirk), and with row 0 and
 *        column 0 of its output cleared, because pad drops them (tron.cu:449-450), so
 *        B = s A^H W', s = 1/(nxos npe1work) (tron.cu:532);
 *   W' = the weights gridding really applies: the ramp of precompensate (tron.cu:405-416),
 *        doubled at ro = nro/2 (r = 0 is visited twice, tron.cu:512,521) and zero at ro = 0
 *        (r = -nro/2 lies outside Rhi <= nxos/2-1, tron.cu:499).  Needs nro == nxos.
 *     r = M y (M clears ro = 0);  z = B r;  p = z;  x = 0
 *     repeat niter times:  v = A p;  alpha = |z|^2 / (s <v, W' v>);  x += alpha p
 *                          (the last iteration stops here)
 *                          r -= alpha M v;  z' = B r;  beta = |z'|^2 / |z|^2;  p = z' + beta p;  z = z'
 * What remains unsymmetric is the annulus clipping of gridding (F4, ~6e-4 of <r, A p>).  All coils
 * share alpha and beta (the vectors hold every coil, as in tron.cu:679-680); inner products are
 * accumulated in double.  x (per-coil images) is then coil-combined. */
static void zero_first_row_col(ocplx *z, int nx, int nchan)
{
    for (int i = 0; i < nx; ++i)
        for (int ch = 0; ch < nchan; ++ch) {
            z[(size_t)nchan * i + ch].x = z[(size_t)nchan * i + ch].y = 0.f;                       /* row 0 */
            z[(size_t)nchan * ((size_t)i * nx) + ch].x = z[(size_t)nchan * ((size_t)i * nx) + ch].y = 0.f;   /* column 0 */
        }
}

void oracle_cgnr_coils(const oracle_cfg *c, ocplx *x, const ocplx *samples, int peoffset, int niter)
{
    int nchan = c->nc * c->nt;
    size_t N = (size_t)nchan * c->nx * c->nx, n = (size_t)nchan * c->nro * c->npe1work;
    ocplx *r = (ocplx *)malloc(sizeof(ocplx) * n), *v = (ocplx *)malloc(sizeof(ocplx) * n);
    ocplx *z = (ocplx *)malloc(sizeof(ocplx) * N), *p = (ocplx *)malloc(sizeof(ocplx) * N);
    const float s = 1.f / (float)c->nxos / (float)c->npe1work;
    const float wa = (2.f - 2.f / (float)c->npe1work) / (float)c->nro, wb = 1.f / (float)c->npe1work;
    memcpy(r, samples, sizeof(ocplx) * n);
    for (size_t i = 0; i < n; ++i)                         /* ro = 0 has weight 0 in W' */
        if ((i / (size_t)nchan) % (size_t)c->nro == 0) r[i].x = r[i].y = 0.f;
    adj_coils_impl(c, z, r, peoffset, 1);
    zero_first_row_col(z, c->nx, nchan);
    memcpy(p, z, sizeof(ocplx) * N);
    memset(x, 0, sizeof(ocplx) * N);
    double zz = 0.0;
    for (size_t i = 0; i < N; ++i) zz += (double)z[i].x * z[i].x + (double)z[i].y * z[i].y;
    for (int t = 0; t < niter; ++t) {
        fwd_coils_grid_angles(c, v, p, peoffset);
        double vwv = 0.0;
        for (size_t i = 0; i < n; ++i) {
            int ro = (int)((i / (size_t)nchan) % (size_t)c->nro);
            float w = wa * fabsf((float)ro - (float)(c->nro / 2)) + wb;
            if (ro == c->nro / 2) w += w;                  /* gridding visits r = 0 twice (tron.cu:512,521) */
            if (ro == 0) w = 0.f;                          /* r = -nro/2 lies outside Rhi <= nxos/2-1 (tron.cu:499) */
            vwv += (double)w * ((double)v[i].x * v[i].x + (double)v[i].y * v[i].y);
        }
        float alpha = vwv > 0.0 ? (float)(zz / ((double)s * vwv)) : 0.f;
        for (size_t i = 0; i < N; ++i) { x[i].x += alpha * p[i].x; x[i].y += alpha * p[i].y; }
        if (t == niter - 1) break;
        for (size_t i = 0; i < n; ++i) {
            if ((i / (size_t)nchan) % (size_t)c->nro == 0) continue;
            r[i].x -= alpha * v[i].x; r[i].y -= alpha * v[i].y;
        }
        adj_coils_impl(c, z, r, peoffset, 1);
        zero_first_row_col(z, c->nx, nchan);
        double zz2 = 0.0;
        for (size_t i = 0; i < N; ++i) zz2 += (double)z[i].x * z[i].x + (double)z[i].y * z[i].y;
        float beta = zz > 0.0 ? (float)(zz2 / zz) : 0.f;
        for (size_t i = 0; i < N; ++i) { p[i].x = z[i].x + beta * p[i].x; p[i].y = z[i].y + beta * p[i].y; }
        zz = zz2;
    }
    free(r); free(v); free(z); free(p);
}

/* tron.cu:753-764: one adjoint slice = (CGNR | plain adjoint) then the coil combine */
void oracle_nufft_adj_slice(const oracle_cfg *c, ocplx *img_out, const ocplx *samples, int peoffset)
{
    int nchan = c->nc * c->nt;
    ocplx *u = (ocplx *)malloc(sizeof(ocplx) * (size_t)nchan * c->nx * c->nx);
    if (c->niter > 0) oracle_cgnr_coils(c, u, samples, peoffset, c->niter);
    else oracle_nufft_adj_coils(c, u, samples, peoffset);
    combine_coils(c, img_out, u);
    free(u);
}

/* tron.cu:639-649: one forward slice. */
void oracle_nufft_fwd_slice(const oracle_cfg *c, ocplx *samples_out, const ocplx *img)
{
    int nchan = c->nc * c->nt, n = c->nxos;
    size_t ng = (size_t)nchan * n * n, ns = (size_t)nchan * c->nro * c->npe1work;
    ocplx *u = (ocplx *)malloc(sizeof(ocplx) * (ns > ng ? ns : ng));
    ocplx *v = (ocplx *)malloc(sizeof(ocplx) * (ns > ng ? ns : ng));
    oracle_pad(v, n, img, c->nx, nchan);
    oracle_deapod(v, n, nchan, c->kernwidth, 1.f);
    oracle_fftshift(u, v, n, nchan, 0);
    oracle_fft2(u, n, nchan, -1);                          /* CUFFT_FORWARD */
    oracle_fftshift(v, u, n, nchan, 1);
    oracle_degridradial2d(samples_out, v, n, nchan, c->nro, c->npe1work, c->kernwidth,
                          c->skip_angles, c->golden_angle);
    free(u); free(v);
}

/* tron.cu:726-786: slice loop.  Forward mode is restated for nz == 1 only
 * (the reference reads every slice from offset 0 and overruns its output for
 * nz > 1). */
int oracle_recon_radial2d(const oracle_cfg *c, ocplx *h_out, const ocplx *h_in)
{
    if (c->adjoint) {
        for (int z = 0; z < c->nz; ++z) {
            int peoffset = z * c->prof_slide;
            oracle_nufft_adj_slice(c, h_out + (size_t)c->nt * c->nx * c->ny * z,
                                   h_in + (size_t)c->nc * c->nt * c->nro * peoffset, peoffset);
        }
    } else {
        if (c->nz != 1) return -3;
        oracle_nufft_fwd_slice(c, h_out, h_in);
    }
    return 0;
}

/* ------------------------------------------------------------------ */
/* RA container (ra.h:38-48; ra.cu:87-174)                             */
/* ------------------------------------------------------------------ */

static const uint64_t ORACLE_RA_MAGIC = 0x7961727261776172ULL;   /* ra.h:51 */

static int rd_all(int fd, void *buf, uint64_t n)
{
    uint8_t *p = (uint8_t *)buf;
    while (n) {
        size_t chunk = n < (1ULL << 30) ? (size_t)n : (size_t)(1ULL << 30);
        ssize_t got = read(fd, p, chunk);
        if (got <= 0) return -1;
        p += got; n -= (uint64_t)got;
    }
    return 0;
}

static int wr_all(int fd, const void *buf, uint64_t n)
{
    const uint8_t *p = (const uint8_t *)buf;
    while (n) {
        size_t chunk = n < (1ULL << 30) ? (size_t)n : (size_t)(1ULL << 30);
        ssize_t put = write(fd, p, chunk);
        if (put <= 0) return -1;
        p += put; n -= (uint64_t)put;
    }
    return 0;
}

/* header = magic, flags, eltype, elbyte, size, ndims (6 x u64), then ndims x u64 */
int oracle_ra_read(oracle_ra *a, const char *path)
{
    int fd = open(path, O_RDONLY);
    if (fd < 0) return -errno;
    uint64_t h[6];
    if (rd_all(fd, h, sizeof h) || h[0] != ORACLE_RA_MAGIC) { close(fd); return -EINVAL; }
    a->flags = h[1]; a->eltype = h[2]; a->elbyte = h[3]; a->size = h[4]; a->ndims = h[5];
    a->dims = (uint64_t *)malloc(sizeof(uint64_t) * a->ndims);
    a->data = (uint8_t *)malloc(a->size ? a->size : 1);
    int bad = rd_all(fd, a->dims, sizeof(uint64_t) * a->ndims) || rd_all(fd, a->data, a->size);
    close(fd);
    return bad ? -EIO : 0;
}

int oracle_ra_write(const oracle_ra *a, const char *path)
{
    int fd = open(path, O_WRONLY | O_TRUNC | O_CREAT, 0644);
    if (fd < 0) return -errno;
    uint64_t h[6] = { ORACLE_RA_MAGIC, a->flags, a->eltype, a->elbyte, a->size, a->ndims };
    int bad = wr_all(fd, h, sizeof h) || wr_all(fd, a->dims, sizeof(uint64_t) * a->ndims)
              || wr_all(fd, a->data, a->size);
    close(fd);
    return bad ? -EIO : 0;
}

void oracle_ra_free(oracle_ra *a)
{
    free(a->dims); free(a->data); a->dims = NULL; a->data = NULL;
}

/* ------------------------------------------------------------------ */
/* binary16 (float16.cu:76-166, 261-291)                               */
/* ------------------------------------------------------------------ */

/* Round-to-nearest-even on the 13 dropped bits.  In the subnormal branch the
 * significand is shifted right BEFORE the tie test, so sticky bits below the
 * shifted-out position are lost (float16.cu:112-126); this differs from IEEE
 * RNE for a few inputs with 2^-25 < |f| < 2^-14 and is kept on purpose. */
uint16_t oracle_floatbits_to_halfbits(uint32_t f)
{
    uint16_t sign = (uint16_t)((f >> 16) & 0x8000u);
    uint32_t e = f & 0x7f800000u, m = f & 0x007fffffu;
    if (e >= 0x47800000u) {                       /* |f| >= 65536, inf, nan */
        if (e == 0x7f800000u && m) {
            uint16_t q = (uint16_t)(0x7c00u + (m >> 13));
            if (q == 0x7c00u) q++;                /* keep it a NaN */
            return (uint16_t)(sign + q);
        }
        return (uint16_t)(sign + 0x7c00u);
    }
    if (e <= 0x38000000u) {                       /* result is zero or subnormal */
        if (e < 0x33000000u) return sign;
        uint32_t s = (0x00800000u + m) >> (113 - (e >> 23));
        if ((s & 0x3fffu) != 0x1000u) s += 0x1000u;
        return (uint16_t)(sign + (uint16_t)(s >> 13));
    }
    uint16_t he = (uint16_t)((e - 0x38000000u) >> 13);
    if ((m & 0x3fffu) != 0x1000u) m += 0x1000u;
    return (uint16_t)(sign + he + (uint16_t)(m >> 13));  /* carry may bump the exponent */
}

uint32_t oracle_halfbits_to_floatbits(uint16_t h)
{
    uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
    uint32_t e = h & 0x7c00u, m = h & 0x03ffu;
    if (e == 0x7c00u) return sign + 0x7f800000u + (m << 13);
    if (e == 0) {
        if (!m) return sign;
        int shift = 0;                            /* normalise the subnormal */
        while (!(m & 0x0400u)) { m <<= 1; shift++; }
        return sign + ((uint32_t)(127 - 15 - shift + 1) << 23) + ((m & 0x03ffu) << 13);
    }
    return sign + (((uint32_t)(h & 0x7fffu) + 0x1c000u) << 13);
}
