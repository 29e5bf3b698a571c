/*
 * legacy.cu -- the reference's own exported symbols, implemented on the plan API.
 *
 * Surface replaced (all extern "C" in the reference, tron.cu:460-788, tron.h:55-72):
 *   gridradial2d, degridradial2d   __global__ kernels, callable with any <<<blocks,threads>>>
 *   tron_init, tron_shutdown, tron_nufft_adj_radial2d, tron_nufft_radial2d
 *   recon_radial2d (and the header's spelling recon_radial_2d)
 *   tron_cgnr_radial2d, copy, Caxpy (tron.cu:651-720)
 *
 * In the reference these read file-static globals that only main() assigns
 * (tron.cu:54-87); tron_set_config() is the replacement for those assignments.
 * The void entry points have no error channel; like the reference they print
 * and exit non-zero on failure (without the reference's getchar()).
 */
#include "tron_internal.h"

#include <stdlib.h>

using namespace tronb;

static tron_config g_cfg;
static bool g_cfg_set = false;
static tron_plan *g_adj = nullptr, *g_fwd = nullptr, *g_cg = nullptr;

static void die(const char *where)
{
    fprintf(stderr, "tron: %s failed: %s\n", where, tron_last_error());
    exit(EXIT_FAILURE);
}

extern "C" int tron_set_config(const tron_config *cfg)
{
    tron_geometry g;
    int rc = tron_geometry_compute(cfg, &g);
    if (rc) return rc;
    g_cfg = *cfg; g_cfg_set = true;
    return TRON_OK;
}

extern "C" void tron_shutdown(void)
{
    tron_plan_destroy(g_adj); tron_plan_destroy(g_fwd); tron_plan_destroy(g_cg);
    g_adj = g_fwd = g_cg = nullptr;
}

/* tron.cu:579-606 allocates two stream slots; here one plan per direction owns
 * its streams.  The adjoint plan reconstructs the first window (peoffset = 0)
 * and returns per-coil images, which is what tron_nufft_adj_radial2d yields. */
extern "C" void tron_init(void)
{
    if (!g_cfg_set) { set_error("tron_set_config() has not been called"); die("tron_init"); }
    tron_shutdown();
    tron_config c = g_cfg;
    c.slice_begin = 0; c.slice_end = 1; c.per_coil_out = 1; c.sos_partial = 0;
    if (c.adjoint) { if (tron_plan_create(&g_adj, &c)) die("tron_init"); }
    else { if (tron_plan_create(&g_fwd, &c)) die("tron_init"); }
}

extern "C" void tron_nufft_adj_radial2d(tron_float2 *d_out, tron_float2 *d_in, const int j)
{
    (void)j;
    if (!g_adj) { set_error("tron_init() has not created an adjoint plan"); die("tron_nufft_adj_radial2d"); }
    if (tron_recon_device(g_adj, d_out, d_in, g_adj->stream)) die("tron_nufft_adj_radial2d");
    cudaStreamSynchronize(g_adj->stream);
}

extern "C" void tron_nufft_radial2d(tron_float2 *d_out, tron_float2 *d_in, const int j)
{
    (void)j;
    if (!g_fwd) { set_error("tron_init() has not created a forward plan"); die("tron_nufft_radial2d"); }
    if (tron_recon_device(g_fwd, d_out, d_in, g_fwd->stream)) die("tron_nufft_radial2d");
    cudaStreamSynchronize(g_fwd->stream);
}

/* tron.cu:665-720: CGNR on the first window; per-coil images out, like the reference's d_p.
 * The plan is kept between calls with the same iteration count. */
extern "C" void tron_cgnr_radial2d(tron_float2 *d_out, tron_float2 *d_in, const int j, const int niter)
{
    (void)j;
    if (!g_cfg_set) { set_error("tron_set_config() has not been called"); die("tron_cgnr_radial2d"); }
    if (g_cg && g_cg->cfg.niter != niter) { tron_plan_destroy(g_cg); g_cg = nullptr; }
    if (!g_cg) {
        tron_config c = g_cfg;
        c.adjoint = 1; c.niter = niter; c.slice_begin = 0; c.slice_end = 1; c.per_coil_out = 1; c.sos_partial = 0;
        c.coil_combine = 0;
        if (tron_plan_create(&g_cg, &c)) die("tron_cgnr_radial2d");
    }
    if (tron_recon_device(g_cg, d_out, d_in, g_cg->stream)) die("tron_cgnr_radial2d");
    cudaStreamSynchronize(g_cg->stream);
}

/* tron.cu:651-655 */
extern "C" void copy(tron_float2 *d_dst, tron_float2 *d_src, const size_t N, const int j)
{
    (void)j;
    tron_plan *p = g_adj ? g_adj : (g_fwd ? g_fwd : g_cg);
    cudaStream_t s = p ? p->stream : nullptr;
    if (cudaMemcpyAsync(d_dst, d_src, N * sizeof(float2), cudaMemcpyDeviceToDevice, s) != cudaSuccess) {
        set_error("cudaMemcpyAsync failed"); die("copy");
    }
}

/* tron.cu:658-663, same parameter list; size_t index so that N >= 2^31 works */
extern "C" __global__ void Caxpy(float2 *d_z, float2 *d_y, float2 *d_x, float alpha, const size_t N)
{
    for (size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x; id < N; id += (size_t)blockDim.x * gridDim.x) {
        const float2 y = d_y[id], x = d_x[id];
        d_z[id] = make_float2(fmaf(alpha, x.x, y.x), fmaf(alpha, x.y, y.y));
    }
}

extern "C" int tron_launch_Caxpy(void *d_z, const void *d_y, const void *d_x, float alpha, size_t N,
                                 int blocks, int threads, void *stream)
{
    if (blocks < 1 || threads < 1 || threads > 1024) { set_error("bad launch configuration"); return TRON_EINVAL; }
    Caxpy<<<blocks, threads, 0, (cudaStream_t)stream>>>((float2 *)d_z, (float2 *)d_y, (float2 *)d_x, alpha, N);
    TRON_CUDA(cudaGetLastError());
    return TRON_OK;
}

/* tron.cu:726-786: init, slice loop, shutdown -- here: plan, whole job, destroy */
extern "C" void recon_radial2d(tron_float2 *h_outdata, const tron_float2 *h_indata)
{
    if (!g_cfg_set) { set_error("tron_set_config() has not been called"); die("recon_radial2d"); }
    tron_plan *p = nullptr;
    if (tron_plan_create(&p, &g_cfg)) die("recon_radial2d");
    if (tron_recon_host(p, h_outdata, h_indata)) die("recon_radial2d");
    tron_plan_destroy(p);
}

extern "C" void recon_radial_2d(tron_float2 *h_outdata, const tron_float2 *h_indata)
{
    recon_radial2d(h_outdata, h_indata);
}

/* ---------------------------------------------------------------------- */
/* compatibility kernels                                                   */
/* ---------------------------------------------------------------------- */

/* Same contract as tron.cu:465-536 (no density compensation, channel-interleaved
 * output, scale 1/nxos/npe).  Grid-stride over cells; per spoke the candidate
 * radii come from the two support intervals, each then decided by the
 * reference predicate. */
extern "C" __global__ void
gridradial2d(float2 *udata, const float2 *__restrict__ nudata, const int nxos, const int nchan,
             const int nro, const int npe, const float kernwidth, const float gridos,
             const int skip_angles, const int flag_golden_angle)
{
    (void)gridos;
    const KbParams kb = make_kb_basic(kernwidth);
    const float W = kernwidth;
    const float scale = div_approx(rcp_approx((float)nxos), (float)npe);
    for (int id = blockIdx.x * blockDim.x + threadIdx.x; id < nxos * nxos; id += blockDim.x * gridDim.x) {
        const int Y = id / nxos - nxos / 2, X = id % nxos - nxos / 2;
        const float Xf = (float)X, Yf = (float)Y;
        const float R = ref_hypotf(Xf, Yf);
        const int Rhi = (int)fminf(floorf(R + W), (float)(nxos / 2 - 1));
        const int Rlo = (int)fmaxf(ceilf(R - W), 0.f);
        float2 *out = udata + (size_t)nchan * id;
        for (int ch = 0; ch < nchan; ++ch) out[ch] = make_float2(0.f, 0.f);
        for (int pe = 0; pe < npe && Rlo <= Rhi; ++pe) {
            const float t = ref_angle_grid(pe, npe, skip_angles, flag_golden_angle);
            const float st = sin_approx(t), ct = cos_approx(t);
            const float ic = fabsf(ct) > 1e-18f ? 1.0f / ct : copysignf(1e18f, ct);
            const float is = fabsf(st) > 1e-18f ? 1.0f / st : copysignf(1e18f, st);
            float ax = (Xf - W) * ic, bx = (Xf + W) * ic, ay = (Yf - W) * is, by = (Yf + W) * is;
            float lo = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)) - 1e-3f, -(float)Rhi);
            float hi = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)) + 1e-3f, (float)Rhi);
            for (int r = (int)ceilf(lo); r <= (int)floorf(hi); ++r) {
                if (abs(r) < Rlo) continue;
                float dx = fma_ftz(ct, (float)r, -Xf), dy = fma_ftz(st, (float)r, -Yf);
                if (!(fabsf(dx) < W) || !(fabsf(dy) < W)) continue;
                float w = kb_weight(dx, kb) * kb_weight(dy, kb);
                if (!(w > 0.f)) continue;
                if (r == 0) w += w;
                const float2 *s = nudata + (size_t)nchan * ((size_t)nro * pe + (r * nro) / nxos + nro / 2);
                for (int ch = 0; ch < nchan; ++ch) {
                    out[ch].x = fmaf(w, s[ch].x, out[ch].x);
                    out[ch].y = fmaf(w, s[ch].y, out[ch].y);
                }
            }
        }
        for (int ch = 0; ch < nchan; ++ch) { out[ch].x *= scale; out[ch].y *= scale; }
    }
}

/* Same contract as tron.cu:540-577 (channel-interleaved grid in, samples out). */
extern "C" __global__ void
degridradial2d(float2 *nudata, const float2 *__restrict__ udata, const int n, const int nrep,
               const int nro, const int npe, const float W, const float gridos,
               const int skip_angles, const int flag_golden_angle)
{
    (void)gridos;
    const KbParams kb = make_kb_basic(W);
    const float c0 = (float)((n + 1) / 2);
    const float inv_nro = rcp_approx((float)nro);
    for (int id = blockIdx.x * blockDim.x + threadIdx.x; id < nro * npe; id += blockDim.x * gridDim.x) {
        float2 *out = nudata + (size_t)nrep * id;
        for (int c = 0; c < nrep; ++c) out[c] = make_float2(0.f, 0.f);
        const int pe = id / nro, ro = id % nro;
        const float T = ref_angle_degrid(pe, npe, skip_angles, flag_golden_angle);
        const float nR = mul_ftz(fma_ftz((float)ro, inv_nro, -0.5f), (float)n);
        const float X = fma_ftz(sin_approx(T), nR, c0), Y = fma_ftz(cos_approx(T), nR, c0);
        for (int xu = (int)ceilf(X - W); (float)xu <= X + W; ++xu) {
            float dx = (float)xu - X;
            if (!(fabsf(dx) < W)) continue;
            const float wx = kb_weight(dx, kb);
            for (int yu = (int)ceilf(Y - W); (float)yu <= Y + W; ++yu) {
                float dy = (float)yu - Y;
                if (!(fabsf(dy) < W)) continue;
                const float w = wx * kb_weight(dy, kb);
                const float2 *src = udata + (size_t)nrep * ((size_t)((xu + n) % n) * n + (yu + n) % n);
                for (int c = 0; c < nrep; ++c) {
                    out[c].x = fmaf(w, src[c].x, out[c].x);
                    out[c].y = fmaf(w, src[c].y, out[c].y);
                }
            }
        }
    }
}

/* host launchers so that FFI users (no <<<>>> syntax) can drive the two kernels */
extern "C" int tron_launch_gridradial2d(void *udata, const void *nudata, int nxos, int nchan, int nro, int npe,
                                        float kernwidth, float gridos, int skip_angles, int golden,
                                        int blocks, int threads, void *stream)
{
    gridradial2d<<<blocks, threads, 0, (cudaStream_t)stream>>>((float2 *)udata, (const float2 *)nudata, nxos, nchan,
                                                              nro, npe, kernwidth, gridos, skip_angles, golden);
    TRON_CUDA(cudaGetLastError());
    return TRON_OK;
}

extern "C" int tron_launch_degridradial2d(void *nudata, const void *udata, int n, int nrep, int nro, int npe,
                                          float W, float gridos, int skip_angles, int golden,
                                          int blocks, int threads, void *stream)
{
    degridradial2d<<<blocks, threads, 0, (cudaStream_t)stream>>>((float2 *)nudata, (const float2 *)udata, n, nrep,
                                                                nro, npe, W, gridos, skip_angles, golden);
    TRON_CUDA(cudaGetLastError());
    return TRON_OK;
}
