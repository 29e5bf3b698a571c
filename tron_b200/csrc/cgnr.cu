/*
 * cgnr.cu -- iterative reconstruction (-i niter): conjugate gradients on the weighted normal
 * equations, on top of the gridding / degridding pair.
 *
 * Replaces tron_cgnr_radial2d (/root/reference/src/tron.cu:665-720, call at tron.cu:754-755),
 * which the reference marks "NOT WORKING CORRECTLY YET" (tron.cu:670): it divides norms where
 * Knopp et al. 2007 (Algorithm 1) divide squared norms, sizes its image vectors nxos^2 although
 * the adjoint returns nx^2, clears p with a byte count, leaves peoffset out of the forward angles,
 * and its two operators are not an adjoint pair.  What runs here keeps the reference's operators
 * and repairs what breaks the symmetry of B A (DESIGN.md section 3.6 derives each repair):
 *   A  = pad, deapod(nxos, 1), FFT, degridding with the spoke angles of the window's GRIDDING
 *        operator (tron.cu:509 with skip + peoffset);
 *   B  = gridding + inverse FFT + crop with the forward model's deapodisation table and with row 0
 *        and column 0 cleared (pad drops them, tron.cu:449-450): B = s A^H W', s = 1/(nxos npe);
 *   W' = ramp of precompensate (tron.cu:405-416), doubled at ro = nro/2 (r = 0 is gridded twice,
 *        tron.cu:512,521), zero at ro = 0 (never gridded, tron.cu:499).
 *     r = y;  z = B r;  p = z;  x = 0
 *     niter times:  v = A p;  alpha = |z|^2 / (s <v, W' v>);  x += alpha p;  (last: stop)
 *                   r -= alpha v;  z' = B r;  beta = |z'|^2 / |z|^2;  p = z' + beta p
 * All coils of a slice share alpha and beta (tron.cu:679-680).
 *
 * Everything stays on the device and on one stream: a batch of slices advances together, each
 * slice with its own scalars; inner products are two-stage (64 double partials per slice, summed by
 * the consumer kernel in a fixed order, so results are reproducible run to run), and the scalars
 * never visit the host (the reference synchronises on cuBLAS results four times per iteration).
 */
#include "tron_internal.h"

namespace tronb {

constexpr int CG_PARTS = 64;
constexpr int CG_THREADS = 256;

__device__ __forceinline__ double block_sum(double v)
{
    __shared__ double sh[CG_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int i = 0; i < CG_THREADS / 32; ++i) t += sh[i];
    return t;                                            /* valid in thread 0 */
}

/* z [nb][nx][nx][nc]: clear row 0 and column 0, part[b][blockIdx.x] = partial |z_b|^2.
 * One thread per pixel (one integer modulo per nc elements: with one per element these streaming kernels
 * were bound by the division sequence, not by HBM). */
__global__ void __launch_bounds__(CG_THREADS)
cg_zz_kernel(float2 *__restrict__ z, double *__restrict__ part, int nx, int nc)
{
    const size_t npix = (size_t)nx * nx;
    float2 *zb = z + (size_t)blockIdx.y * npix * nc;
    float acc = 0.f;
    for (size_t pix = (size_t)blockIdx.x * CG_THREADS + threadIdx.x; pix < npix; pix += (size_t)gridDim.x * CG_THREADS) {
        float2 *q = zb + pix * nc;
        if (pix < (size_t)nx || pix % nx == 0) {
            for (int c = 0; c < nc; ++c) q[c] = make_float2(0.f, 0.f);
            continue;
        }
        for (int c = 0; c < nc; ++c) acc = fmaf(q[c].x, q[c].x, fmaf(q[c].y, q[c].y, acc));
    }
    const double t = block_sum((double)acc);
    if (threadIdx.x == 0) part[(size_t)blockIdx.y * CG_PARTS + blockIdx.x] = t;
}

__device__ __forceinline__ float cg_weight(int ro, int nro, float wa, float wb)
{
    float w = fmaf(wa, fabsf((float)(ro - nro / 2)), wb);      /* tron.cu:412 */
    if (ro == nro / 2) w += w;
    if (ro == 0) w = 0.f;
    return w;
}

/* v [nb][npe][nro][nc]: part[b][blockIdx.x] = partial <v_b, W' v_b>; npos = npe * nro sample positions */
__global__ void __launch_bounds__(CG_THREADS)
cg_vwv_kernel(const float2 *__restrict__ v, double *__restrict__ part, size_t npos, int nro, int nc, float wa, float wb)
{
    const float2 *vb = v + (size_t)blockIdx.y * npos * nc;
    float acc = 0.f;
    for (size_t pos = (size_t)blockIdx.x * CG_THREADS + threadIdx.x; pos < npos; pos += (size_t)gridDim.x * CG_THREADS) {
        const float w = cg_weight((int)(pos % nro), nro, wa, wb);
        const float2 *q = vb + pos * nc;
        float sq = 0.f;
        for (int c = 0; c < nc; ++c) sq = fmaf(q[c].x, q[c].x, fmaf(q[c].y, q[c].y, sq));
        acc = fmaf(w, sq, acc);
    }
    const double t = block_sum((double)acc);
    if (threadIdx.x == 0) part[(size_t)blockIdx.y * CG_PARTS + blockIdx.x] = t;
}

__device__ __forceinline__ double sum_parts(const double *part)
{
    double t = 0.0;
    for (int i = 0; i < CG_PARTS; ++i) t += part[i];
    return t;
}

/* x += alpha p;  r -= alpha v (ro != 0) unless this is the last iteration */
__global__ void __launch_bounds__(CG_THREADS)
cg_step_kernel(float2 *__restrict__ x, const float2 *__restrict__ p, float2 *__restrict__ r, const float2 *__restrict__ v,
               const double *__restrict__ part_zz, const double *__restrict__ part_vwv,
               size_t N, size_t n, int nro, int nc, float s, int update_r)
{
    __shared__ float sh_alpha;
    const int b = blockIdx.y;
    if (threadIdx.x == 0) {
        const double zz = sum_parts(part_zz + (size_t)b * CG_PARTS), vwv = sum_parts(part_vwv + (size_t)b * CG_PARTS);
        sh_alpha = vwv > 0.0 ? (float)(zz / ((double)s * vwv)) : 0.f;
    }
    __syncthreads();
    const float alpha = sh_alpha;
    float2 *xb = x + (size_t)b * N; const float2 *pb = p + (size_t)b * N;
    for (size_t i = (size_t)blockIdx.x * CG_THREADS + threadIdx.x; i < N; i += (size_t)gridDim.x * CG_THREADS) {
        float2 a = xb[i]; const float2 q = pb[i];
        a.x = fmaf(alpha, q.x, a.x); a.y = fmaf(alpha, q.y, a.y);
        xb[i] = a;
    }
    if (!update_r) return;
    float2 *rb = r + (size_t)b * n; const float2 *vb = v + (size_t)b * n;
    for (size_t i = (size_t)blockIdx.x * CG_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * CG_THREADS) {
        float2 a = rb[i]; const float2 q = vb[i];          /* ro = 0 rows of r are zero and stay unused by the */
        a.x = fmaf(-alpha, q.x, a.x); a.y = fmaf(-alpha, q.y, a.y);   /* gridding (W' = 0 there): no test needed */
        rb[i] = a;
    }
}

/* p = z + beta p, beta = |z'|^2 / |z|^2 (first: p = z) */
__global__ void __launch_bounds__(CG_THREADS)
cg_dir_kernel(float2 *__restrict__ p, const float2 *__restrict__ z, const double *__restrict__ part_new,
              const double *__restrict__ part_old, size_t N, int first)
{
    __shared__ float sh_beta;
    const int b = blockIdx.y;
    if (threadIdx.x == 0) {
        float beta = 0.f;
        if (!first) {
            const double zn = sum_parts(part_new + (size_t)b * CG_PARTS), zo = sum_parts(part_old + (size_t)b * CG_PARTS);
            beta = zo > 0.0 ? (float)(zn / zo) : 0.f;
        }
        sh_beta = beta;
    }
    __syncthreads();
    const float beta = sh_beta;
    float2 *pb = p + (size_t)b * N; const float2 *zb = z + (size_t)b * N;
    for (size_t i = (size_t)blockIdx.x * CG_THREADS + threadIdx.x; i < N; i += (size_t)gridDim.x * CG_THREADS) {
        const float2 q = zb[i];
        float2 a = first ? make_float2(0.f, 0.f) : pb[i];
        a.x = fmaf(beta, a.x, q.x); a.y = fmaf(beta, a.y, q.y);
        pb[i] = a;
    }
}

/* r_b = window of slice z0 + b (complex64 or complex-half) */
__global__ void __launch_bounds__(CG_THREADS)
cg_init_r_kernel(float2 *__restrict__ r, const void *__restrict__ in, size_t n, size_t window_stride, int half_in)
{
    float2 *rb = r + (size_t)blockIdx.y * n;
    const size_t off = (size_t)blockIdx.y * window_stride;
    for (size_t i = (size_t)blockIdx.x * CG_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * CG_THREADS)
        rb[i] = half_in ? __half22float2(((const __half2 *)in)[off + i]) : ((const float2 *)in)[off + i];
}

/* B: gridding + inverse FFT passes -> per-coil images in `coil`, forward deapodisation table */
static int cg_apply_adjoint(tron_plan *p, const void *samples, int half, int data_slide, float2 *coil, int z0, int nb,
                            const float *deapod, cudaStream_t s)
{
    const tron_geometry &g = p->g;
    GridLaunch L = make_grid_launch(p, samples, p->d_grid, z0, nb);
    L.slide = data_slide; L.half_in = half; L.zero_r2 = p->zero_r2;
    int rc = launch_grid(L, s);
    if (rc) return rc;
    AdjFftLaunch a;
    a.grid = p->d_grid; a.tmp = p->d_tmp; a.deapod = deapod; a.out = coil;
    a.nslices = nb; a.nch = p->nch; a.nc_total = g.nc * g.nt; a.ch0 = g.coil_begin;
    a.mode = 2; a.half_out = 0; a.zero_r2 = p->zero_r2;
    p->last_launches += 3;
    return launch_adj_fft(p->fft, a, s);
}

/* per-coil images of slices [z0, z0 + nb) -> p->d_coil: plain adjoint (niter = 0) or CGNR */
int run_percoil_batch(tron_plan *p, const void *d_in, int z0, int nb, cudaStream_t s)
{
    const tron_geometry &g = p->g;
    const int nc = g.nc * g.nt, niter = p->cfg.niter;
    if (niter <= 0)
        return cg_apply_adjoint(p, d_in, p->cfg.half_in, g.prof_slide, p->d_coil, z0, nb, p->deapod_adj, s);

    const size_t N = (size_t)g.nx * g.nx * nc, n = (size_t)g.npe1work * g.nro * nc;
    const size_t in_esz = p->in_elem_bytes;
    const float sc = 1.f / (float)g.nxos / (float)g.npe1work;
    const float wa = (2.f - 2.f / (float)g.npe1work) / (float)g.nro, wb = 1.f / (float)g.npe1work;
    const dim3 gr(CG_PARTS, nb);
    double *part_zz[2] = { p->cg_part, p->cg_part + (size_t)p->batch * CG_PARTS };
    double *part_vwv = p->cg_part + 2 * (size_t)p->batch * CG_PARTS;
    float2 *x = p->d_coil;

    int rc = cg_apply_adjoint(p, d_in, p->cfg.half_in, g.prof_slide, p->cg_z, z0, nb, p->deapod_fwd, s);
    if (rc) return rc;
    cg_zz_kernel<<<gr, CG_THREADS, 0, s>>>(p->cg_z, part_zz[0], g.nx, nc);
    cg_dir_kernel<<<gr, CG_THREADS, 0, s>>>(p->cg_p, p->cg_z, part_zz[0], part_zz[0], N, 1);
    TRON_CUDA(cudaMemsetAsync(x, 0, (size_t)nb * N * sizeof(float2), s));
    if (niter > 1) {
        const size_t window_stride = (size_t)g.prof_slide * g.nro * nc;
        cg_init_r_kernel<<<gr, CG_THREADS, 0, s>>>(p->cg_r, (const char *)d_in + (size_t)z0 * window_stride * in_esz, n,
                                                   window_stride, p->cfg.half_in);
    }
    TRON_CUDA(cudaGetLastError());
    p->last_launches += 3;
    for (int t = 0; t < niter; ++t) {
        {                                                /* v_b = A p_b, the whole batch per launch */
            FwdFftLaunch f;
            f.img = p->cg_p; f.tmp = p->d_tmp; f.grid = p->d_grid; f.deapod = p->deapod_fwd;
            f.nch = p->nch; f.nc_total = nc; f.ch0 = g.coil_begin; f.half_in = 0; f.nimg = nb;
            rc = launch_fwd_fft(p->fft, f, s);
            if (rc) return rc;
            DegridLaunch d;
            const int per_slice = p->tabs.ntab > 1;
            d.samples = p->cg_v; d.grid = p->d_grid;
            d.cs = p->tabs.cs_lin + (size_t)(per_slice ? z0 : 0) * p->tabs.npe;
            d.cs_stride = per_slice ? p->tabs.npe : 0;
            d.n = g.nxos; d.nro = g.nro; d.npe = g.npe1work;
            d.nc_total = nc; d.ch0 = g.coil_begin; d.nch = p->nch;
            d.kb = p->kb; d.half_out = 0; d.nimg = nb;
            rc = launch_degrid(d, s);
            if (rc) return rc;
        }
        p->last_launches += 3;
        const int last = t == niter - 1;
        cg_vwv_kernel<<<gr, CG_THREADS, 0, s>>>(p->cg_v, part_vwv, (size_t)g.npe1work * g.nro, g.nro, nc, wa, wb);
        cg_step_kernel<<<gr, CG_THREADS, 0, s>>>(x, p->cg_p, p->cg_r, p->cg_v, part_zz[t & 1], part_vwv, N, n,
                                                 g.nro, nc, sc, !last);
        TRON_CUDA(cudaGetLastError());
        p->last_launches += 2;
        if (last) break;
        /* the residual buffer is batch local: slice z0 + b sits at index b, windows do not overlap */
        const char *r0 = (const char *)p->cg_r - (ptrdiff_t)z0 * (ptrdiff_t)(n * sizeof(float2));
        rc = cg_apply_adjoint(p, r0, 0, g.npe1work, p->cg_z, z0, nb, p->deapod_fwd, s);
        if (rc) return rc;
        cg_zz_kernel<<<gr, CG_THREADS, 0, s>>>(p->cg_z, part_zz[(t + 1) & 1], g.nx, nc);
        cg_dir_kernel<<<gr, CG_THREADS, 0, s>>>(p->cg_p, p->cg_z, part_zz[(t + 1) & 1], part_zz[t & 1], N, 0);
        TRON_CUDA(cudaGetLastError());
        p->last_launches += 2;
    }
    return TRON_OK;
}

size_t cg_part_doubles(int batch) { return 3 * (size_t)batch * CG_PARTS; }

} // namespace tronb
