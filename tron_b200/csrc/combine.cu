/*
 * combine.cu -- coil combination of per-coil images held channel-interleaved
 * ([slice][row][col][nc], the layout tron_nufft_adj_radial2d returns).
 *
 * Replaces
 *   coilcombinesos    /root/reference/src/tron.cu:255-268  (call: tron.cu:764)
 *   coilcombinewalsh  tron.cu:270-302 with powit tron.cu:222-253 (call: tron.cu:766, commented
 *                     out in the reference; reachable here with tron_config.coil_combine = 1)
 *
 * The plain adjoint path never comes through here: its root sum of squares is fused into the
 * last FFT pass (fft.cu).  These kernels serve the paths that need every coil image of a slice
 * at once -- the adaptive (Walsh) combine and the output of the CGNR iteration (cgnr.cu).
 *
 * Walsh combine, per pixel: A = sum over the (2 npatch + 1)^2 patch (clipped at the border) of
 * z z^H; five power iterations from x = (1,..,1), normalising with a multiplication by the
 * float reciprocal of |y| (float2math.h:24-28); img = x^H z.  Two kernels:
 *   walsh_small_kernel<NC>  nc <= 8: one thread per pixel, the Hermitian matrix (upper triangle,
 *       accumulated px-outer / py-inner like tron.cu:284-290) and the iteration live in registers;
 *   walsh_wide_kernel<CPL>  nc > 8: one warp per pixel, lanes = channels, matrix free:
 *       A x = sum_q z_q (z_q^H x), one coalesced load and one complex shuffle reduction per patch
 *       pixel and iteration -- 2 nc (2p+1)^2 complex FMAs per iteration instead of nc^2 plus the
 *       nc^2 (2p+1)^2 of forming A, and no nc^2 storage (the reference keeps A in local memory
 *       and is limited to MAXCHAN = 6).
 * A patch of zeros (e.g. the image row/column the CGNR iteration clears) gives 0 here; the reference would
 * produce 0 * (1/0) = NaN there.
 */
#include "tron_internal.h"

namespace tronb {

__device__ __forceinline__ float2 cmul_(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
/* a * conj(b) */
__device__ __forceinline__ float2 cmulc_(float2 a, float2 b) { return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }

template <bool HALF>
__device__ __forceinline__ void store_px(void *out, size_t i, float2 v)
{
    if (HALF) ((__half2 *)out)[i] = __float22half2_rn(v);
    else ((float2 *)out)[i] = v;
}

/* ---------------------------------------------------------------------- */
/* root sum of squares and friends from interleaved coil images            */
/* ---------------------------------------------------------------------- */
/* mode 0: (sqrt(sum_c |z_c|^2), 0), sequential over c (tron.cu:259-264); 1: single channel passes
 * through as complex (tron.cu:265-266); 2: copy of the per-coil images; 3: the sum itself, float32 */
__global__ void coil_combine_kernel(void *__restrict__ out, const float2 *__restrict__ coil, size_t npix, int nc,
                                    int mode, int half_out)
{
    for (size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x; id < npix; id += (size_t)gridDim.x * blockDim.x) {
        if (mode == 2) {
            for (int c = 0; c < nc; ++c) {
                float2 z = coil[id * nc + c];
                if (half_out) store_px<true>(out, id * nc + c, z); else store_px<false>(out, id * nc + c, z);
            }
            continue;
        }
        if (mode == 1) {
            float2 z = coil[id];
            if (half_out) store_px<true>(out, id, z); else store_px<false>(out, id, z);
            continue;
        }
        float val = 0.f;
        for (int c = 0; c < nc; ++c) { float2 z = coil[id * nc + c]; val += z.x * z.x + z.y * z.y; }
        if (mode == 3) ((float *)out)[id] = val;
        else if (half_out) store_px<true>(out, id, make_float2(sqrtf(val), 0.f));
        else store_px<false>(out, id, make_float2(sqrtf(val), 0.f));
    }
}

int launch_coil_combine(void *out, const float2 *coil, size_t npix, int nc, int mode, int half_out, cudaStream_t s)
{
    size_t blocks = (npix + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    coil_combine_kernel<<<(unsigned)blocks, 256, 0, s>>>(out, coil, npix, nc, mode, half_out);
    TRON_CUDA(cudaGetLastError());
    return 0;
}

/* ---------------------------------------------------------------------- */
/* Walsh, nc <= 8: registers                                               */
/* ---------------------------------------------------------------------- */
template <int NC>
__device__ __forceinline__ void load_px(float2 (&z)[NC], const float2 *p)
{
#pragma unroll
    for (int i = 0; i < NC / 2; ++i) {
        float4 q = __ldg((const float4 *)p + i);
        z[2 * i] = make_float2(q.x, q.y); z[2 * i + 1] = make_float2(q.z, q.w);
    }
}

/* acc += w * v on both halves: one packed FP32x2 FMA (FFMA2 on sm_100a) */
__device__ __forceinline__ void cfma2(float2 &acc, float w, float2 v)
{
    unsigned long long a = *reinterpret_cast<unsigned long long *>(&acc);
    float2 ww = make_float2(w, w);
    asm("fma.rn.f32x2 %0, %1, %2, %0;"
        : "+l"(a)
        : "l"(*reinterpret_cast<unsigned long long *>(&ww)), "l"(*reinterpret_cast<unsigned long long *>(&v)));
    acc = *reinterpret_cast<float2 *>(&a);
}

/* Complex products as two packed FMAs each: a * b = a b.x + (-a.y, a.x) b.y and
 * a * conj(b) = a b.x + (a.y, -a.x) b.y; the rotated operand is formed once per vector element. */
template <int NC, bool HALF>
__global__ void __launch_bounds__(128)
walsh_small_kernel(void *__restrict__ out, const float2 *__restrict__ coil, int nimg, int npatch)
{
    constexpr int NT = NC * (NC + 1) / 2;
    const size_t npix = (size_t)nimg * nimg;
    const float2 *src = coil + (size_t)blockIdx.y * npix * NC;
    for (int id = blockIdx.x * blockDim.x + threadIdx.x; id < (int)npix; id += gridDim.x * blockDim.x) {
        const int x = id / nimg, y = id % nimg;
        float2 A[NT];                                   /* A[j][k], j <= k, at j*NC - j*(j-1)/2 + (k-j) */
#pragma unroll
        for (int k = 0; k < NT; ++k) A[k] = make_float2(0.f, 0.f);
        const int x0 = max(0, x - npatch), x1 = min(nimg - 1, x + npatch);
        const int y0 = max(0, y - npatch), y1 = min(nimg - 1, y + npatch);
        for (int px = x0; px <= x1; ++px)
            for (int py = y0; py <= y1; ++py) {         /* px outer, py inner: tron.cu:284-285 */
                float2 z[NC], zr[NC];
                load_px<NC>(z, src + ((size_t)px * nimg + py) * NC);
#pragma unroll
                for (int j = 0; j < NC; ++j) zr[j] = make_float2(z[j].y, -z[j].x);
                int t = 0;
#pragma unroll
                for (int j = 0; j < NC; ++j)
#pragma unroll
                    for (int k = j; k < NC; ++k, ++t) {
                        if (k == j) A[t].x = fmaf(z[j].x, z[j].x, fmaf(z[j].y, z[j].y, A[t].x));
                        else { cfma2(A[t], z[k].x, z[j]); cfma2(A[t], z[k].y, zr[j]); }   /* z_j conj(z_k) */
                    }
            }
        float2 xv[NC], yv[NC];
#pragma unroll
        for (int k = 0; k < NC; ++k) xv[k] = make_float2(1.f, 0.f);
#pragma unroll 1
        for (int it = 0; it < 5; ++it) {                /* tron.cu:291 */
            float2 xr[NC];
#pragma unroll
            for (int k = 0; k < NC; ++k) xr[k] = make_float2(-xv[k].y, xv[k].x);
#pragma unroll
            for (int j = 0; j < NC; ++j) {
                float2 acc = make_float2(0.f, 0.f);
#pragma unroll
                for (int k = 0; k < NC; ++k) {          /* y_j += A[j][k] x_k, A[j][k] = conj(A[k][j]) for j > k */
                    const int lo = j < k ? j : k, hi = j < k ? k : j;
                    const float2 a = A[lo * NC - lo * (lo - 1) / 2 + (hi - lo)];
                    cfma2(acc, a.x, xv[k]);
                    if (j < k) cfma2(acc, a.y, xr[k]);
                    else if (j > k) cfma2(acc, -a.y, xr[k]);
                }
                yv[j] = acc;
            }
            float nsq = 0.f;
#pragma unroll
            for (int k = 0; k < NC; ++k) nsq = fmaf(yv[k].x, yv[k].x, fmaf(yv[k].y, yv[k].y, nsq));
            const float inv = nsq > 0.f ? 1.0f / sqrtf(nsq) : 0.f;
#pragma unroll
            for (int k = 0; k < NC; ++k) xv[k] = make_float2(yv[k].x * inv, yv[k].y * inv);
        }
        float2 z[NC];
        load_px<NC>(z, src + (size_t)id * NC);
        float2 o = make_float2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < NC; ++c) {                  /* conj(x_c) z_c, tron.cu:294 */
            cfma2(o, xv[c].x, z[c]);
            cfma2(o, xv[c].y, make_float2(z[c].y, -z[c].x));
        }
        store_px<HALF>(out, (size_t)blockIdx.y * npix + id, o);
    }
}

/* ---------------------------------------------------------------------- */
/* Walsh, nc > 8: one warp per pixel, lanes = channels, matrix free        */
/* ---------------------------------------------------------------------- */
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int CPL, bool HALF>
__global__ void __launch_bounds__(256)
walsh_wide_kernel(void *__restrict__ out, const float2 *__restrict__ coil, int nimg, int nc, int npatch)
{
    constexpr int WCH = 1;   /* measured: the kernel is bound by SHFL throughput (10 per patch pixel), not by their latency;
                                reducing 4 pixels per round was no faster (nc = 32) or slower (nc = 64) */
    const int lane = threadIdx.x & 31;
    const size_t npix = (size_t)nimg * nimg;
    const float2 *src = coil + (size_t)blockIdx.y * npix * nc;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int id = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; id < (int)npix; id += warps) {
        const int x = id / nimg, y = id % nimg;
        const int x0 = max(0, x - npatch), x1 = min(nimg - 1, x + npatch);
        const int y0 = max(0, y - npatch), y1 = min(nimg - 1, y + npatch);
        float2 xv[CPL], yv[CPL];
#pragma unroll
        for (int i = 0; i < CPL; ++i) xv[i] = lane + 32 * i < nc ? make_float2(1.f, 0.f) : make_float2(0.f, 0.f);
#pragma unroll 1
        for (int it = 0; it < 5; ++it) {
#pragma unroll
            for (int i = 0; i < CPL; ++i) yv[i] = make_float2(0.f, 0.f);
            /* a patch row in chunks of WCH pixels whose dot products are reduced together */
            for (int px = x0; px <= x1; ++px)
                for (int pyc = y0; pyc <= y1; pyc += WCH) {
                    const float2 *p = src + ((size_t)px * nimg + pyc) * nc;
                    float2 z[WCH][CPL], d[WCH];
#pragma unroll
                    for (int j = 0; j < WCH; ++j) {
                        const bool okj = pyc + j <= y1;
                        d[j] = make_float2(0.f, 0.f);    /* z_q^H x */
#pragma unroll
                        for (int i = 0; i < CPL; ++i) {
                            z[j][i] = (okj && lane + 32 * i < nc) ? __ldg(p + (size_t)j * nc + lane + 32 * i) : make_float2(0.f, 0.f);
                            float2 m = cmulc_(xv[i], z[j][i]);
                            d[j].x += m.x; d[j].y += m.y;
                        }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                        for (int j = 0; j < WCH; ++j) {
                            d[j].x += __shfl_xor_sync(0xffffffffu, d[j].x, o);
                            d[j].y += __shfl_xor_sync(0xffffffffu, d[j].y, o);
                        }
#pragma unroll
                    for (int j = 0; j < WCH; ++j)
#pragma unroll
                        for (int i = 0; i < CPL; ++i) {
                            float2 m = cmul_(z[j][i], d[j]);
                            yv[i].x += m.x; yv[i].y += m.y;
                        }
                }
            float nsq = 0.f;
#pragma unroll
            for (int i = 0; i < CPL; ++i) nsq += yv[i].x * yv[i].x + yv[i].y * yv[i].y;
            nsq = warp_sum(nsq);
            const float inv = nsq > 0.f ? 1.0f / sqrtf(nsq) : 0.f;
#pragma unroll
            for (int i = 0; i < CPL; ++i) xv[i] = make_float2(yv[i].x * inv, yv[i].y * inv);
        }
        const float2 *p = src + (size_t)id * nc;
        float2 o = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
            float2 z = lane + 32 * i < nc ? __ldg(p + lane + 32 * i) : make_float2(0.f, 0.f);
            float2 m = cmulc_(z, xv[i]);
            o.x += m.x; o.y += m.y;
        }
        o.x = warp_sum(o.x); o.y = warp_sum(o.y);
        if (lane == 0) store_px<HALF>(out, (size_t)blockIdx.y * npix + id, o);
    }
}

template <bool HALF>
static int launch_walsh_h(void *out, const float2 *coil, int nimg, int nc, int npatch, int nslices, cudaStream_t s)
{
    const size_t npix = (size_t)nimg * nimg;
    const bool aligned = ((uintptr_t)coil % 16) == 0;
    if (nc <= 8 && nc % 2 == 0 && aligned) {
        dim3 grid((unsigned)((npix + 127) / 128), nslices);
        switch (nc) {
        case 2: walsh_small_kernel<2, HALF><<<grid, 128, 0, s>>>(out, coil, nimg, npatch); break;
        case 4: walsh_small_kernel<4, HALF><<<grid, 128, 0, s>>>(out, coil, nimg, npatch); break;
        case 6: walsh_small_kernel<6, HALF><<<grid, 128, 0, s>>>(out, coil, nimg, npatch); break;
        default: walsh_small_kernel<8, HALF><<<grid, 128, 0, s>>>(out, coil, nimg, npatch); break;
        }
    } else {
        size_t blocks = (npix + 7) / 8;                  /* 8 warps per block, one pixel per warp and trip */
        if (blocks > 148 * 8) blocks = 148 * 8;
        dim3 grid((unsigned)blocks, nslices);
        const int cpl = (nc + 31) / 32;
        if (cpl == 1) walsh_wide_kernel<1, HALF><<<grid, 256, 0, s>>>(out, coil, nimg, nc, npatch);
        else if (cpl == 2) walsh_wide_kernel<2, HALF><<<grid, 256, 0, s>>>(out, coil, nimg, nc, npatch);
        else if (cpl <= 4) walsh_wide_kernel<4, HALF><<<grid, 256, 0, s>>>(out, coil, nimg, nc, npatch);
        else { set_error("Walsh combine supports at most 128 channels (nc = %d)", nc); return TRON_EUNSUPPORTED; }
    }
    TRON_CUDA(cudaGetLastError());
    return 0;
}

/* coil [nslices][nimg][nimg][nc] complex64 -> out [nslices][nimg][nimg] complex64 (or complex-half) */
int launch_walsh(void *out, const float2 *coil, int nimg, int nc, int npatch, int nslices, int half_out, cudaStream_t s)
{
    if (nslices <= 0) return 0;
    if (nc == 1) return launch_coil_combine(out, coil, (size_t)nimg * nimg * nslices, 1, 1, half_out, s);   /* tron.cu:277-278 */
    return half_out ? launch_walsh_h<true>(out, coil, nimg, nc, npatch, nslices, s)
                    : launch_walsh_h<false>(out, coil, nimg, nc, npatch, nslices, s);
}

} // namespace tronb

using namespace tronb;

extern "C" int tron_coilcombine_walsh_device(void *d_img, const void *d_coilimg, int nimg, int nchan, int npatch,
                                             int nslices, void *stream)
{
    if (!d_img || !d_coilimg || nimg < 1 || nchan < 1 || npatch < 0 || nslices < 1) { set_error("bad argument"); return TRON_EINVAL; }
    return launch_walsh(d_img, (const float2 *)d_coilimg, nimg, nchan, npatch, nslices, 0, (cudaStream_t)stream);
}

extern "C" int tron_coilcombine_sos_device(void *d_img, const void *d_coilimg, int nimg, int nchan, int nslices, void *stream)
{
    if (!d_img || !d_coilimg || nimg < 1 || nchan < 1 || nslices < 1) { set_error("bad argument"); return TRON_EINVAL; }
    return launch_coil_combine(d_img, (const float2 *)d_coilimg, (size_t)nimg * nimg * nslices, nchan,
                               nchan > 1 ? 0 : 1, 0, (cudaStream_t)stream);
}
