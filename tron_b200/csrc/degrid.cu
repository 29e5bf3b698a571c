/*
 * degrid.cu -- forward interpolation ("degridding"): Cartesian grid -> radial samples.
 *
 * Replaces degridradial2d (/root/reference/src/tron.cu:540-577).
 *
 * One thread per (sample, channel chunk); taps accumulate in registers and the
 * sample is written once (the reference read-modify-writes global memory per
 * tap per channel, tron.cu:572-573).  The Kaiser-Bessel factor along the
 * column axis is evaluated once per tap column and reused across rows.
 *
 * Index map, as the reference's SASS evaluates it:
 *   R  = fma(float(ro), rcp.approx(float(nro)), -0.5)
 *   nR = R * float(n)
 *   X  = fma(sinT, nR, c),  Y = fma(cosT, nR, c),  c = (n+1)/2 (integer division)
 *   xu = ceil(X - W) .. while float(xu) <= X + W;  live iff |float(xu) - X| < W
 *   cell = ((xu + n) % n, (yu + n) % n), X walks rows, Y walks columns.
 * No output scaling (the reference applies none in this direction).
 *
 * The grid is planar [ch][row][col]; samples are channel-interleaved.
 */
#include "tron_internal.h"

namespace tronb {

/* NT = number of column taps held in registers, >= floor(2W)+1 */

template <int CH, bool HALF>
__device__ __forceinline__ void store_sample(void *samples, size_t idx, const float2 (&acc)[CH])
{
    if (!HALF) {
        float2 *p = (float2 *)samples + idx;
        if (CH % 2 == 0) {
#pragma unroll
            for (int i = 0; i < CH / 2; ++i)
                ((float4 *)p)[i] = make_float4(acc[2 * i].x, acc[2 * i].y, acc[2 * i + 1].x, acc[2 * i + 1].y);
        } else {
#pragma unroll
            for (int i = 0; i < CH; ++i) p[i] = acc[i];
        }
    } else {
        __half2 *p = (__half2 *)samples + idx;
#pragma unroll
        for (int i = 0; i < CH; ++i) p[i] = __float22half2_rn(acc[i]);
    }
}

/* (u + n) % n of tron.cu:569-570 for -n <= u < 2n (taps reach at most W < n cells past the grid):
 * two selects instead of an integer division */
__device__ __forceinline__ int wrap_cell(int u, int n)
{
    u += u < 0 ? n : 0;
    return u - (u >= n ? n : 0);
}

template <int CH, bool HALF, int NT>
__global__ void __launch_bounds__(256)
degrid_gather_kernel(const DegridLaunch d)
{
    const int n = d.n;
    const size_t nsamp = (size_t)d.nro * d.npe;
    const int nchunk = d.nch / CH;
    const size_t total = nsamp * nchunk;
    const float W = d.kb.W;
    const float c0 = (float)((n + 1) / 2);
    const float inv_nro = rcp_approx((float)d.nro);
    const size_t plane = (size_t)n * n;
    /* blockIdx.y = grid of a batch (CGNR): its own plane set, spoke table and sample block */
    const float2 *grid_b = d.grid + (size_t)blockIdx.y * d.nch * plane;
    const float2 *cs_b = d.cs + (size_t)blockIdx.y * d.cs_stride;
    const size_t samp_b = (size_t)blockIdx.y * nsamp * d.nc_total;

    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (size_t)gridDim.x * blockDim.x) {
        /* sample fastest: the lanes of a warp are consecutive samples of one spoke, so their taps fall
         * on neighbouring cells of the same channel plane (chunk-fastest made every load touch 32 planes) */
        const int chunk = (int)(t / nsamp);
        const size_t id = t - (size_t)chunk * nsamp;
        const int pe = (int)(id / d.nro), ro = (int)(id - (size_t)pe * d.nro);
        const float2 cs = __ldg(cs_b + pe);
        const float R = fma_ftz((float)ro, inv_nro, -0.5f);
        const float nR = mul_ftz(R, (float)n);
        const float X = fma_ftz(cs.y, nR, c0);          /* rows:    sin */
        const float Y = fma_ftz(cs.x, nR, c0);          /* columns: cos */

        /* column taps: weights and wrapped indices, once */
        float wy[NT]; int jy[NT];
        const int yu0 = (int)ceilf(Y - W);
        const float ytop = Y + W;
#pragma unroll
        for (int k = 0; k < NT; ++k) {
            int yu = yu0 + k;
            float dy = (float)yu - Y;
            bool in = ((float)yu <= ytop) && (fabsf(dy) < W);
            wy[k] = in ? kb_weight(dy, d.kb) : 0.f;
            jy[k] = wrap_cell(yu, n);
        }
        const int ntapy = min(NT, (int)floorf(ytop) - yu0 + 1);

        float2 acc[CH];
#pragma unroll
        for (int i = 0; i < CH; ++i) acc[i] = make_float2(0.f, 0.f);
        const float2 *g0 = grid_b + (size_t)chunk * CH * plane;

        const float xtop = X + W;
        for (int xu = (int)ceilf(X - W); (float)xu <= xtop; ++xu) {
            float dx = (float)xu - X;
            if (!(fabsf(dx) < W)) continue;
            const float wx = kb_weight(dx, d.kb);
            const float2 *row = g0 + (size_t)wrap_cell(xu, n) * n;
#pragma unroll
            for (int k = 0; k < NT; ++k) {
                if (k < ntapy) {
                    const float w = wx * wy[k];
#pragma unroll
                    for (int i = 0; i < CH; ++i) {
                        float2 v = __ldg(row + (size_t)i * plane + jy[k]);
                        acc[i].x = fmaf(w, v.x, acc[i].x);
                        acc[i].y = fmaf(w, v.y, acc[i].y);
                    }
                }
            }
        }
        store_sample<CH, HALF>(d.samples, samp_b + id * d.nc_total + d.ch0 + (size_t)chunk * CH, acc);
    }
}

template <int CH, int NT>
static int launch_degrid_nt(const DegridLaunch &d, cudaStream_t s)
{
    size_t total = (size_t)d.nro * d.npe * (d.nch / CH);
    int bx = (int)((total + 255) / 256);
    if (bx > 148 * 64) bx = 148 * 64;
    dim3 blocks(bx, d.nimg > 0 ? d.nimg : 1);
    if (d.half_out) degrid_gather_kernel<CH, true, NT><<<blocks, 256, 0, s>>>(d);
    else            degrid_gather_kernel<CH, false, NT><<<blocks, 256, 0, s>>>(d);
    TRON_CUDA(cudaGetLastError());
    return 0;
}

template <int CH>
static int launch_degrid_ch(const DegridLaunch &d, cudaStream_t s)
{
    int taps = (int)floorf(2.f * d.kb.W) + 1;
    if (taps <= 5) return launch_degrid_nt<CH, 5>(d, s);
    if (taps <= 7) return launch_degrid_nt<CH, 7>(d, s);
    if (taps <= 9) return launch_degrid_nt<CH, 9>(d, s);
    if (taps <= 13) return launch_degrid_nt<CH, 13>(d, s);
    return launch_degrid_nt<CH, 16>(d, s);
}

int launch_degrid(const DegridLaunch &d, cudaStream_t s)
{
    if (d.kb.W > 7.5f) { set_error("kernel width %.2f too large for the degridding kernel (max 7.5)", d.kb.W); return TRON_EUNSUPPORTED; }
    size_t esz = d.half_out ? 4 : 8;
    bool aligned = (((uintptr_t)d.samples) % (2 * esz) == 0) && (d.nc_total % 2 == 0) && (d.ch0 % 2 == 0);
    if (!aligned || d.nch % 2) return launch_degrid_ch<1>(d, s);
    if (d.nch % 4 == 0) return launch_degrid_ch<4>(d, s);
    return launch_degrid_ch<2>(d, s);
}

} // namespace tronb
