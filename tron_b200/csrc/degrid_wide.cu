/*
 * degrid_wide.cu -- degridding for many receive channels (nc a multiple of 32): lanes = channels.
 *
 * Same operator and tap set as degrid.cu (reference: degridradial2d,
 * /root/reference/src/tron.cu:540-577).  A warp owns four consecutive samples of one spoke; their
 * tap windows overlap almost completely (neighbouring samples are one cell apart), so the warp
 * walks the union window once:
 *
 *   A  lanes = rows / columns of the union window: the Kaiser-Bessel factor of every row and
 *      every column for each of the four samples (zero outside that sample's own support, decided
 *      by the reference predicate |xu - X| < W on the reference's coordinates, refmath.cuh);
 *   B  lanes = channels: every grid cell of the union window is loaded once, as one coalesced
 *      request from the channel-interleaved grid, and feeds the four samples' accumulators with
 *      packed FFMA2s.
 *
 * With -k 6 (13 x 13 taps) this is FP32 bound; the point of the blocking is
 * that a cell costs one load + four FMAs instead of four loads + four FMAs, all requests are full
 * lines, and the per-tap weight work is amortised over the channels.
 *
 * Input grid: channel-interleaved g[(row*n + col)*nch + ch] (fwd FFT output transposed by
 * planar_to_interleaved_kernel); output: samples[(pe*nro + ro)*nc_total + ch0 + ch].
 */
#include "tron_internal.h"

namespace tronb {

#define DW_S 4            /* samples per warp */
#define DW_MAXU 20        /* union window side: floor(2W)+1 + DW_S-1 + slack <= 20  (W <= 7.5) */

struct __align__(16) DwWeights {
    float4 wx[DW_MAXU];   /* row factor of samples 0..3 */
    float4 wy[DW_MAXU];   /* column factor */
    float4 wp[4][DW_MAXU];/* wx[i] * wy[j] of the rows in flight (one per sub-warp, see LPC): formed once per row by lanes
                             j < nuy, so the channel loop reads the finished tap weight instead of multiplying per cell */
    int coff[DW_MAXU + 4];/* element offset of every column of the union window: the periodic wrap is applied once
                             per column here instead of an integer modulo per cell (ncu: 45 -> 25 instructions per
                             cell, cfg5 forward 18.4 -> 13.6 ms); padded by repeating the last column */
};

__device__ __forceinline__ void ffma2d(float2 &acc, float w, float2 v)
{
    unsigned long long a = *reinterpret_cast<unsigned long long *>(&acc);
    float2 ww = make_float2(w, w);
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;"
                 : "+l"(a)
                 : "l"(*reinterpret_cast<unsigned long long *>(&ww)), "l"(*reinterpret_cast<unsigned long long *>(&v)));
    acc = *reinterpret_cast<float2 *>(&a);
}

__device__ __forceinline__ float2 ldg2v(const float2 *p)
{
    float2 v;
    asm volatile("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}

/* LPC = lanes per cell = channels fetched by one request.  32 (or 64 channels with NCHUNK = 2): the warp walks the
 * rows of the union window one by one.  16 / 8 (coil shards of a many-coil job, cfg5 on 4 / 8 GPUs): the warp's
 * 2 / 4 sub-warps take 2 / 4 rows at a time -- every lane still owns a channel, every request is still a whole
 * 128- / 64-byte piece -- and their partial sums are added by shuffles at the end. */
template <int NCHUNK, bool HALF, int LPC>
__global__ void __launch_bounds__(256)
degrid_wide_kernel(const DegridLaunch d, const float2 *__restrict__ gi /* interleaved grid */)
{
    constexpr int RPI = 32 / LPC;                    /* rows in flight per warp */
    __shared__ DwWeights sw[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    DwWeights &S = sw[warp];
    const int n = d.n, nch = d.nch;
    const int groups_per_spoke = (d.nro + DW_S - 1) / DW_S;
    const long long ngroups = (long long)groups_per_spoke * d.npe;
    const float W = d.kb.W;
    const float c0 = (float)((n + 1) / 2);
    const float inv_nro = rcp_approx((float)d.nro);
    const int chan0 = blockIdx.y * 32 * NCHUNK;
    const int clane = lane % LPC, rsub = lane / LPC; /* channel lane, sub-warp (row) */

    /* Work order.  The 8 warps of a block take 8 ADJACENT SPOKES at the same radial position, and consecutive
     * blocks walk outwards along that bundle: the few hundred blocks in flight then cover one thin bundle of
     * spokes whose tap windows overlap (13 cells wide at -k 6, neighbouring spokes <= 1.6 cells apart), i.e. a
     * footprint of ~20 MB that stays in the 126 MB L2, and the next bundle re-uses its inner part.  Walking
     * along one spoke per warp instead (round 1) put ~560 MB of windows in flight at once: ncu measured 31.9 GB
     * of DRAM reads for a 2.1 GB grid (profiles/r02_ncu_cfg5_wide.txt). */
    const long long nwork = (long long)((d.npe + 7) / 8) * groups_per_spoke;
    (void)ngroups;
    for (long long wb = blockIdx.x; wb < nwork; wb += gridDim.x) {
        const int pe = (int)(wb / groups_per_spoke) * 8 + warp;
        const int ro0 = (int)(wb % groups_per_spoke) * DW_S;
        if (pe >= d.npe) continue;
        const float2 cs = __ldg(d.cs + pe);
        /* coordinates of the four samples, exactly as tron.cu:554-561 compiles */
        float X[DW_S], Y[DW_S];
        int xlo = 1 << 30, xhi = -(1 << 30), ylo = 1 << 30, yhi = -(1 << 30);
#pragma unroll
        for (int s = 0; s < DW_S; ++s) {
            const float R = fma_ftz((float)(ro0 + s), inv_nro, -0.5f);
            const float nR = mul_ftz(R, (float)n);
            X[s] = fma_ftz(cs.y, nR, c0);            /* rows:    sin */
            Y[s] = fma_ftz(cs.x, nR, c0);            /* columns: cos */
            if (ro0 + s < d.nro) {
                xlo = min(xlo, (int)ceilf(X[s] - W)); xhi = max(xhi, (int)floorf(X[s] + W));
                ylo = min(ylo, (int)ceilf(Y[s] - W)); yhi = max(yhi, (int)floorf(Y[s] + W));
            }
        }
        const int nux = min(xhi - xlo + 1, DW_MAXU), nuy = min(yhi - ylo + 1, DW_MAXU);

        /* phase A: lanes 0..nux-1 rows, the others (offset 16 would not cover 20) -> two passes */
        __syncwarp();
        for (int i = lane; i < nux + nuy; i += 32) {
            const bool isrow = i < nux;
            const int u = isrow ? xlo + i : ylo + (i - nux);
            float w4[DW_S];
#pragma unroll
            for (int s = 0; s < DW_S; ++s) {
                const float dd = (float)u - (isrow ? X[s] : Y[s]);
                const bool live = (ro0 + s < d.nro) && fabsf(dd) < W;     /* tron.cu:343 via gridkernel */
                w4[s] = live ? kb_weight(dd, d.kb) : 0.f;
            }
            if (isrow) S.wx[i] = make_float4(w4[0], w4[1], w4[2], w4[3]);
            else {
                S.wy[i - nux] = make_float4(w4[0], w4[1], w4[2], w4[3]);
                const int off = ((u + n) % n) * nch;                           /* periodic, tron.cu:570 */
                S.coff[i - nux] = off;
                if (i - nux == nuy - 1) { S.coff[nuy] = off; S.coff[nuy + 1] = off; S.coff[nuy + 2] = off; }
            }
        }
        __syncwarp();

        /* phase B: lanes = channels */
        float2 acc[NCHUNK][DW_S];
#pragma unroll
        for (int c = 0; c < NCHUNK; ++c)
#pragma unroll
            for (int s = 0; s < DW_S; ++s) acc[c][s] = make_float2(0.f, 0.f);
        for (int i0 = 0; i0 < nux; i0 += RPI) {
            __syncwarp();
#pragma unroll
            for (int rr = 0; rr < RPI; ++rr) {
                const float4 a = i0 + rr < nux ? S.wx[i0 + rr] : make_float4(0.f, 0.f, 0.f, 0.f);
                if (lane < nuy) {
                    const float4 b = S.wy[lane];
                    S.wp[rr][lane] = make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
                }
            }
            __syncwarp();
            const int i = min(i0 + rsub, nux - 1);                          /* (rows past the window: zero weights) */
            const int row = (xlo + i + n) % n;                              /* periodic, tron.cu:569 */
            const float2 *grow = gi + ((size_t)row * n) * nch + chan0 + clane;
            for (int j0 = 0; j0 < nuy; j0 += 4) {
                float2 v[4][NCHUNK];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int co = S.coff[j0 + jj];
#pragma unroll
                    for (int c = 0; c < NCHUNK; ++c) v[jj][c] = ldg2v(grow + co + c * 32);
                }
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    if (j0 + jj < nuy) {
                        const float4 b = S.wp[rsub][j0 + jj];
                        const float w0 = b.x, w1 = b.y, w2 = b.z, w3 = b.w;
#pragma unroll
                        for (int c = 0; c < NCHUNK; ++c) {
                            ffma2d(acc[c][0], w0, v[jj][c]); ffma2d(acc[c][1], w1, v[jj][c]);
                            ffma2d(acc[c][2], w2, v[jj][c]); ffma2d(acc[c][3], w3, v[jj][c]);
                        }
                    }
                }
            }
        }
        if (RPI > 1) {                                                      /* add the sub-warps' partial sums */
#pragma unroll
            for (int c = 0; c < NCHUNK; ++c)
#pragma unroll
                for (int s = 0; s < DW_S; ++s)
#pragma unroll
                    for (int o = LPC; o < 32; o <<= 1) {
                        acc[c][s].x += __shfl_xor_sync(0xffffffffu, acc[c][s].x, o);
                        acc[c][s].y += __shfl_xor_sync(0xffffffffu, acc[c][s].y, o);
                    }
        }
#pragma unroll
        for (int s = 0; s < DW_S; ++s) {
            if (ro0 + s >= d.nro) continue;
            const size_t base = ((size_t)pe * d.nro + ro0 + s) * d.nc_total + d.ch0 + chan0 + clane;
#pragma unroll
            for (int c = 0; c < NCHUNK; ++c) {
                if (rsub != 0 || chan0 + c * 32 + clane >= nch) continue;
                if (HALF) ((__half2 *)d.samples)[base + c * 32] = __float22half2_rn(acc[c][s]);
                else ((float2 *)d.samples)[base + c * 32] = acc[c][s];
            }
        }
    }
}

/* planar [ch][cell] -> interleaved [cell][ch], 32 x 32 tiles through shared memory */
__global__ void __launch_bounds__(256)
planar_to_interleaved_kernel(float2 *__restrict__ dst, const float2 *__restrict__ src, int nch, size_t ncell)
{
    __shared__ float2 tile[32][33];
    const size_t cell0 = (size_t)blockIdx.x * 32;
    const int ch0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      /* 32 x 8 threads */
    for (int r = ty; r < 32; r += 8) {
        const int ch = ch0 + r;
        const size_t cell = cell0 + tx;
        tile[r][tx] = (ch < nch && cell < ncell) ? src[(size_t)ch * ncell + cell] : make_float2(0.f, 0.f);
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const size_t cell = cell0 + r;
        const int ch = ch0 + tx;
        if (ch < nch && cell < ncell) dst[cell * nch + ch] = tile[tx][r];
    }
}

bool degrid_wide_applicable(const DegridLaunch &d)
{
    /* 8 or 16 channels: the thread-per-sample kernel wins for narrow kernels; from W = 3 on (cfg5's coil shards,
     * -k 6: 18.0 ms for a 16-coil 2048^2 shard) the shared window does */
    const bool few = (d.nch == 8 || d.nch == 16) && d.kb.W >= 3.f;
    if (!few && (d.nch < 32 || d.nch % 32 != 0)) return false;
    if ((int)floorf(2.f * d.kb.W) + 1 + DW_S - 1 + 1 > DW_MAXU) return false;
    return true;
}

/* scratch: nch*n*n float2, receives the channel-interleaved copy of the planar grid */
int launch_degrid_wide(const DegridLaunch &d, float2 *scratch, cudaStream_t s)
{
    const size_t ncell = (size_t)d.n * d.n;
    dim3 tg((unsigned)((ncell + 31) / 32), (unsigned)((d.nch + 31) / 32));
    planar_to_interleaved_kernel<<<tg, 256, 0, s>>>(scratch, d.grid, d.nch, ncell);
    TRON_CUDA(cudaGetLastError());
    const long long nwork = (long long)((d.nro + DW_S - 1) / DW_S) * ((d.npe + 7) / 8);
    int bx = (int)(nwork < 148 * 32 ? nwork : 148 * 32);
    if (d.nch == 8) {
        if (d.half_out) degrid_wide_kernel<1, true, 8><<<bx, 256, 0, s>>>(d, scratch);
        else            degrid_wide_kernel<1, false, 8><<<bx, 256, 0, s>>>(d, scratch);
    } else if (d.nch == 16) {
        if (d.half_out) degrid_wide_kernel<1, true, 16><<<bx, 256, 0, s>>>(d, scratch);
        else            degrid_wide_kernel<1, false, 16><<<bx, 256, 0, s>>>(d, scratch);
    } else if (d.nch % 64 == 0) {
        dim3 grid(bx, d.nch / 64);
        if (d.half_out) degrid_wide_kernel<2, true, 32><<<grid, 256, 0, s>>>(d, scratch);
        else            degrid_wide_kernel<2, false, 32><<<grid, 256, 0, s>>>(d, scratch);
    } else {
        dim3 grid(bx, d.nch / 32);
        if (d.half_out) degrid_wide_kernel<1, true, 32><<<grid, 256, 0, s>>>(d, scratch);
        else            degrid_wide_kernel<1, false, 32><<<grid, 256, 0, s>>>(d, scratch);
    }
    TRON_CUDA(cudaGetLastError());
    return 0;
}

} // namespace tronb
