/*
 * degrid_wide.cu -- degridding for many receive channels (nc a multiple of 32, or 8 / 16 at wide
 * kernels): lanes = channels.
 *
 * Same operator and tap set as degrid.cu (reference: degridradial2d,
 * /root/reference/src/tron.cu:540-577).  A warp owns four consecutive samples of one spoke -- of
 * two neighbouring spokes when the angle order is linear -- whose tap windows overlap almost
 * completely (neighbouring samples are one cell apart), so the warp walks the union window once:
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
 * Two kernels: degrid_wide_kernel (a warp, or 2 / 4 sub-warps on different ROWS, per cell: 64 and 16
 * channels) and degrid_cols_kernel (sub-warps on adjacent COLUMNS of one row: 8 and 32 channels).
 * Both are bound by L1 wavefronts per FFMA2, which is what the blocking is chosen by (DESIGN.md
 * section 3.3).
 *
 * Input grid: channel-interleaved g[(row*n + col)*nch + ch] (fwd FFT output transposed by
 * planar_to_interleaved_kernel); output: samples[(pe*nro + ro)*nc_total + ch0 + ch].
 */
#include "tron_internal.h"
#include <stdlib.h>

namespace tronb {

#define DW_S 4            /* samples per warp */
#define DW_MAXU 20        /* union window side: floor(2W)+1 + DW_S-1 + slack <= 20  (W <= 7.5) */

/* P = spokes per warp (DW_S samples each) */
template <int P> struct __align__(16) DwWeights {
    float4 wx[P][DW_MAXU];   /* row factor of samples 0..3 of spoke p */
    float4 wy[P][DW_MAXU];   /* column factor */
    float4 wp[4][P][DW_MAXU];/* wx[i] * wy[j] of the rows in flight (one per sub-warp, see LPC): formed once per row by lanes
                                j < nuy, so the channel loop reads the finished tap weight instead of multiplying per cell */
    int coff[DW_MAXU + 8];   /* wrapped index of every column of the union window: the periodic wrap is applied once
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

/* the NCHUNK adjacent channels a lane holds of one grid cell: one 8- or 16-byte request */
template <int NCHUNK>
__device__ __forceinline__ void ldg_cell(float2 (&v)[NCHUNK], const char *p)
{
    if (NCHUNK == 2) {
        asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(v[0].x), "=f"(v[0].y), "=f"(v[NCHUNK - 1].x), "=f"(v[NCHUNK - 1].y) : "l"(p));
    } else {
        v[0] = ldg2v((const float2 *)p);
    }
}

template <int NCHUNK, bool HALF, int NS>
__device__ __forceinline__ void store_chan(void *samples, size_t base, const float2 (&acc)[NCHUNK][NS], int s)
{
    if (HALF) {
        if (NCHUNK == 2) {
            const __half2 a = __float22half2_rn(acc[0][s]), b = __float22half2_rn(acc[NCHUNK - 1][s]);
            uint2 u = make_uint2(*reinterpret_cast<const unsigned *>(&a), *reinterpret_cast<const unsigned *>(&b));
            *reinterpret_cast<uint2 *>((__half2 *)samples + base) = u;
        } else ((__half2 *)samples)[base] = __float22half2_rn(acc[0][s]);
    } else {
        if (NCHUNK == 2)
            *reinterpret_cast<float4 *>((float2 *)samples + base) =
                make_float4(acc[0][s].x, acc[0][s].y, acc[NCHUNK - 1][s].x, acc[NCHUNK - 1][s].y);
        else ((float2 *)samples)[base] = acc[0][s];
    }
}

/* (u + n) % n of tron.cu:569-570 for -n <= u < 2n (taps reach at most W < n cells past the grid) */
__device__ __forceinline__ int wrap_cell_w(int u, int n)
{
    u += u < 0 ? n : 0;
    return u - (u >= n ? n : 0);
}

/* four neighbouring columns of one grid row */
template <int NCHUNK>
__device__ __forceinline__ void load_quad(float2 (&v)[4][NCHUNK], const char *grow, const int *coff, unsigned cstride)
{
    const int4 co = *reinterpret_cast<const int4 *>(coff);
    ldg_cell<NCHUNK>(v[0], grow + (unsigned long long)(unsigned)co.x * cstride);
    ldg_cell<NCHUNK>(v[1], grow + (unsigned long long)(unsigned)co.y * cstride);
    ldg_cell<NCHUNK>(v[2], grow + (unsigned long long)(unsigned)co.z * cstride);
    ldg_cell<NCHUNK>(v[3], grow + (unsigned long long)(unsigned)co.w * cstride);
}

template <int NCHUNK, int P>
__device__ __forceinline__ void fma_quad(float2 (&acc)[NCHUNK][P * 4], const float2 (&v)[4][NCHUNK], const float4 *w /* [P][DW_MAXU] */)
{
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const float4 b = w[p * DW_MAXU + jj];
#pragma unroll
            for (int c = 0; c < NCHUNK; ++c) {
                ffma2d(acc[c][4 * p + 0], b.x, v[jj][c]); ffma2d(acc[c][4 * p + 1], b.y, v[jj][c]);
                ffma2d(acc[c][4 * p + 2], b.z, v[jj][c]); ffma2d(acc[c][4 * p + 3], b.w, v[jj][c]);
            }
        }
    }
}

/* LPC = lanes per cell = channels fetched by one request.  32 (or 64 channels with NCHUNK = 2): the warp walks the
 * rows of the union window one by one.  16 / 8 (coil shards of a many-coil job, cfg5 on 4 / 8 GPUs): the warp's
 * 2 / 4 sub-warps take 2 / 4 rows at a time -- every lane still owns a channel, every request is still a whole
 * 128- / 64-byte piece -- and their partial sums are added by shuffles at the end. */
template <int NCHUNK, bool HALF, int LPC, int P>
__global__ void __launch_bounds__(256)
degrid_wide_kernel(const DegridLaunch d, const float2 *__restrict__ gi /* interleaved grid */)
{
    constexpr int RPI = 32 / LPC;                    /* rows in flight per warp */
    constexpr int NS = P * DW_S;                     /* samples per warp: s = DW_S * spoke + position */
    __shared__ DwWeights<P> sw[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    DwWeights<P> &S = sw[warp];
    const int n = d.n, nch = d.nch;
    const int groups_per_spoke = (d.nro + DW_S - 1) / DW_S;
    const float W = d.kb.W;
    const float c0 = (float)((n + 1) / 2);
    const float inv_nro = rcp_approx((float)d.nro);
    const int chan0 = blockIdx.y * 32 * NCHUNK;
    const int clane = lane % LPC, rsub = lane / LPC; /* channel lane, sub-warp (row) */
    const unsigned cstride = (unsigned)nch * (unsigned)sizeof(float2);   /* bytes between neighbouring columns */

    /* Work order.  The 8 warps of a block take 8 P ADJACENT SPOKES at the same radial position, and consecutive
     * blocks walk outwards along that bundle: the few hundred blocks in flight then cover one thin bundle of
     * spokes whose tap windows overlap (13 cells wide at -k 6, neighbouring spokes <= 1.6 cells apart), i.e. a
     * footprint of ~20 MB that stays in the 126 MB L2, and the next bundle re-uses its inner part.  Walking
     * along one spoke per warp instead (round 1) put ~560 MB of windows in flight at once: ncu measured 31.9 GB
     * of DRAM reads for a 2.1 GB grid (profiles/r02_ncu_cfg5_wide.txt).
     * P = 2 (linear angle order: spokes pe, pe + 1 are neighbours): the warp's union window barely grows, a loaded
     * cell feeds 8 samples instead of 4 -- the kernel is bound by L1 wavefronts (ncu: l1tex 79 %, a 64-channel cell
     * is 4 wavefronts + 1 for its weights per 8 FFMA2), so cells per sample is what counts.  A pair whose union
     * window would not fit (spokes that are NOT neighbours) is walked spoke by spoke. */
    const long long nwork = (long long)((d.npe + 8 * P - 1) / (8 * P)) * groups_per_spoke;
    for (long long wb = blockIdx.x; wb < nwork; wb += gridDim.x) {
        const int pe0 = ((int)(wb / groups_per_spoke) * 8 + warp) * P;
        const int ro0 = (int)(wb % groups_per_spoke) * DW_S;
        if (pe0 >= d.npe) continue;
        /* coordinates of the samples, exactly as tron.cu:554-561 compiles */
        float X[NS], Y[NS];
        int bx0[P], bx1[P], by0[P], by1[P];
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const float2 cs = __ldg(d.cs + min(pe0 + p, d.npe - 1));
            bx0[p] = by0[p] = 1 << 30; bx1[p] = by1[p] = -(1 << 30);
#pragma unroll
            for (int k = 0; k < DW_S; ++k) {
                const int s = p * DW_S + k;
                const float R = fma_ftz((float)(ro0 + k), inv_nro, -0.5f);
                const float nR = mul_ftz(R, (float)n);
                X[s] = fma_ftz(cs.y, nR, c0);            /* rows:    sin */
                Y[s] = fma_ftz(cs.x, nR, c0);            /* columns: cos */
                if (ro0 + k < d.nro && pe0 + p < d.npe) {
                    bx0[p] = min(bx0[p], (int)ceilf(X[s] - W)); bx1[p] = max(bx1[p], (int)floorf(X[s] + W));
                    by0[p] = min(by0[p], (int)ceilf(Y[s] - W)); by1[p] = max(by1[p], (int)floorf(Y[s] + W));
                }
            }
        }
        /* one pass over the union window of all spokes if it fits, else one pass per spoke */
        int npass = 1;
        if (P > 1) {
            int x0 = bx0[0], x1 = bx1[0], y0 = by0[0], y1 = by1[0];
#pragma unroll
            for (int p = 1; p < P; ++p) { x0 = min(x0, bx0[p]); x1 = max(x1, bx1[p]); y0 = min(y0, by0[p]); y1 = max(y1, by1[p]); }
            if (x1 - x0 + 1 > DW_MAXU || y1 - y0 + 1 > DW_MAXU) npass = P;
        }

        float2 acc[NCHUNK][NS];
#pragma unroll
        for (int c = 0; c < NCHUNK; ++c)
#pragma unroll
            for (int s = 0; s < NS; ++s) acc[c][s] = make_float2(0.f, 0.f);

        for (int pass = 0; pass < npass; ++pass) {
            int xlo = 1 << 30, xhi = -(1 << 30), ylo = 1 << 30, yhi = -(1 << 30);
#pragma unroll
            for (int p = 0; p < P; ++p)
                if (npass == 1 || p == pass) {
                    xlo = min(xlo, bx0[p]); xhi = max(xhi, bx1[p]); ylo = min(ylo, by0[p]); yhi = max(yhi, by1[p]);
                }
            if (xhi < xlo) continue;                                            /* (a spoke past the last one) */
            const int nux = min(xhi - xlo + 1, DW_MAXU), nuy = min(yhi - ylo + 1, DW_MAXU);

            /* phase A: lanes 0..nux-1 rows, the others (offset 16 would not cover 20) -> two passes */
            const int nuy4 = (nuy + 3) & ~3;                                    /* columns in whole quads: <= DW_MAXU */
            __syncwarp();
            for (int i = lane; i < nux + nuy4; i += 32) {
                const bool isrow = i < nux;
                const int u = isrow ? xlo + i : ylo + min(i - nux, nuy - 1);
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    float w4[DW_S], dd[DW_S];
#pragma unroll
                    for (int k = 0; k < DW_S; ++k) dd[k] = (float)u - (isrow ? X[p * DW_S + k] : Y[p * DW_S + k]);
                    /* two factors per packed evaluation (same FP32 operations as kb_weight); arguments outside the
                     * support give garbage that the selects drop */
                    const float2 k01 = kb_weight_pair(dd[0], dd[1], d.kb), k23 = kb_weight_pair(dd[2], dd[3], d.kb);
                    const float kk[DW_S] = { k01.x, k01.y, k23.x, k23.y };
#pragma unroll
                    for (int k = 0; k < DW_S; ++k) {
                        const bool live = (ro0 + k < d.nro) && (pe0 + p < d.npe) && (npass == 1 || p == pass)
                            && fabsf(dd[k]) < W && (isrow || i - nux < nuy);     /* tron.cu:343 via gridkernel */
                        w4[k] = live ? kk[k] : 0.f;
                    }
                    /* the pad columns of the last quad repeat the last column with zero weights (no new address) */
                    if (isrow) S.wx[p][i] = make_float4(w4[0], w4[1], w4[2], w4[3]);
                    else S.wy[p][i - nux] = make_float4(w4[0], w4[1], w4[2], w4[3]);
                }
                if (!isrow) S.coff[i - nux] = wrap_cell_w(u, n);                 /* periodic, tron.cu:570 */
            }
            if (lane < 8) S.coff[nuy4 + lane] = wrap_cell_w(ylo + nuy - 1, n);
            __syncwarp();

            /* phase B: lanes = channels (NCHUNK = 2: lane l holds channels 2l, 2l + 1 of the CTA's 64 -- one 16-byte load) */
            for (int i0 = 0; i0 < nux; i0 += RPI) {
                __syncwarp();
#pragma unroll
                for (int rr = 0; rr < RPI; ++rr)
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        const float4 a = i0 + rr < nux ? S.wx[p][i0 + rr] : make_float4(0.f, 0.f, 0.f, 0.f);
                        if (lane < nuy4) {
                            const float4 b = S.wy[p][lane];
                            S.wp[rr][p][lane] = make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
                        }
                    }
                __syncwarp();
                const int i = min(i0 + rsub, nux - 1);                          /* (rows past the window: zero weights) */
                const int row = wrap_cell_w(xlo + i, n);                        /* periodic, tron.cu:569 */
                const char *grow = (const char *)(gi + ((size_t)row * n) * nch + chan0 + clane * NCHUNK);
                const float4 *wrow = &S.wp[rsub][0][0];
                asm volatile("mov.b64 %0, %0;" : "+l"(grow));      /* one register pair: no uniform part added per address */
                /* the next quad of cells is in flight while this one is consumed (ptxas otherwise pairs every load with
                 * its first use); a column's address is one IMAD.WIDE: wrapped column index x cell stride + row base.
                 * (past the last quad: coff is padded with the last column, those loads are discarded) */
                float2 va[4][NCHUNK], vb[4][NCHUNK];
                load_quad<NCHUNK>(va, grow, &S.coff[0], cstride);
                for (int j0 = 0; j0 < nuy4; j0 += 8) {
                    load_quad<NCHUNK>(vb, grow, &S.coff[j0 + 4], cstride);
                    fma_quad<NCHUNK, P>(acc, va, wrow + j0);
                    if (j0 + 4 >= nuy4) break;
                    load_quad<NCHUNK>(va, grow, &S.coff[j0 + 8], cstride);
                    fma_quad<NCHUNK, P>(acc, vb, wrow + j0 + 4);
                }
            }
        }
        if (RPI > 1) {                                                      /* add the sub-warps' partial sums */
#pragma unroll
            for (int c = 0; c < NCHUNK; ++c)
#pragma unroll
                for (int s = 0; s < NS; ++s)
#pragma unroll
                    for (int o = LPC; o < 32; o <<= 1) {
                        acc[c][s].x += __shfl_xor_sync(0xffffffffu, acc[c][s].x, o);
                        acc[c][s].y += __shfl_xor_sync(0xffffffffu, acc[c][s].y, o);
                    }
        }
        if (rsub == 0 && chan0 + clane * NCHUNK < nch) {
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                const int pe = pe0 + s / DW_S, ro = ro0 + s % DW_S;
                if (ro >= d.nro || pe >= d.npe) continue;
                const size_t base = ((size_t)pe * d.nro + ro) * d.nc_total + d.ch0 + chan0 + clane * NCHUNK;
                store_chan<NCHUNK, HALF, NS>(d.samples, base, acc, s);
            }
        }
    }
}

/* ---------------------------------------------------------------------------------------------------------------
 * 8 channels (cfg5's coil shards on 8 GPUs).  In the kernel above the four quarter-warps take four ROWS of the window
 * at a time: every request is four 64-byte pieces of four different rows and the weights of four rows -- ncu: the L1
 * data pipe 91 % busy, 2.5 wavefronts per cell for four FFMA2.  Here the quarter-warps take four ADJACENT COLUMNS of
 * one row: in the channel-interleaved grid that is one contiguous 256-byte piece (two wavefronts for four cells), and
 * the four weights are neighbours in shared memory.  Loads run one step (16 columns) ahead, across rows.
 * ------------------------------------------------------------------------------------------------------------- */
template <int P> struct __align__(16) Dw8Weights {
    float4 wx[P][DW_MAXU];   /* row factor of samples 0..3 of spoke p */
    float4 wy[P][32];        /* column factor, zero past the window */
    float4 wp[P][32];        /* wx[row] * wy[column] of the row in flight */
    int coffT[32];           /* wrapped column indices, [step][sub-warp][quad]: one 16-byte load per lane */
};

/* CPL = lanes per cell: 8 (8 channels, four quarter-warps on four adjacent columns) or 16 with NCHUNK = 2 (32
 * channels, a lane holds two adjacent ones, the half-warps on two adjacent columns: cfg5's shards on 2 GPUs -- with a
 * whole warp per cell and one channel per lane a cell cost 2 + 4 wavefronts for 8 FFMA2, here a pair of cells costs
 * 4 + 2 for 32).
 * P = spokes per warp: with linear angle order the neighbours pe, pe + 1 share a window (see degrid_wide_kernel), a
 * loaded cell and the row's bookkeeping then serve 8 samples.  A pair whose union window does not fit is walked
 * spoke by spoke. */
template <bool HALF, int P, int CPL, int NCHUNK>
__global__ void __launch_bounds__(256, 2)
degrid_cols_kernel(const DegridLaunch d, const float2 *__restrict__ gi /* interleaved grid */)
{
    constexpr int NS = P * DW_S;
    constexpr int NSUB = 32 / CPL;                       /* sub-warps = adjacent columns taken together */
    constexpr int STEPC = 4 * NSUB;                      /* columns per step: four quads in flight */
    __shared__ Dw8Weights<P> sw8[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Dw8Weights<P> &S = sw8[warp];
    const int n = d.n, nch = d.nch;
    const int groups_per_spoke = (d.nro + DW_S - 1) / DW_S;
    const float W = d.kb.W;
    const float c0 = (float)((n + 1) / 2);
    const float inv_nro = rcp_approx((float)d.nro);
    const int clane = lane % CPL, rsub = lane / CPL;     /* channel lane, sub-warp = column inside a quad of columns */
    const int chan0 = blockIdx.y * CPL * NCHUNK;
    const unsigned cstride = (unsigned)nch * (unsigned)sizeof(float2);   /* bytes between neighbouring columns */
    const long long nwork = (long long)((d.npe + 8 * P - 1) / (8 * P)) * groups_per_spoke;
    for (long long wb = blockIdx.x; wb < nwork; wb += gridDim.x) {
        const int pe0 = ((int)(wb / groups_per_spoke) * 8 + warp) * P;
        const int ro0 = (int)(wb % groups_per_spoke) * DW_S;
        if (pe0 >= d.npe) continue;
        /* coordinates of the samples, exactly as tron.cu:554-561 compiles */
        float X[NS], Y[NS];
        int bx0[P], bx1[P], by0[P], by1[P];
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const float2 cs = __ldg(d.cs + min(pe0 + p, d.npe - 1));
            bx0[p] = by0[p] = 1 << 30; bx1[p] = by1[p] = -(1 << 30);
#pragma unroll
            for (int k = 0; k < DW_S; ++k) {
                const int s = p * DW_S + k;
                const float R = fma_ftz((float)(ro0 + k), inv_nro, -0.5f);
                const float nR = mul_ftz(R, (float)n);
                X[s] = fma_ftz(cs.y, nR, c0);            /* rows:    sin */
                Y[s] = fma_ftz(cs.x, nR, c0);            /* columns: cos */
                if (ro0 + k < d.nro && pe0 + p < d.npe) {
                    bx0[p] = min(bx0[p], (int)ceilf(X[s] - W)); bx1[p] = max(bx1[p], (int)floorf(X[s] + W));
                    by0[p] = min(by0[p], (int)ceilf(Y[s] - W)); by1[p] = max(by1[p], (int)floorf(Y[s] + W));
                }
            }
        }
        int npass = 1;
        if (P > 1) {
            int x0 = bx0[0], x1 = bx1[0], y0 = by0[0], y1 = by1[0];
#pragma unroll
            for (int p = 1; p < P; ++p) { x0 = min(x0, bx0[p]); x1 = max(x1, bx1[p]); y0 = min(y0, by0[p]); y1 = max(y1, by1[p]); }
            if (x1 - x0 + 1 > DW_MAXU || y1 - y0 + 1 > DW_MAXU) npass = P;
        }
        float2 acc[NCHUNK][NS];
#pragma unroll
        for (int c = 0; c < NCHUNK; ++c)
#pragma unroll
            for (int k = 0; k < NS; ++k) acc[c][k] = make_float2(0.f, 0.f);

        for (int pass = 0; pass < npass; ++pass) {
            int xlo = 1 << 30, xhi = -(1 << 30), ylo = 1 << 30, yhi = -(1 << 30);
#pragma unroll
            for (int p = 0; p < P; ++p)
                if (npass == 1 || p == pass) {
                    xlo = min(xlo, bx0[p]); xhi = max(xhi, bx1[p]); ylo = min(ylo, by0[p]); yhi = max(yhi, by1[p]);
                }
            if (xhi < xlo) continue;                                            /* (a spoke past the last one) */
            const int nux = min(xhi - xlo + 1, DW_MAXU), nuy = min(yhi - ylo + 1, DW_MAXU);
            const int nsteps = (nuy + STEPC - 1) / STEPC;                       /* <= 32 / STEPC */
            const int nuyP = nsteps * STEPC;

            /* phase A: lanes = rows, then columns */
            __syncwarp();
            for (int i = lane; i < nux + nuyP; i += 32) {
                const bool isrow = i < nux;
                const int c = i - nux;                                          /* column (when not a row) */
                const int u = isrow ? xlo + i : ylo + min(c, nuy - 1);
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    float dd[DW_S], w4[DW_S];
#pragma unroll
                    for (int k = 0; k < DW_S; ++k) dd[k] = (float)u - (isrow ? X[p * DW_S + k] : Y[p * DW_S + k]);
                    const float2 k01 = kb_weight_pair(dd[0], dd[1], d.kb), k23 = kb_weight_pair(dd[2], dd[3], d.kb);
                    const float kk[DW_S] = { k01.x, k01.y, k23.x, k23.y };
#pragma unroll
                    for (int k = 0; k < DW_S; ++k) {
                        const bool live = (ro0 + k < d.nro) && (pe0 + p < d.npe) && (npass == 1 || p == pass)
                            && fabsf(dd[k]) < W && (isrow || c < nuy);           /* tron.cu:343 via gridkernel */
                        w4[k] = live ? kk[k] : 0.f;
                    }
                    if (isrow) S.wx[p][i] = make_float4(w4[0], w4[1], w4[2], w4[3]);
                    else S.wy[p][c] = make_float4(w4[0], w4[1], w4[2], w4[3]);
                }
                /* column c = STEPC step + NSUB quad + sub-warp: stored as [step][sub-warp][quad] */
                if (!isrow) S.coffT[(c / STEPC) * STEPC + (c % NSUB) * 4 + ((c / NSUB) & 3)] = wrap_cell_w(u, n);   /* periodic, tron.cu:570 */
            }
            __syncwarp();

            /* phase B: lanes = (sub-warp = column of a quad, channel); the loads run one step ahead, across rows */
            int li = 0, ls = 0, fi = 0, fs = 0;                                 /* (row, step) cursors of loads and FMAs */
            auto load = [&](float2 (&v)[4][NCHUNK]) {
                const int row = wrap_cell_w(xlo + li, n);                       /* periodic, tron.cu:569 */
                const char *base = (const char *)(gi + ((size_t)row * n) * nch + chan0 + clane * NCHUNK);
                const int4 co = *reinterpret_cast<const int4 *>(&S.coffT[ls * STEPC + rsub * 4]);
                ldg_cell<NCHUNK>(v[0], base + (unsigned long long)(unsigned)co.x * cstride);
                ldg_cell<NCHUNK>(v[1], base + (unsigned long long)(unsigned)co.y * cstride);
                ldg_cell<NCHUNK>(v[2], base + (unsigned long long)(unsigned)co.z * cstride);
                ldg_cell<NCHUNK>(v[3], base + (unsigned long long)(unsigned)co.w * cstride);
                if (++ls == nsteps) { ls = 0; ++li; }
            };
            auto fma = [&](const float2 (&v)[4][NCHUNK]) {
                if (fs == 0) {                                                  /* a new row: its tap weights */
                    __syncwarp();
                    if (lane < nuyP) {
#pragma unroll
                        for (int p = 0; p < P; ++p) {
                            const float4 a = S.wx[p][fi], b = S.wy[p][lane];
                            S.wp[p][lane] = make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
                        }
                    }
                    __syncwarp();
                }
#pragma unroll
                for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        const float4 b = S.wp[p][fs * STEPC + NSUB * jj + rsub];
#pragma unroll
                        for (int c = 0; c < NCHUNK; ++c) {
                            ffma2d(acc[c][4 * p + 0], b.x, v[jj][c]); ffma2d(acc[c][4 * p + 1], b.y, v[jj][c]);
                            ffma2d(acc[c][4 * p + 2], b.z, v[jj][c]); ffma2d(acc[c][4 * p + 3], b.w, v[jj][c]);
                        }
                    }
                if (++fs == nsteps) { fs = 0; ++fi; }
            };
            const int nq = nux * nsteps;
            float2 va[4][NCHUNK], vb[4][NCHUNK];
            load(va);
            for (int q = 0; q < nq; q += 2) {
                if (q + 1 < nq) load(vb);
                fma(va);
                if (q + 1 >= nq) break;
                if (q + 2 < nq) load(va);
                fma(vb);
            }
        }
        /* add the sub-warps' partial sums */
#pragma unroll
        for (int c = 0; c < NCHUNK; ++c)
#pragma unroll
            for (int k = 0; k < NS; ++k)
#pragma unroll
                for (int o = CPL; o < 32; o <<= 1) {
                    acc[c][k].x += __shfl_xor_sync(0xffffffffu, acc[c][k].x, o);
                    acc[c][k].y += __shfl_xor_sync(0xffffffffu, acc[c][k].y, o);
                }
        if (rsub == 0 && chan0 + clane * NCHUNK < nch) {
#pragma unroll
            for (int k = 0; k < NS; ++k) {
                const int pe = pe0 + k / DW_S, ro = ro0 + k % DW_S;
                if (ro >= d.nro || pe >= d.npe) continue;
                const size_t base = ((size_t)pe * d.nro + ro) * d.nc_total + d.ch0 + chan0 + clane * NCHUNK;
                store_chan<NCHUNK, HALF, NS>(d.samples, base, acc, k);
            }
        }
    }
}

template <int P, int CPL, int NCHUNK>
static void launch_cols(const DegridLaunch &d, const float2 *scratch, cudaStream_t s)
{
    const long long nw = (long long)((d.nro + DW_S - 1) / DW_S) * ((d.npe + 8 * P - 1) / (8 * P));
    dim3 grid((unsigned)(nw < 148 * 32 ? nw : 148 * 32), (unsigned)(d.nch / (CPL * NCHUNK)));
    if (d.half_out) degrid_cols_kernel<true, P, CPL, NCHUNK><<<grid, 256, 0, s>>>(d, scratch);
    else            degrid_cols_kernel<false, P, CPL, NCHUNK><<<grid, 256, 0, s>>>(d, scratch);
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

template <int NCHUNK, int LPC, int P>
static void launch_dw(const DegridLaunch &d, const float2 *scratch, dim3 grid, cudaStream_t s)
{
    if (d.half_out) degrid_wide_kernel<NCHUNK, true, LPC, P><<<grid, 256, 0, s>>>(d, scratch);
    else            degrid_wide_kernel<NCHUNK, false, LPC, P><<<grid, 256, 0, s>>>(d, scratch);
}

/* scratch: nch*n*n float2, receives the channel-interleaved copy of the planar grid */
int launch_degrid_wide(const DegridLaunch &d, float2 *scratch, cudaStream_t s)
{
    const size_t ncell = (size_t)d.n * d.n;
    dim3 tg((unsigned)((ncell + 31) / 32), (unsigned)((d.nch + 31) / 32));
    planar_to_interleaved_kernel<<<tg, 256, 0, s>>>(scratch, d.grid, d.nch, ncell);
    TRON_CUDA(cudaGetLastError());
    /* spokes in pairs when consecutive spokes are neighbours in angle (linear order) */
    const int pair_env = getenv("TRON_DEGRID_PAIR") ? atoi(getenv("TRON_DEGRID_PAIR")) : -1;
    /* (from 16 channels on: the 16-coil cfg5 shard 4.01 -> 3.77 ms; the 8-channel ROW kernel was slower in pairs,
     * 2.81 -> 3.02 ms -- 8 channels have their own kernel below) */
    const bool pair = pair_env >= 0 ? pair_env != 0 : (d.pair_spokes != 0 && d.nch >= 16);
    const int P = pair ? 2 : 1;
    const long long nwork = (long long)((d.nro + DW_S - 1) / DW_S) * ((d.npe + 8 * P - 1) / (8 * P));
    int bx = (int)(nwork < 148 * 32 ? nwork : 148 * 32);
    const bool even2 = d.nc_total % 2 == 0 && d.ch0 % 2 == 0 && ((uintptr_t)d.samples) % 16 == 0;
    if (d.nch == 8 && getenv("TRON_DEGRID_ROWS8") == nullptr) {
        /* (the column kernel pairs spokes whenever the angle order is linear: it is bound by instructions, not L1) */
        const bool pair8 = pair_env >= 0 ? pair_env != 0 : d.pair_spokes != 0;
        if (pair8) launch_cols<2, 8, 1>(d, scratch, s); else launch_cols<1, 8, 1>(d, scratch, s);
    } else if (d.nch == 32 && even2 && getenv("TRON_DEGRID_ROWS32") == nullptr) {
        /* 32 channels (cfg5's shards on 2 GPUs): half-warps on adjacent columns, two channels per lane */
        if (pair) launch_cols<2, 16, 2>(d, scratch, s); else launch_cols<1, 16, 2>(d, scratch, s);
    } else if (d.nch == 8) {
        if (pair) launch_dw<1, 8, 2>(d, scratch, dim3(bx), s); else launch_dw<1, 8, 1>(d, scratch, dim3(bx), s);
    } else if (d.nch == 16) {
        if (pair) launch_dw<1, 16, 2>(d, scratch, dim3(bx), s); else launch_dw<1, 16, 1>(d, scratch, dim3(bx), s);
    } else if (d.nch % 64 == 0 && d.nc_total % 2 == 0 && d.ch0 % 2 == 0 && ((uintptr_t)d.samples) % 16 == 0) {
        /* two adjacent channels per lane: the 8- / 16-byte sample stores need even channel offsets */
        dim3 grid(bx, d.nch / 64);
        if (pair) launch_dw<2, 32, 2>(d, scratch, grid, s); else launch_dw<2, 32, 1>(d, scratch, grid, s);
    } else {
        dim3 grid(bx, d.nch / 32);
        if (pair) launch_dw<1, 32, 2>(d, scratch, grid, s); else launch_dw<1, 32, 1>(d, scratch, grid, s);
    }
    TRON_CUDA(cudaGetLastError());
    return 0;
}

} // namespace tronb
