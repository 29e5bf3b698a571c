/*
 * grid_wide.cu -- gridding for many receive channels (nc = 16, 32, 64, ...): lanes = channels.
 *
 * Same operator and the same tap set as grid.cu (reference: precompensate + gridradial2d,
 * /root/reference/src/tron.cu:405-416, 465-536); different decomposition.  With one thread per
 * cell every tap costs nc scattered 8-byte loads per thread; here a warp owns a block of 4 x 2
 * cells and works in two phases:
 *
 *   A  lanes = spokes of the block's angular window.  Each lane walks the radii of its spoke that
 *      fall into the block's support box, evaluates the reference predicate for the 8 cells
 *      (fused fma, annulus, r = 0 twice -- refmath.cuh), the separable Kaiser-Bessel weights
 *      (4 + 2 evaluations serve 8 cells) and the density ramp, and appends
 *      (sample offset, 8 weights, slice mask) to a per-warp list in shared memory.
 *   B  lanes = channels.  For every list entry the warp loads the sample's channels with one
 *      coalesced request (256 B for 32 channels) and feeds the 8 cells' accumulators with packed
 *      FFMA2s: each sample is fetched once per block instead of once per cell (3.7x fewer
 *      requests), and every request is a full cache line.
 *
 * The list is drained whenever it could overflow, so blocks near DC (thousands of taps) need no
 * special path.  Sliding golden-angle windows share phase A across GS = 4 slices like grid.cu.
 * nc = 16 / 8 (coil shards at wide kernels) run two / four entries per step on sub-warps and fold
 * them at the end; nc = 64 gives a lane two ADJACENT channels (one request per sample).
 * Output: planar grid[slice][ch][row][col]; the CTA's eight blocks tile 16 x 4 cells and are
 * transposed through shared memory, so a channel's rows leave as 128 contiguous bytes.
 *
 * What bounds it (ncu, profiles/r02_ncu_cfg3_grid_wide_v2.txt): the L1 data pipe -- an entry is
 * 4 wavefronts for its weights (a broadcast LDS.128 costs two) + 2 for the sample per 8 FFMA2 --
 * not the instruction count.
 */
#include "tron_internal.h"
#include <stdlib.h>

namespace tronb {

#define PI_F 3.14159274101257324219f
#define WCAP 96                       /* list entries per warp */
#ifndef WIDE_PIPE_SUB
#define WIDE_PIPE_SUB 1               /* sub-warp instantiations (8 / 16 channels) with fp16 samples: pipelined drain (16-coil cfg5 shard 4.05 -> 2.85 ms; 32 channels per warp: 3.81 -> 4.00, stays off) */
#endif

struct __align__(16) WideList {
    float4 wa[WCAP];                  /* weights of cells (0..3, row 0) */
    float4 wb[WCAP];                  /* weights of cells (0..3, row 1) */
    unsigned off[WCAP + 8];           /* sample offset in elements from the group's channel base (launch checks it fits 32 bits) */
    int mask[WCAP];                   /* slices of the group whose window holds the spoke */
};

/* volatile: the drain loop relies on program order (all sample loads of a step, then the FMAs) --
 * left to itself ptxas pairs each load with its first use and no load is ever in flight */
__device__ __forceinline__ void ffma2w(float2 &acc, float w, float2 v)
{
    unsigned long long a = *reinterpret_cast<unsigned long long *>(&acc);
    float2 ww = make_float2(w, w);
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;"
        : "+l"(a)
        : "l"(*reinterpret_cast<unsigned long long *>(&ww)), "l"(*reinterpret_cast<unsigned long long *>(&v)));
    acc = *reinterpret_cast<float2 *>(&a);
}

/* the NCHUNK adjacent channels a lane holds of one sample: ONE request of 4 .. 16 bytes, kept as loaded (half
 * pairs stay packed until they are used: a step in flight costs one register per channel) */
template <int NCHUNK, bool HALF> struct LaneRaw { unsigned long long q[HALF ? (NCHUNK + 1) / 2 : NCHUNK]; };

template <int NCHUNK, bool HALF>
__device__ __forceinline__ void load_lane(LaneRaw<NCHUNK, HALF> &x, const char *p)
{
    if (HALF) {
        if (NCHUNK == 2) {
            asm volatile("ld.global.nc.b64 %0, [%1];" : "=l"(x.q[0]) : "l"(p));
        } else {
            unsigned raw;
            asm volatile("ld.global.nc.b32 %0, [%1];" : "=r"(raw) : "l"(p));
            x.q[0] = raw;
        }
    } else {
        if (NCHUNK == 2)
            asm volatile("ld.global.nc.v2.b64 {%0, %1}, [%2];" : "=l"(x.q[0]), "=l"(x.q[NCHUNK - 1]) : "l"(p));
        else
            asm volatile("ld.global.nc.b64 %0, [%1];" : "=l"(x.q[0]) : "l"(p));
    }
}

template <int NCHUNK, bool HALF>
__device__ __forceinline__ float2 lane_value(const LaneRaw<NCHUNK, HALF> &x, int c)
{
    if (HALF) {
        const unsigned raw = (unsigned)(c ? (x.q[0] >> 32) : x.q[0]);
        return __half22float2(*reinterpret_cast<const __half2 *>(&raw));
    }
    return *reinterpret_cast<const float2 *>(&x.q[c]);
}

/* the sample loads of DEPTH list steps */
template <int LPC, int NCHUNK, bool HALF, int DEPTH>
__device__ __forceinline__ void drain_load(LaneRaw<NCHUNK, HALF> (&x)[DEPTH], const WideList &L, int e0, const char *lbase, int sub)
{
    constexpr int EPI = 32 / LPC;
    constexpr unsigned ROWB = HALF ? 4u : 8u;       /* bytes per channel */
    if (EPI == 1 && DEPTH % 4 == 0) {                /* the step's offsets: 16-byte shared loads */
#pragma unroll
        for (int d = 0; d < DEPTH; d += 4) {
            const uint4 o4 = *reinterpret_cast<const uint4 *>(&L.off[e0 + d]);
            load_lane<NCHUNK, HALF>(x[d + 0], lbase + (unsigned long long)o4.x * ROWB);
            load_lane<NCHUNK, HALF>(x[d + 1], lbase + (unsigned long long)o4.y * ROWB);
            load_lane<NCHUNK, HALF>(x[d + 2], lbase + (unsigned long long)o4.z * ROWB);
            load_lane<NCHUNK, HALF>(x[d + 3], lbase + (unsigned long long)o4.w * ROWB);
        }
    } else {
#pragma unroll
        for (int d = 0; d < DEPTH; ++d)
            load_lane<NCHUNK, HALF>(x[d], lbase + (unsigned long long)L.off[e0 + d * EPI + sub] * ROWB);
    }
}

template <int LPC, int NCHUNK, int GS, bool HALF, int DEPTH>
__device__ __forceinline__ void drain_fma(float2 (&acc)[NCHUNK][GS][8], const LaneRaw<NCHUNK, HALF> (&x)[DEPTH], const WideList &L,
                                          int e0, int sub)
{
    constexpr int EPI = 32 / LPC;
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) {
        const int e = e0 + d * EPI + sub;
        const float4 p = L.wa[e], q = L.wb[e];
        const int m = GS > 1 ? L.mask[e] : 1;
        float2 v[NCHUNK];
#pragma unroll
        for (int c = 0; c < NCHUNK; ++c) v[c] = lane_value<NCHUNK, HALF>(x[d], c);
#pragma unroll
        for (int s = 0; s < GS; ++s)
            if (GS == 1 || (m >> s) & 1) {
#pragma unroll
                for (int c = 0; c < NCHUNK; ++c) {
                    ffma2w(acc[c][s][0], p.x, v[c]); ffma2w(acc[c][s][1], p.y, v[c]);
                    ffma2w(acc[c][s][2], p.z, v[c]); ffma2w(acc[c][s][3], p.w, v[c]);
                    ffma2w(acc[c][s][4], q.x, v[c]); ffma2w(acc[c][s][5], q.y, v[c]);
                    ffma2w(acc[c][s][6], q.z, v[c]); ffma2w(acc[c][s][7], q.w, v[c]);
                }
            }
    }
}

/* phase B: consume the list.  LPC = lanes per entry (32, or 16: two entries per step).
 * The list is first padded to a whole number of steps with zero-weight copies of its last entry
 * (same sample, so no new address is touched), which keeps the loop body free of branches.
 * Two buffers of DEPTH steps alternate: the loads of the next DEPTH steps are issued before the FMAs of the
 * current ones, so a warp always has sample loads in flight (ncu: a quarter of the stall samples sat on the
 * first use of a step's samples when loads and FMAs took turns).
 * `lbase` = the lane's first channel of the group's first sample; a sample's address is one
 * IMAD.WIDE (element offset x element size + lane base). */
template <int LPC, int NCHUNK, int GS, bool HALF, bool PIPE>
__device__ __forceinline__ void wide_drain(float2 (&acc)[NCHUNK][GS][8], WideList &L, int cnt,
                                           const char *lbase, int lane)
{
    constexpr int EPI = 32 / LPC;                    /* entries per step */
    /* one channel chunk (64-register instantiations, cfg3): the second buffer would spill -- 8 steps at a time, loads
     * then FMAs (measured: 2.91 vs 3.12 ms per 32 cfg3 slices) */
    constexpr int DEPTH = !PIPE ? 8 / EPI : ((HALF || NCHUNK == 1) ? (EPI == 4 ? 2 : 4) : 2);
    constexpr int STEP = DEPTH * EPI;                /* <= 8 */
    if (cnt == 0) return;
    const int padded = ((cnt + STEP - 1) / STEP) * STEP;     /* <= WCAP: WCAP is a multiple of STEP */
    __syncwarp();                                    /* the entries other lanes appended (the pad copies the last offset) */
    if (cnt + lane < padded + STEP) {                /* (offsets one step further: the loads run one step ahead) */
        if (cnt + lane < padded) {
            L.wa[cnt + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
            L.wb[cnt + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
            L.mask[cnt + lane] = 0;
        }
        L.off[cnt + lane] = L.off[cnt - 1];
    }
    __syncwarp();
    const int sub = lane / LPC;
    if (!PIPE) {
        for (int e0 = 0; e0 < padded; e0 += STEP) {
            LaneRaw<NCHUNK, HALF> x[DEPTH];
            drain_load<LPC, NCHUNK, HALF, DEPTH>(x, L, e0, lbase, sub);
            drain_fma<LPC, NCHUNK, GS, HALF, DEPTH>(acc, x, L, e0, sub);
        }
        __syncwarp();
        return;
    }
    LaneRaw<NCHUNK, HALF> xa[DEPTH], xb[DEPTH];
    drain_load<LPC, NCHUNK, HALF, DEPTH>(xa, L, 0, lbase, sub);
    for (int e0 = 0; e0 < padded; e0 += 2 * STEP) {
        drain_load<LPC, NCHUNK, HALF, DEPTH>(xb, L, e0 + STEP, lbase, sub);
        drain_fma<LPC, NCHUNK, GS, HALF, DEPTH>(acc, xa, L, e0, sub);
        if (e0 + STEP >= padded) break;
        drain_load<LPC, NCHUNK, HALF, DEPTH>(xa, L, min(e0 + 2 * STEP, padded), lbase, sub);
        drain_fma<LPC, NCHUNK, GS, HALF, DEPTH>(acc, xb, L, e0 + STEP, sub);
    }
    __syncwarp();
}

/* one channel chunk, one slice per group (cfg3): 4 CTAs/SM (64 registers, 26 words of spill) measured 8 % faster
 * than 2 (128 registers); the wider instantiations spill too much below 128 registers (cfg5: 3 CTAs/SM 35 % slower) */
template <int LPC, int NCHUNK, int GS, bool HALF, int MINB>
__global__ void __launch_bounds__(256, MINB)
grid_wide_kernel(const GridLaunch g)
{
    __shared__ WideList lists[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WideList &L = lists[warp];
    const int n = g.n;
    const int tiles_x = (n + 15) >> 4;
    /* a CTA takes a quarter of a 16 x 16 tile (8 blocks of 4 x 2 cells, one per warp): with a single slice in
     * flight (cfg5) whole-tile CTAs near DC ran 100x longer than the average and a third of the SM time was tail */
    const int unit = blockIdx.x / g.ngroups, grp = blockIdx.x - unit * g.ngroups;
    const int rank = unit >> 2, quarter = unit & 3;
    const int tile = __ldg(g.tile_order + rank);
    const int ty = tile >> 16, tx = tile & 0xffff;
    const int chan0 = blockIdx.y * (LPC * NCHUNK);           /* first plan-local channel of this CTA */

    const int ug = g.z0 / GS + grp;
    const int tabi = g.tab_per_slice ? ug : 0;
    const float4 *tab = g.tab_cs + (size_t)tabi * g.npe;
    const int *tpe = g.tab_pe + (size_t)tabi * g.npe;
    const int *lut = g.lut + (size_t)tabi * (g.nbins + 1);
    const size_t esz = HALF ? sizeof(__half2) : sizeof(float2);
    /* lanes = channels: NCHUNK = 2 gives a lane the ADJACENT channels 2l, 2l + 1 (one request per sample); lanes past
     * the last channel of a partial chunk recompute the last one and never store */
    const int nvalid = g.nch - chan0;
    const int chl = NCHUNK == 2 ? 2 * lane : min(lane % LPC, nvalid - 1);
    const char *lbase = (const char *)g.samples
        + ((size_t)ug * GS * g.slide * g.nro * g.nc_total + (size_t)(g.ch0 + chan0 + chl)) * esz;
    const char *pbase = (const char *)g.samples
        + ((size_t)ug * GS * g.slide * g.nro * g.nc_total + (size_t)(g.ch0 + chan0)) * esz;
    /* opaque to the compiler: otherwise it keeps the kernel argument in a uniform register and adds it with two more
     * instructions per sample address */
    asm volatile("mov.b64 %0, %0;" : "+l"(lbase));

    const float W = g.kb.W;
    const bool same = g.nro == g.n;
    const float Rmaxf = (float)(n / 2 - 1);
    const float lut_scale = (float)g.nbins / PI_F;

    {
        const int bi = quarter * 8 + warp;
        const int x0 = tx * 16 + (bi & 3) * 4, y0 = ty * 16 + (bi >> 2) * 2;
        const bool inside = x0 < n && y0 < n;                 /* (n not a multiple of 16: blocks past the edge idle) */
        const int X0 = x0 - n / 2, Y0 = y0 - n / 2;

        /* annulus of the 8 cells (cell c = cy*4 + cx), broadcast from lanes 0..7 */
        int myband = 0;
        if (lane < 8 && inside) myband = __ldg(g.cells + (size_t)(y0 + (lane >> 2)) * n + x0 + (lane & 3)).x;
        int band[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) band[c] = __shfl_sync(0xffffffffu, myband, c);
        bool dead = true;
#pragma unroll
        for (int c = 0; c < 8; ++c) dead = dead && ((band[c] & 0xffff) > (band[c] >> 16));
        dead = dead || !inside;

        float2 acc[NCHUNK][GS][8];
#pragma unroll
        for (int c = 0; c < NCHUNK; ++c)
#pragma unroll
            for (int s = 0; s < GS; ++s)
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[c][s][i] = make_float2(0.f, 0.f);

        if (!dead) {
            /* spokes whose line passes within reach of the block (centre (X0+1.5, Y0+0.5), half diagonal 1.59) */
            const float Xc = (float)X0 + 1.5f, Yc = (float)Y0 + 0.5f;
            const float Rc = sqrtf(Xc * Xc + Yc * Yc);
            const float reach = W * 1.41421368f + 1.59f + 2e-3f;
            const float xr = reach / fmaxf(Rc, 1e-6f);
            int kstart = 0, count = g.npe;
            if (xr <= 0.7f) {
                float T = atan2f(Yc, Xc);
                if (T < 0.f) T += PI_F;
                if (T >= PI_F) T -= PI_F;
                const float delta = xr * fmaf(0.25f * xr, xr, 1.0f) + 2e-4f;
                int b0 = (int)floorf((T - delta) * lut_scale);
                int b1 = (int)floorf((T + delta) * lut_scale);
                if (b1 - b0 + 1 < g.nbins) {
                    bool wrap = false;
                    if (b0 < 0) { b0 += g.nbins; wrap = true; }
                    if (b1 >= g.nbins) { b1 -= g.nbins; wrap = true; }
                    const int ks = __ldg(lut + b0), ke = __ldg(lut + b1 + 1);
                    kstart = ks;
                    count = wrap ? (g.npe - ks) + ke : ke - ks;
                }
            }
            const float xlo = (float)X0 - W, xhi = (float)(X0 + 3) + W;
            const float ylo = (float)Y0 - W, yhi = (float)(Y0 + 1) + W;
            int cnt = 0;

            for (int it0 = 0; it0 < count; it0 += 32) {
                /* phase A, step 1: lane = spoke of the window: its radii inside the block's support box */
                const int it = it0 + lane;
                const bool have = it < count;
                int k = kstart + (have ? it : 0);
                if (k >= g.npe) k -= g.npe;
                const float4 e = __ldg(tab + k);                 /* ct, st, 1/ct, 1/st */
                const int pm = __ldg(tpe + k);
                float ax = xlo * e.z, bx = xhi * e.z, ay = ylo * e.w, by = yhi * e.w;
                float lo = fmaxf(fminf(ax, bx), fminf(ay, by)) - 1e-3f;
                float hi = fminf(fmaxf(ax, bx), fmaxf(ay, by)) + 1e-3f;
                /* two-sided clamps: with ct or st near 0 the bounds reach +-1e18 and the int length would wrap */
                lo = fminf(fmaxf(lo, -Rmaxf), Rmaxf + 1.f); hi = fmaxf(fminf(hi, Rmaxf), -Rmaxf - 1.f);
                const int r0 = (int)ceilf(lo);
                int len = (int)floorf(hi) - r0 + 1;
                if (len < 0 || !have || (GS > 1 && (pm >> 24) == 0)) len = 0;
                /* exclusive prefix of the lengths: candidate f of this batch belongs to the last lane with pre <= f */
                int pre = len;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, pre, o);
                    if (lane >= o) pre += t;
                }
                const int total = __shfl_sync(0xffffffffu, pre, 31);
                pre -= len;
                /* step 2: lane = candidate (spoke, radius), flattened so that all lanes work */
                for (int f0 = 0; f0 < total; f0 += 32) {
                    const int f = f0 + lane;
                    int own = 0;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const int cand = own + o;
                        const int pc = __shfl_sync(0xffffffffu, pre, cand & 31);
                        if (cand < 32 && pc <= f) own = cand;
                    }
                    const float ct = __shfl_sync(0xffffffffu, e.x, own), st = __shfl_sync(0xffffffffu, e.y, own);
                    const int pmo = __shfl_sync(0xffffffffu, pm, own);
                    const int r = __shfl_sync(0xffffffffu, r0, own) + (f - __shfl_sync(0xffffffffu, pre, own));
                    const float rf = (float)r;
                    const int ar = abs(r);
                    const int ridx = same ? r : (r * g.nro) / g.n;                      /* tron.cu:517 */
                    float w8[8];
                    bool any = false;
                    if (f < total) {
                        float wx[4], wy[2];
                        float dxs[4], dys[2];
#pragma unroll
                        for (int cx = 0; cx < 4; ++cx) dxs[cx] = fma_ftz(ct, rf, -(float)(X0 + cx));   /* tron.cu:514 as compiled */
#pragma unroll
                        for (int cy = 0; cy < 2; ++cy) dys[cy] = fma_ftz(st, rf, -(float)(Y0 + cy));
                        float fs = fmaf(g.sdc_a, fabsf((float)ridx), g.sdc_b);          /* tron.cu:412 */
                        if (r == 0) fs += fs;                                            /* r = 0 visited twice */
                        {
                            /* three packed evaluations serve the six factors (the same FP32 operations per factor as
                             * kb_weight -- fitted polynomial or rational form -- so the weights are bit-identical);
                             * arguments outside the support give garbage that the selects drop */
                            const float2 a = kb_weight_pair(dxs[0], dxs[1], g.kb), b = kb_weight_pair(dxs[2], dxs[3], g.kb);
                            const float2 c = kb_weight_pair(dys[0], dys[1], g.kb);
                            wx[0] = fabsf(dxs[0]) < W ? a.x : 0.f; wx[1] = fabsf(dxs[1]) < W ? a.y : 0.f;
                            wx[2] = fabsf(dxs[2]) < W ? b.x : 0.f; wx[3] = fabsf(dxs[3]) < W ? b.y : 0.f;
                            wy[0] = fabsf(dys[0]) < W ? c.x * fs : 0.f; wy[1] = fabsf(dys[1]) < W ? c.y * fs : 0.f;
                        }
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const bool inband = ar >= (band[c] & 0xffff) && ar <= (band[c] >> 16);
                            const float w = inband ? wx[c & 3] * wy[c >> 2] : 0.f;
                            w8[c] = w > 0.f ? w : 0.f;
                            any = any || w > 0.f;
                        }
                    }
                    const unsigned hit = __ballot_sync(0xffffffffu, any);
                    if (any) {
                        const int slot = cnt + __popc(hit & ((1u << lane) - 1u));
                        L.wa[slot] = make_float4(w8[0], w8[1], w8[2], w8[3]);
                        L.wb[slot] = make_float4(w8[4], w8[5], w8[6], w8[7]);
                        const unsigned off = (unsigned)(((pmo & 0xffffff) * g.nro + g.nro / 2 + ridx)) * (unsigned)g.nc_total;
                        L.off[slot] = off;
                        L.mask[slot] = pmo >> 24;
                        /* the drain comes up to WCAP entries later: start the sample's lines towards L2 / L1 now (one
                         * warp instruction covers 32 samples), so that its loads do not wait for DRAM */
                        if (g.wide_prefetch) {
                            const char *pf = pbase + (unsigned long long)off * (HALF ? 4u : 8u);
#pragma unroll
                            for (int b = 0; b < LPC * NCHUNK * (HALF ? 4 : 8); b += 128) {
                                if (g.wide_prefetch == 1) asm volatile("prefetch.global.L1 [%0];" :: "l"(pf + b));
                                else asm volatile("prefetch.global.L2 [%0];" :: "l"(pf + b));
                            }
                        }
                    }
                    cnt += __popc(hit);
                    if (cnt + 32 > WCAP) {                       /* phase B: lanes = channels */
                        wide_drain<LPC, NCHUNK, GS, HALF, (NCHUNK == 2 || MINB == 3 || (HALF && LPC < 32 && WIDE_PIPE_SUB))>(acc, L, cnt, lbase, lane);
                        cnt = 0;
                    }
                }
            }
            wide_drain<LPC, NCHUNK, GS, HALF, (NCHUNK == 2 || MINB == 3 || (HALF && LPC < 32 && WIDE_PIPE_SUB))>(acc, L, cnt, lbase, lane);
        }

        /* fold the two half-warps (nc = 16 mode), then every lane writes its channel plane */
        if (LPC < 32) {
#pragma unroll
            for (int s = 0; s < GS; ++s)
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int o = 16; o >= LPC; o >>= 1) {
                        acc[0][s][i].x += __shfl_down_sync(0xffffffffu, acc[0][s][i].x, o);
                        acc[0][s][i].y += __shfl_down_sync(0xffffffffu, acc[0][s][i].y, o);
                    }
        }
        /* Store.  A lane holds ONE channel of the warp's 4 x 2 cells: written from here, every lane of a store would
         * touch its own 128-byte line (ncu: half of the kernel's L1 tag requests were these stores).  The CTA's 8
         * blocks tile 16 x 4 cells, so they are transposed through shared memory (the lists are no longer needed):
         * per channel the tile is 4 rows of 128 contiguous bytes, one 16-byte store per lane. */
        constexpr int CSTR = 66;                          /* float2 per channel: 64 cells + 16 bytes (bank groups) */
        float2 *stage = reinterpret_cast<float2 *>(lists);
        static_assert(sizeof(lists) >= (size_t)LPC * CSTR * sizeof(float2), "staging tile does not fit the lists");
        const size_t plane = (size_t)n * n;
        const int zg = ug * GS;
        const int xb = tx * 16, yb = ty * 16 + quarter * 4;
        const int brow = ((bi >> 2) & 1) * 2, bcol = (bi & 3) * 4;
        __syncthreads();                                  /* every warp has drained its last list */
#pragma unroll
        for (int s = 0; s < GS; ++s) {
            const int zl = zg + s - g.z0;
            if (zl < 0 || zl >= g.nslices) continue;      /* (uniform over the CTA) */
#pragma unroll
            for (int c = 0; c < NCHUNK; ++c) {
                if (lane < LPC) {
#pragma unroll
                    for (int cy = 0; cy < 2; ++cy) {
                        float4 *o4 = reinterpret_cast<float4 *>(stage + lane * CSTR + (brow + cy) * 16 + bcol);
                        const float2 *a = &acc[c][s][cy * 4];
                        o4[0] = make_float4(a[0].x * g.scale, a[0].y * g.scale, a[1].x * g.scale, a[1].y * g.scale);
                        o4[1] = make_float4(a[2].x * g.scale, a[2].y * g.scale, a[3].x * g.scale, a[3].y * g.scale);
                    }
                }
                __syncthreads();
                const int row = lane >> 3, col = (lane & 7) * 2;
                for (int chl = warp; chl < LPC; chl += 8) {
                    const int ch = NCHUNK == 2 ? chan0 + 2 * chl + c : chan0 + chl;
                    if (ch >= g.nch || yb + row >= n || xb + col >= n) continue;
                    const float4 v = *reinterpret_cast<const float4 *>(stage + chl * CSTR + row * 16 + col);
                    *reinterpret_cast<float4 *>(g.grid + ((size_t)zl * g.nch + ch) * plane + (size_t)(yb + row) * n + xb + col) = v;
                }
                __syncthreads();
            }
        }
    }
}

template <int LPC, int NCHUNK, int GS, int MINB = ((NCHUNK == 1 && GS == 1) ? 4 : 2)>
static int launch_wide(GridLaunch g, cudaStream_t s)
{
    int tiles = ((g.n + 15) / 16) * ((g.n + 15) / 16);
    g.ngroups = (g.z0 + g.nslices - 1) / GS - g.z0 / GS + 1;
    int per_cta = LPC * NCHUNK;
    dim3 grid(tiles * 4 * g.ngroups, (g.nch + per_cta - 1) / per_cta);
    if (g.half_in) grid_wide_kernel<LPC, NCHUNK, GS, true, MINB><<<grid, 256, 0, s>>>(g);
    else           grid_wide_kernel<LPC, NCHUNK, GS, false, MINB><<<grid, 256, 0, s>>>(g);
    TRON_CUDA(cudaGetLastError());
    return 0;
}

bool grid_wide_applicable(const GridLaunch &g)
{
    /* nc = 16 is faster on the thread-per-cell kernel (measured: 9.3 vs 18.3 us/slice on cfg4) */
    /* ... for the default kernel width; with wide kernels (cfg5's 16-coil shards at -k 6) a sample serves dozens of
     * blocks and fetching it once per block wins again */
    const bool few = (g.nch == 16 || (g.nch == 8 && getenv("TRON_NO_WIDE8") == nullptr)) && g.kb.W >= 3.f;
    if (!few && (g.nch < 32 || g.nch % 16 != 0)) return false;
    if (g.n % 4 != 0) return false;
    if (g.gs != 1 && g.gs != 4) return false;
    /* sample offsets inside a table's window are 32-bit element counts: (window spokes * nro) * nc must fit */
    if ((unsigned long long)g.npe * (unsigned long long)g.nro * (unsigned long long)g.nc_total >= (1ull << 32)) return false;
    if ((unsigned long long)g.npe * (unsigned long long)g.nro >= (1ull << 31)) return false;
    return true;
}

int launch_grid_wide(const GridLaunch &g_in, cudaStream_t s)
{
    if (g_in.nslices <= 0 || g_in.nch <= 0) return 0;
    GridLaunch g = g_in;
    static const int pf = getenv("TRON_WIDE_PREFETCH") ? atoi(getenv("TRON_WIDE_PREFETCH")) : 2;
    g.wide_prefetch = pf;
    if (g.gs == 4) {
        if (g.nch == 16) return launch_wide<16, 1, 4>(g, s);
        if (g.nch == 8) return launch_wide<8, 1, 4>(g, s);
        return launch_wide<32, 1, 4>(g, s);              /* 32 channels per CTA row */
    }
    if (g.nch == 16) return launch_wide<16, 1, 1>(g, s);
    if (g.nch == 8) return launch_wide<8, 1, 1>(g, s);           /* cfg5's shards on 8 GPUs: four list entries per step */
    /* (32 channels at 3 blocks per SM, 80 registers, pipelined drain: 3.27 vs 2.93 ms per 32 cfg3 slices -- 4 blocks stay) */
    /* two adjacent channels per lane: the 8- / 16-byte sample loads need even channel offsets */
    const size_t esz = g.half_in ? 4 : 8;
    if (g.nch % 64 == 0 && g.nc_total % 2 == 0 && g.ch0 % 2 == 0 && ((uintptr_t)g.samples) % (2 * esz) == 0)
        return launch_wide<32, 2, 1>(g, s);
    return launch_wide<32, 1, 1>(g, s);
}

} // namespace tronb
