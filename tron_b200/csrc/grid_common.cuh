/*
 * grid_common.cuh -- device helpers shared by the gridding kernels (grid.cu: one thread block per
 * (tile, slice group), taps read through L1; grid_tile.cu: persistent tile blocks with the spoke
 * segments staged in shared memory by bulk copies).  See grid.cu for the method.
 */
#pragma once
#include "tron_internal.h"

namespace tronb {

#define PI_F 3.14159274101257324219f
#define CELL_ALL_SPOKES 0x7fff

__device__ __forceinline__ int angle_bin(float a, float lut_scale, int nbins)
{
    int b = (int)(a * lut_scale);
    return min(b, nbins - 1);
}

/* ---------------------------------------------------------------------- */
/* the gather                                                              */
/* ---------------------------------------------------------------------- */

/* acc.xy += w * v.xy as one packed FP32x2 FMA (FFMA2 on sm_100a) */
__device__ __forceinline__ void ffma2(float2 &acc, float w, float2 v)
{
    unsigned long long a = *reinterpret_cast<unsigned long long *>(&acc);
    float2 ww = make_float2(w, w);
    asm("fma.rn.f32x2 %0, %1, %2, %0;"
        : "+l"(a)
        : "l"(*reinterpret_cast<unsigned long long *>(&ww)), "l"(*reinterpret_cast<unsigned long long *>(&v)));
    acc = *reinterpret_cast<float2 *>(&a);
}

/* CH channels of one sample (contiguous, channel fastest); fp16 storage converts on load.
 * The loads are volatile asm so that they stay where the source puts them -- ahead of the weight
 * evaluation -- instead of being sunk below the last branch that could still drop the tap. */
__device__ __forceinline__ float4 ldg_nc_f4(const void *p)
{
    float4 q;
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "l"(p));
    return q;
}
__device__ __forceinline__ float2 ldg_nc_f2(const void *p)
{
    float2 q;
    asm volatile("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(q.x), "=f"(q.y) : "l"(p));
    return q;
}
__device__ __forceinline__ uint2 ldg_nc_u2(const void *p)
{
    uint2 q;
    asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(q.x), "=r"(q.y) : "l"(p));
    return q;
}
__device__ __forceinline__ unsigned ldg_nc_u1(const void *p)
{
    unsigned q;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(q) : "l"(p));
    return q;
}

template <int CH, bool HALF>
__device__ __forceinline__ void load_sample(float2 (&v)[CH], const char *p)
{
    if (!HALF) {
        if (CH % 2 == 0) {
#pragma unroll
            for (int i = 0; i < CH / 2; ++i) {
                float4 q = ldg_nc_f4((const float4 *)p + i);
                v[2 * i] = make_float2(q.x, q.y); v[2 * i + 1] = make_float2(q.z, q.w);
            }
        } else {
#pragma unroll
            for (int i = 0; i < CH; ++i) v[i] = ldg_nc_f2((const float2 *)p + i);
        }
    } else {
        if (CH % 2 == 0) {
#pragma unroll
            for (int i = 0; i < CH / 2; ++i) {
                uint2 raw = ldg_nc_u2((const uint2 *)p + i);
                v[2 * i] = __half22float2(*reinterpret_cast<__half2 *>(&raw.x));
                v[2 * i + 1] = __half22float2(*reinterpret_cast<__half2 *>(&raw.y));
            }
        } else {
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                unsigned raw = ldg_nc_u1((const unsigned *)p + i);
                v[i] = __half22float2(*reinterpret_cast<__half2 *>(&raw));
            }
        }
    }
}

struct CellGeom { int X, Y, Rlo, Rhi, kstart, count; };

/* window [kstart, kstart+count) of sorted spokes that can reach the cell (circular in the table) */
__device__ __forceinline__ void cell_setup(const GridLaunch &g, const int *__restrict__ lut, int x, int y, CellGeom &c)
{
    const int n = g.n;
    c.X = x - n / 2; c.Y = y - n / 2;
    const int2 t = __ldg(g.cells + (size_t)y * n + x);
    c.Rlo = t.x & 0xffff; c.Rhi = t.x >> 16;
    c.kstart = 0; c.count = g.npe;
    const int lo16 = t.y & 0xffff;
    if (lo16 != CELL_ALL_SPOKES) {
        int b0 = (int)(short)lo16, b1 = t.y >> 16;
        bool wrap = false;
        if (b0 < 0) { b0 += g.nbins; wrap = true; }
        if (b1 >= g.nbins) { b1 -= g.nbins; wrap = true; }
        int ks = __ldg(lut + b0), ke = __ldg(lut + b1 + 1);
        c.kstart = ks;
        c.count = wrap ? (g.npe - ks) + ke : ke - ks;
    }
    if (c.Rlo > c.Rhi) c.count = 0;
}

/* Visit sorted-table entries kstart + first, kstart + first + step, ... (< count; the table is stored
 * twice, so the circular window is a plain range).  `samples` points at sample ro = 0 of the group's
 * first spoke (this thread's channel chunk).  The accumulators receive the final value: the output
 * scale 1/(nxos*npe) (tron.cu:532) rides on the density compensation factor.
 * PLAIN: the fitted Kaiser-Bessel polynomial is in use and nro == nxos (ridx = r) -- the usual case,
 * compiled without the run-time alternatives. */
template <int CH, int GS, bool HALF, bool PLAIN>
__device__ __forceinline__ void gather_cell(float2 (&acc)[GS][CH], const GridLaunch &g,
                                            const float4 *__restrict__ tab, const char *samples,
                                            const CellGeom &c, int first, int step)
{
    const float W = g.kb.W;
    const float Xf = (float)c.X, Yf = (float)c.Y, Rhif = (float)c.Rhi;
    const bool same = PLAIN || g.nro == g.n;              /* ridx = r (gridos 2) */
    const unsigned samp_bytes = (unsigned)g.nc_total * (unsigned)(HALF ? sizeof(__half2) : sizeof(float2));
    const int half_nro = g.nro >> 1;
    /* the table entry of the next spoke is fetched while the current one is processed */
    const float4 *tp = tab + c.kstart + first;
    float4 e = first < c.count ? __ldg(tp) : make_float4(1.f, 1.f, 0.f, 0.f);
    for (int left = c.count - first; left > 0; left -= step) {
        const float4 ec = e;                              /* ct, st, spoke index, slice mask */
        tp += step;
        if (left > step) e = __ldg(tp);
        /* candidate radii: integer points of {|r ct - X| < W} n {|r st - Y| < W} with a margin; a zero
         * cosine/sine gives +-inf bounds (or NaN when the cell cannot be reached: no candidates) */
        const float icx = rcp_approx(ec.x), icy = rcp_approx(ec.y);
        float ax = (Xf - W) * icx, bx = (Xf + W) * icx, ay = (Yf - W) * icy, by = (Yf + W) * icy;
        float lo = fmaxf(fminf(ax, bx), fminf(ay, by)) - 1e-3f;
        float hi = fminf(fmaxf(ax, bx), fmaxf(ay, by)) + 1e-3f;
        lo = fmaxf(lo, -Rhif); hi = fminf(hi, Rhif);
        if (!(lo <= hi)) continue;
        int r0 = (int)ceilf(lo), r1 = (int)floorf(hi);
        if (r0 > r1) continue;
        const int mask = GS > 1 ? __float_as_int(ec.w) : 1;
        if (GS > 1 && mask == 0) continue;               /* spoke outside every window of a partial group */
        const int centre = __float_as_int(ec.z) * g.nro + half_nro;       /* sample ro = nro/2 of this spoke */
        for (int r = r0; r <= r1; ++r) {
            if (abs(r) < c.Rlo) continue;                /* annulus, tron.cu:501-502,512,521 */
            float rf = (float)r;
            float dx = fma_ftz(ec.x, rf, -Xf);           /* tron.cu:514,516 as compiled */
            if (!(fabsf(dx) < W)) continue;
            float dy = fma_ftz(ec.y, rf, -Yf);
            if (!(fabsf(dy) < W)) continue;
            /* the tap is live: start the sample load, evaluate the weight while it is in flight */
            int ridx = same ? r : (r * g.nro) / g.n;     /* tron.cu:517 */
            float2 v[CH];
            load_sample<CH, HALF>(v, samples + (size_t)(unsigned)(centre + ridx) * samp_bytes);
            float w = PLAIN ? kb_poly_xy(dx, dy, g.kb) : kb_weight_xy(dx, dy, g.kb);
            float sdc = fmaf(g.sdc_as, fabsf((float)ridx), g.sdc_bs);        /* tron.cu:412, times the scale */
            w *= (r == 0) ? sdc + sdc : sdc;             /* both loops visit r = 0 */
            w = w > 0.f ? w : 0.f;                       /* the reference's wgt > 0 guard, branch-free so that
                                                            the loads above are not sunk below it */
#pragma unroll
            for (int s = 0; s < GS; ++s) {
                if (GS == 1 || (mask >> s) & 1) {
#pragma unroll
                    for (int i = 0; i < CH; ++i) ffma2(acc[s][i], w, v[i]);
                }
            }
        }
    }
}

/* blockIdx.x = slice group (fastest, so the groups of one tile run together), .y = tile rank, .z = chunk */
template <int CH, int GS, bool HALF>
__device__ __forceinline__ void group_pointers(const GridLaunch &g, int grp, int chunk, const float4 *&tab,
                                               const int *&lut, const char *&samples)
{
    const int ug = g.z0 / GS + grp;                       /* slice group, shard-local */
    const int tabi = g.tab_per_slice ? ug : 0;
    tab = g.tab_gx + (size_t)tabi * 2 * g.npe;
    lut = g.lut + (size_t)tabi * (g.nbins + 1);
    const size_t esz = HALF ? sizeof(__half2) : sizeof(float2);
    samples = (const char *)g.samples
        + ((size_t)ug * GS * g.slide * g.nro * g.nc_total + (size_t)(g.ch0 + chunk * CH)) * esz;
}

template <int CH, int GS>
__device__ __forceinline__ void store_cell(const GridLaunch &g, const float2 (&acc)[GS][CH], int grp, int chunk,
                                           int x, int y)
{
    const size_t plane = (size_t)g.n * g.n;
    const int zl0 = (g.z0 / GS + grp) * GS - g.z0;          /* slice index inside this launch of the group's first */
    float2 *out = g.grid + ((ptrdiff_t)zl0 * g.nch + (ptrdiff_t)chunk * CH) * (ptrdiff_t)plane + (size_t)y * g.n + x;
    const size_t slice_stride = (size_t)g.nch * plane;
#pragma unroll
    for (int s = 0; s < GS; ++s) {
        if (zl0 + s >= 0 && zl0 + s < g.nslices) {
            float2 *o = out;
#pragma unroll
            for (int i = 0; i < CH; ++i) { __stcs(o, acc[s][i]); o += plane; }   /* streaming: keep the samples in L2 */
        }
        out += slice_stride;
    }
}

/* main path: one thread per cell; cells inside the heavy disc are left to the heavy path */
template <int CH, int GS, bool HALF, int BT, bool PLAIN>
__device__ __forceinline__ void grid_tile_path(const GridLaunch &g, int rank, int grp, int chunk)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = g.n;
    const int tile = __ldg((BT == 128 ? g.tile_order8 : g.tile_order) + rank);      /* ty << 16 | tx */
    const int x = (tile & 0xffff) * 16 + (warp & 1) * 8 + (lane & 7);
    const int y = (tile >> 16) * (BT / 16) + (warp >> 1) * 4 + (lane >> 3);
    if (x >= n || y >= n) return;
    {
        const int X = x - n / 2, Y = y - n / 2;
        if (X * X + Y * Y > g.zero_r2) return;            /* beyond the last annulus: never fetched by the FFT pass */
    }

    const float4 *tab; const int *lut; const char *samples;
    group_pointers<CH, GS, HALF>(g, grp, chunk, tab, lut, samples);
    CellGeom c;
    cell_setup(g, lut, x, y, c);
    if (c.X * c.X + c.Y * c.Y <= g.heavy_r2) return;      /* integer test: identical on host and device */

    float2 acc[GS][CH];
#pragma unroll
    for (int s = 0; s < GS; ++s)
#pragma unroll
        for (int i = 0; i < CH; ++i) acc[s][i] = make_float2(0.f, 0.f);
    gather_cell<CH, GS, HALF, PLAIN>(acc, g, tab, samples, c, 0, 1);
    store_cell<CH, GS>(g, acc, grp, chunk, x, y);
}

/* heavy path: one warp per cell, lanes stride over the spokes, shuffle reduction */
template <int CH, int GS, bool HALF, int BT, bool PLAIN>
__device__ __forceinline__ void grid_heavy_path(const GridLaunch &g, int hg, int grp, int chunk)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ci = hg * (BT / 32) + warp;
    if (ci >= g.nheavy) return;
    const int packed = __ldg(g.heavy_cells + ci);
    const int x = packed & 0xffff, y = packed >> 16;

    const float4 *tab; const int *lut; const char *samples;
    group_pointers<CH, GS, HALF>(g, grp, chunk, tab, lut, samples);
    CellGeom c;
    cell_setup(g, lut, x, y, c);
    float2 acc[GS][CH];
#pragma unroll
    for (int s = 0; s < GS; ++s)
#pragma unroll
        for (int i = 0; i < CH; ++i) acc[s][i] = make_float2(0.f, 0.f);
    gather_cell<CH, GS, HALF, PLAIN>(acc, g, tab, samples, c, lane, 32);
#pragma unroll
    for (int s = 0; s < GS; ++s)
#pragma unroll
        for (int i = 0; i < CH; ++i) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                acc[s][i].x += __shfl_xor_sync(0xffffffffu, acc[s][i].x, o);
                acc[s][i].y += __shfl_xor_sync(0xffffffffu, acc[s][i].y, o);
            }
        }
    if (lane == 0) store_cell<CH, GS>(g, acc, grp, chunk, x, y);
}

/* ---------------------------------------------------------------------- */
/* shared-memory plumbing                                                  */
/* ---------------------------------------------------------------------- */
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    unsigned done;
    do {
        asm volatile("{\n.reg .pred p;\n"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                     "selp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
/* global -> shared bulk copy (TMA, no tensor map): 16-byte aligned addresses, size a multiple of 16 */
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ float4 lds_f4(unsigned a)
{
    float4 q;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "r"(a));
    return q;
}
__device__ __forceinline__ uint2 lds_u2(unsigned a)
{
    uint2 q;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(q.x), "=r"(q.y) : "r"(a));
    return q;
}

/* CH channels of one staged sample */
template <int CH, bool HALF>
__device__ __forceinline__ void lds_sample(float2 (&v)[CH], unsigned a)
{
    static_assert(CH % 2 == 0, "staged samples are read in 16-byte (8-byte for fp16 storage) pieces");
    if (!HALF) {
#pragma unroll
        for (int i = 0; i < CH / 2; ++i) {
            float4 q = lds_f4(a + 16 * i);
            v[2 * i] = make_float2(q.x, q.y); v[2 * i + 1] = make_float2(q.z, q.w);
        }
    } else {
#pragma unroll
        for (int i = 0; i < CH / 2; ++i) {
            uint2 raw = lds_u2(a + 8 * i);
            v[2 * i] = __half22float2(*reinterpret_cast<__half2 *>(&raw.x));
            v[2 * i + 1] = __half22float2(*reinterpret_cast<__half2 *>(&raw.y));
        }
    }
}

struct TileWindow { int k0, cnt; };

/* window [k0, k0 + cnt) of the sorted table (stored twice: never wraps) for a packed bin window */
__device__ __forceinline__ TileWindow window_of(int2 w, const int *__restrict__ lut, int nbins, int npe)
{
    TileWindow r; r.k0 = 0; r.cnt = npe;
    if (w.x != CELL_ALL_SPOKES) {
        int b0 = w.x, b1 = w.y;
        if (b1 < b0) { r.cnt = 0; return r; }
        bool wrap = false;
        if (b1 >= nbins) { b1 -= nbins; wrap = true; }
        const int ks = __ldg(lut + b0), ke = __ldg(lut + b1 + 1);
        r.k0 = ks;
        r.cnt = wrap ? (npe - ks) + ke : ke - ks;
    }
    return r;
}

/* 1/c and W/|c| + margin for the candidate run r in (X - W, X + W)/c = X/c -+ W/|c|; a vanishing cosine
 * or sine leaves the run unbounded on that axis (the reference predicate decides) */
__device__ __forceinline__ void axis_terms(float c, float W, float &ic, float &hw)
{
    if (fabsf(c) < 1e-30f) { ic = 0.f; hw = 1e30f; }
    else { ic = rcp_approx(c); hw = fmaf(W, fabsf(ic), 1e-3f); }
}


/* KB(dx) KB(dy): kb_poly_xy (refmath.cuh) with its instructions pinned behind the sample loads */
__device__ __forceinline__ float kb_poly_xy_c2(float dx, float dy, float invW, const unsigned long long (&c2)[TRONB_KB_DEG + 1])
{
    const float qx = dx * invW, qy = dy * invW;
    float2 u = make_float2(fmaf(-qx, qx, 1.0f), fmaf(-qy, qy, 1.0f));
    const unsigned long long U = *reinterpret_cast<unsigned long long *>(&u);
    unsigned long long p = c2[TRONB_KB_DEG];
#pragma unroll
    for (int m = TRONB_KB_DEG - 1; m >= 0; --m)
        asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p) : "l"(U), "l"(c2[m]));
    const float2 r = *reinterpret_cast<float2 *>(&p);
    return r.x * r.y;
}


} // namespace tronb
