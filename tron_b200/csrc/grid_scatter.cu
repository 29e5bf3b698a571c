/*
 * grid_scatter.cu -- gridding with the ACCUMULATORS in shared memory: sample-driven, tile-owned, no atomics.
 *
 * Same operator, same tap set as grid.cu / grid_tile.cu (precompensate + gridradial2d,
 * /root/reference/src/tron.cu:405-416 and 465-536).  Those kernels keep a cell's accumulators in registers and
 * search, per cell, for the samples that tap it; with a sliding window that advances by a few spokes per slice
 * a cell has only one or two new taps per slice, the search dominates and the lanes of a warp run out of step
 * (ncu, profiles/r02_ncu_tile_v2.txt: 14 of 32 lanes active in the tap body, 6 of 32 in the predicated FFMA2).
 *
 * Here a WARP owns a 16 x 16 tile of one slice; the tile's accumulators (16 x 16 cells x nc complex) live in the
 * warp's shared memory.  The loop runs over SAMPLES, not cells:
 *
 *   - the spokes that cross the tile come from the angle-sorted tables through the tile's angular-bin window
 *     (plan time), each clipped to the run of readout indices inside the tile's box grown by the kernel
 *     half-width -- a contiguous piece of memory, staged by ONE cp.async.bulk per spoke (mbarrier completion,
 *     double buffered, the next slice's runs in flight while this one is consumed);
 *   - a sample (spoke, r) at (px, py) = (r ct, r st) taps the cells with |px - X| < W and |py - Y| < W: for
 *     W = 2 at most 4 x 4 cells.  Sixteen lanes take one sample: lane (lx, ly) derives ITS cell from
 *     floor(px), floor(py), evaluates the reference's predicate for it (single FFMA.FTZ, annulus of the cell,
 *     r = 0 counted twice), the Kaiser-Bessel weight, and adds w * sample into its cell's accumulators
 *     (LDS.128, FFMA2, STS.128).  The two half-warps take samples of the same spoke at least six readout
 *     steps apart, whose supports are disjoint: no two lanes of an instruction touch the same cell.
 *   - sliding golden-angle windows: slice z + 1 = slice z - (spokes that left) + (spokes that entered); the
 *     warp walks a chain of consecutive slices, keeps the running tile in shared memory and stores it after
 *     every slice (coalesced 128-byte rows per coil plane).  The first slice of a chain is gridded in full.
 *
 * Candidate cells of a sample.  The reference's predicate |fma(ct, r, -X)| < W can only hold for integers X in
 * the open interval (px - W, px + W), px the exact product: for W = 2 that is {c, .., c + 3}, c = floor(px) - 1.
 * The kernel has px' = RN(ct r), and floor(px') is floor(px) or floor(px) + 1 (rounding is monotone and integers
 * are representable), so x0 = floor(px') - 1 is c or c + 1: lanes 0..2 test x0 .. x0 + 2, lane 3 tests x0 + 3
 * and, if that fails, x0 - 1 (the two are four apart, at most one can pass).  Every candidate is decided by the
 * reference's own expression; the tap set is identical by construction (tests: indicator probes against the
 * reference kernel, and the gather kernels).
 *
 * Tiles next to DC are crossed by every spoke: there the four warps of a block share ONE tile and chain, each
 * taking every fourth spoke into its own copy of the tile; the copies are added when the slice is stored.
 */
#include "grid_common.cuh"
#include <algorithm>
#include <stdlib.h>
#include <vector>

namespace tronb {

#define SC_T 16                 /* tile edge (cells) */
#define SC_SLOTS 32             /* spokes per staging round (one per lane) */
#define SC_WARPS 4
#ifndef SC_UNROLL
#define SC_UNROLL 2             /* steps per trip of the sample loop */
#endif
#define SC_HDR 2112             /* bytes of per-warp bookkeeping ahead of the tile: barriers, round info, descriptors, annuli */

/* Tile geometry and shared-memory layout per channel count.
 *  nc <= 6:  16 x 16 tiles, a lane holds all nc channels of its cell (16-byte units back to back); rows are padded so
 *            that the 16-byte accesses of a quarter warp (2 rows x 4 cells) fall into 8 different bank groups whatever
 *            the sample's position: row stride = 4 (mod 8) units, 1 (mod 8) for nc = 4.
 *  nc = 16:  16 x 8 tiles; BOTH half-warps take the same sample, half h holds channels 8h .. 8h+7 of its cell
 *            (SHARE).  A cell is 128 bytes = all 8 bank groups, so its units are swizzled:
 *            physical unit = logical ^ (cx & 7) ^ ((cy & 1) << 2) -- a quarter warp's 2 x 4 cells then touch 8
 *            different groups for every logical unit. */
template <int CH> struct ScPlane {
    static constexpr bool SHARE = CH > 8;
    static constexpr int TH = SHARE ? 8 : 16;               /* tile height (rows) */
    static constexpr int LCH = SHARE ? CH / 2 : CH;         /* channels per lane */
    static constexpr int CELL = CH * 8;
    static constexpr int UNITS = SC_T * CH / 2;
    static constexpr int WANT = CH == 4 ? 1 : 4;
    static constexpr int PAD = SHARE ? 0 : (WANT + 8 - UNITS % 8) % 8;
    static constexpr int ROW = (UNITS + PAD) * 16;
    static constexpr int BYTES = ROW * TH;
    /* byte offset inside the tile of 16-byte unit `u` (0 .. LCH/2 - 1) of the channels half `hf` holds in cell (cx, cy) */
    __device__ static __forceinline__ unsigned unit(int cx, int cy, int hf, int u)
    {
        if (!SHARE) return (unsigned)(cy * ROW + cx * CELL + 16 * u);
        const int sw = (cx & 7) ^ ((cy & 1) << 2);
        return (unsigned)(cy * ROW + cx * CELL + 16 * (((hf * (LCH / 2) + u) ^ sw) & 7));
    }
};

/* ---------------------------------------------------------------------- */
/* plan-time tables: one per slice, sorted by angle mod pi                 */
/* ---------------------------------------------------------------------- */
/* kind 0 (full):  the `ne` = win spokes of slice `tab`'s window;
 * kind 1 (delta): the 2 * slide spokes by which slice `tab` differs from slice tab - 1: `slide` leaving
 *                 (sign bit set: their taps are subtracted) and `slide` entering.
 * entry = (cos, sin, bits: spoke index relative to the slice's first spoke (negative for leaving spokes),
 *          bits: 1 | sign << 31, or 0 for an entry that must be skipped). */
__global__ void scatter_table_kernel(float4 *gx, int *lut, float *scratch, int ne, int kind, int per_slice, int skip,
                                     int golden, int nbins, int win, int slide, int nslices)
{
    const int tab = blockIdx.x;
    float *ku = scratch + (size_t)tab * 2 * ne, *ks = ku + ne;
    float4 *gxd = gx + (size_t)tab * 2 * ne;
    const int zskip = skip + (per_slice ? tab * slide : 0);
    for (int e = threadIdx.x; e < ne; e += blockDim.x) {
        const int pe = kind == 0 ? e : (e < slide ? e - slide : win - 2 * slide + e);
        const float t = ref_angle_grid(pe, win, zskip, golden);
        float key = fmodf(t, PI_F);
        if (key < 0.f) key += PI_F;
        if (key >= PI_F) key -= PI_F;
        ku[e] = key;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < ne; e += blockDim.x) {
        const float key = ku[e];
        int rank = 0;
        for (int j = 0; j < ne; ++j) { const float kj = ku[j]; rank += (kj < key) || (kj == key && j < e); }
        const int pe = kind == 0 ? e : (e < slide ? e - slide : win - 2 * slide + e);
        const float t = ref_angle_grid(pe, win, zskip, golden);
        int code = 1 | ((kind == 1 && e < slide) ? (int)0x80000000 : 0);
        if (kind == 1 && tab == 0) code = 0;               /* no previous slice */
        gxd[rank] = gxd[rank + ne] = make_float4(cos_approx(t), sin_approx(t), __int_as_float(pe), __int_as_float(code));
        ks[rank] = key;
    }
    __syncthreads();
    const float lut_scale = (float)nbins / PI_F;
    int *l = lut + (size_t)tab * (nbins + 1);
    for (int b = threadIdx.x; b <= nbins; b += blockDim.x) {
        int lo = 0, hi = ne;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (angle_bin(ks[mid], lut_scale, nbins) >= b) hi = mid; else lo = mid + 1;
        }
        l[b] = lo;
    }
    (void)nslices;
}

static int build_scatter_table(float4 **gx, int **lut, int ntab, int ne, int kind, int per_slice, int skip, int golden,
                               int nbins, int win, int slide, int nslices, cudaStream_t s)
{
    float *scratch = nullptr;
    TRON_CUDA(cudaMalloc(gx, (size_t)ntab * 2 * ne * sizeof(float4)));
    TRON_CUDA(cudaMalloc(lut, (size_t)ntab * (nbins + 1) * sizeof(int)));
    TRON_CUDA(cudaMalloc(&scratch, (size_t)ntab * 2 * ne * sizeof(float)));
    scatter_table_kernel<<<ntab, 256, 0, s>>>(*gx, *lut, scratch, ne, kind, per_slice, skip, golden, nbins, win, slide, nslices);
    TRON_CUDA(cudaGetLastError());
    TRON_CUDA(cudaStreamSynchronize(s));
    TRON_CUDA(cudaFree(scratch));
    return 0;
}

void scatter_plan_free(ScatterPlan &sp)
{
    cudaFree(sp.tab_full); cudaFree(sp.lut_full); cudaFree(sp.tab_delta); cudaFree(sp.lut_delta);
    cudaFree(sp.tile_win); cudaFree(sp.sched); cudaFree(sp.sched_short);
    sp = ScatterPlan();
}

/* tile schedule: nearest DC first; tiles whose angular window holds at least `near_frac` of all spokes are
 * "near": one block (4 warps splitting the spokes) per tile and short chain.  order = near | far | empty. */
static int build_scatter_schedule(int **d_sched, int *n_near, int *n_far, int *n_empty, const int2 *d_tile_win, int n,
                                  int nbins, float W, float near_frac, int th)
{
    const int nt1 = (n + SC_T - 1) / SC_T, nty = (n + th - 1) / th, nt = nt1 * nty;
    std::vector<int2> hw(nt);
    TRON_CUDA(cudaMemcpy(hw.data(), d_tile_win, nt * sizeof(int2), cudaMemcpyDeviceToHost));
    const float rz = (float)(n / 2 - 1) + W + 0.5f;            /* beyond: no cell can hold a sample */
    std::vector<std::pair<float, int>> nearv, farv;
    std::vector<int> empty;
    for (int t = 0; t < nt; ++t) {
        const int x0 = (t % nt1) * SC_T - n / 2, y0 = (t / nt1) * th - n / 2;
        const float dx = x0 > 0 ? (float)x0 : (x0 + SC_T - 1 < 0 ? (float)-(x0 + SC_T - 1) : 0.f);
        const float dy = y0 > 0 ? (float)y0 : (y0 + th - 1 < 0 ? (float)-(y0 + th - 1) : 0.f);
        const float d2 = dx * dx + dy * dy;
        const int packed = ((t / nt1) << 16) | (t % nt1);
        const int2 w = hw[t];
        if (w.x != CELL_ALL_SPOKES && w.y < w.x) {             /* no cell of the tile is ever tapped: only zeros to store */
            if (d2 <= rz * rz) farv.push_back(std::make_pair(d2, packed));
            else empty.push_back(packed);                      /* beyond the last annulus: visited only when every cell is stored */
            continue;
        }
        const float frac = w.x == CELL_ALL_SPOKES ? 1.f : (float)(w.y - w.x + 1) / (float)nbins;
        if (frac >= near_frac) nearv.push_back(std::make_pair(d2, packed)); else farv.push_back(std::make_pair(d2, packed));
    }
    std::sort(nearv.begin(), nearv.end());
    std::sort(farv.begin(), farv.end());
    std::vector<int> order;
    for (size_t i = 0; i < nearv.size(); ++i) order.push_back(nearv[i].second);
    for (size_t i = 0; i < farv.size(); ++i) order.push_back(farv[i].second);
    order.insert(order.end(), empty.begin(), empty.end());
    *n_near = (int)nearv.size(); *n_far = (int)farv.size(); *n_empty = (int)empty.size();
    TRON_CUDA(cudaMalloc(d_sched, order.size() * sizeof(int)));
    TRON_CUDA(cudaMemcpy(*d_sched, order.data(), order.size() * sizeof(int), cudaMemcpyHostToDevice));
    return 0;
}

/* `cells` / `nbins`: the per-cell table of the plan's spoke tables (slice independent), shared with the gather kernels */
int scatter_plan_build(ScatterPlan &sp, const int2 *cells, int nbins, int n, int nslices, int win, int slide, int skip,
                       int golden, float W, int nc, cudaStream_t s)
{
    sp = ScatterPlan();
    sp.win = win;
    sp.th = nc > 8 ? 8 : 16;                                /* ScPlane<nc>::TH */
    /* difference tables pay when a slice's difference is clearly fewer spokes than its window */
    const bool sliding = golden && nslices > 1 && 4 * slide <= win;
    sp.per_slice = golden ? 1 : 0;
    const int ntab = sp.per_slice ? nslices : 1;
    int rc = build_scatter_table(&sp.tab_full, &sp.lut_full, ntab, win, 0, sp.per_slice, skip, golden, nbins, win, slide, nslices, s);
    if (rc) return rc;
    const int env_chain = getenv("TRON_SCATTER_CHAIN") ? atoi(getenv("TRON_SCATTER_CHAIN")) : 0;
    const int env_chain_near = getenv("TRON_SCATTER_CHAIN_NEAR") ? atoi(getenv("TRON_SCATTER_CHAIN_NEAR")) : 0;
    sp.chain = 1; sp.chain_near = 1; sp.ne_delta = 0;
    if (sliding && env_chain != 1) {
        sp.ne_delta = 2 * slide;
        rc = build_scatter_table(&sp.tab_delta, &sp.lut_delta, nslices, sp.ne_delta, 1, 1, skip, golden, nbins, win, slide, nslices, s);
        if (rc) return rc;
        /* long jobs (launches of ~500 slices): chains of 64 -- 44.6 instead of 47.1 spoke visits per slice, and still
         * enough (tile, chain) tasks; shorter jobs keep 32 (256-slice launches: 1.30 vs 1.08 ms with 64) */
        sp.chain = env_chain > 1 ? env_chain : (nslices >= 448 ? 64 : 32);
        sp.chain_near = env_chain_near > 0 ? env_chain_near : 16;
        if (sp.chain_near > sp.chain) sp.chain_near = sp.chain;
        while (sp.chain % sp.chain_near) --sp.chain_near;      /* near chains nest in the far ones */
    }
    rc = build_tile_windows(&sp.tile_win, cells, n, nbins, SC_T, sp.th, s);
    if (rc) return rc;
    const float near_frac = getenv("TRON_SCATTER_NEAR") ? (float)atof(getenv("TRON_SCATTER_NEAR")) : 0.25f;
    const float near_frac_short = getenv("TRON_SCATTER_NEAR_SHORT") ? (float)atof(getenv("TRON_SCATTER_NEAR_SHORT")) : 0.1f;
    TRON_CUDA(cudaStreamSynchronize(s));
    rc = build_scatter_schedule(&sp.sched, &sp.n_near, &sp.n_far, &sp.ntiles_empty, sp.tile_win, n, nbins, W, near_frac, sp.th);
    if (rc) return rc;
    int nempty = 0;
    rc = build_scatter_schedule(&sp.sched_short, &sp.n_near_short, &sp.n_far_short, &nempty, sp.tile_win, n, nbins, W, near_frac_short, sp.th);
    if (rc) return rc;
    sp.chain_short = sp.chain; sp.chain_near_short = sp.chain_near;
    if (sp.chain > 1) {
        const int cs = getenv("TRON_SCATTER_CHAIN_SHORT") ? atoi(getenv("TRON_SCATTER_CHAIN_SHORT")) : 16;
        const int cn = getenv("TRON_SCATTER_CHAIN_NEAR_SHORT") ? atoi(getenv("TRON_SCATTER_CHAIN_NEAR_SHORT")) : 8;
        sp.chain_short = cs > 0 && cs < sp.chain ? cs : sp.chain;
        while (sp.chain % sp.chain_short) --sp.chain_short;          /* launches start on long-chain boundaries */
        sp.chain_near_short = cn > 0 && cn < sp.chain_short ? cn : sp.chain_short;
        while (sp.chain_short % sp.chain_near_short) --sp.chain_near_short;
    }
    sp.short_below = getenv("TRON_SCATTER_SHORT_BELOW") ? atoi(getenv("TRON_SCATTER_SHORT_BELOW")) : 160;
    sp.ready = 1;
    return 0;
}

/* ---------------------------------------------------------------------- */
/* the kernel                                                              */
/* ---------------------------------------------------------------------- */
__device__ __forceinline__ void sts_f4(unsigned a, float4 q)
{
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(q.x), "f"(q.y), "f"(q.z), "f"(q.w) : "memory");
}
/* the store happens for lanes with `on` only (dead lanes run along branch free) */
__device__ __forceinline__ void sts_f4_if(unsigned a, float4 q, bool on)
{
    asm volatile("{\n.reg .pred pp;\nsetp.ne.s32 pp, %5, 0;\n@pp st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n}"
                 ::"r"(a), "f"(q.x), "f"(q.y), "f"(q.z), "f"(q.w), "r"((int)on) : "memory");
}
__device__ __forceinline__ unsigned lds_u1(unsigned a)
{
    unsigned q;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(q) : "r"(a));
    return q;
}
__device__ __forceinline__ void bar_sync_block(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

struct ScatCursor { int z, slot, k0, cnt; };

enum { SC_FIRST = 1, SC_LAST = 2, SC_FINAL = 4, SC_FULL = 8 };

/* One task: tile (tx, ty), plan-local slices [zs, ze) as one chain (zs is gridded in full), this warp taking the
 * table entries sub, sub + nsub, ... of the tile's window.  NSUB = 1: the warp owns the tile and stores it;
 * NSUB = SC_WARPS: the block's warps hold partial tiles that are added when a slice is stored. */
template <int CH, bool HALF, int NSUB>
__device__ __forceinline__ void scatter_task(const GridLaunch &g, const ScatterPlan &sp, unsigned char *smem_raw, const int cap,
                                             const int tile, const int zs, const int ze, const int store_all)
{
    using P = ScPlane<CH>;
    constexpr unsigned SAMP = CH * (HALF ? 4u : 8u);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = NSUB > 1 ? warp : 0;
    const int n = g.n;
    const unsigned per_warp = (unsigned)(SC_HDR + P::BYTES + 2 * cap);
    unsigned char *wb = smem_raw + (size_t)warp * per_warp;
    const unsigned w0 = smem_u32(wb);
    const unsigned bar0 = w0;                              /* 2 x 8 bytes */
    int4 *info_p = reinterpret_cast<int4 *>(wb + 16);      /* 2 x 16 bytes */
    const unsigned seg0 = w0 + 64;                         /* 2 x 32 x 16 bytes */
    const unsigned ann0 = w0 + 64 + 1024;                  /* 256 x 4 bytes: Rlo | Rhi << 16 of the tile's cells */
    const unsigned plane = w0 + SC_HDR;
    const unsigned data0 = plane + P::BYTES;

    constexpr int TH = P::TH, LCH = P::LCH;
    const int x0 = (tile & 0xffff) * SC_T, y0 = (tile >> 16) * TH;
    const int XL = x0 - n / 2, YL = y0 - n / 2;
    const int XH = min(x0 + SC_T - 1, n - 1) - n / 2, YH = min(y0 + TH - 1, n - 1) - n / 2;
    const int nt1 = (n + SC_T - 1) / SC_T;
    const int2 tw = __ldg(sp.tile_win + (size_t)(tile >> 16) * nt1 + (tile & 0xffff));

    /* annuli of the tile's cells (tron.cu:498-502), slice independent */
    for (int i = lane; i < SC_T * TH; i += 32) {
        const int x = x0 + (i & 15), y = y0 + (i >> 4);
        unsigned a = 1u;                                   /* Rlo = 1 > Rhi = 0: never tapped */
        if (x < n && y < n) a = (unsigned)__ldg(&g.cells[(size_t)y * n + x].x);
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(ann0 + 4u * (unsigned)i), "r"(a) : "memory");
    }
    if (lane == 0) {
        mbar_init(bar0, 1); mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    const size_t esz = HALF ? sizeof(__half2) : sizeof(float2);
    const int half_nro = g.nro >> 1;
    const float W = g.kb.W;
    const float Rmax = (float)(n / 2 - 1);

    /* ---- staging: the next round of the cursor into buffer `buf` ---- */
    ScatCursor cur; cur.z = zs; cur.slot = 0; cur.k0 = 0; cur.cnt = -1;
    auto stage = [&](int buf) {
        if (cur.z >= ze) return;
        const bool full = cur.z == zs;
        const int ne = full ? sp.win : sp.ne_delta;
        const int tabi = sp.per_slice ? cur.z : 0;
        if (cur.cnt < 0) {
            const TileWindow w = window_of(tw, (full ? sp.lut_full : sp.lut_delta) + (size_t)tabi * (g.nbins + 1), g.nbins, ne);
            cur.k0 = w.k0; cur.cnt = w.cnt;
        }
        const float4 *tab = (full ? sp.tab_full : sp.tab_delta) + (size_t)tabi * 2 * ne + cur.k0;
        const char *samples = (const char *)g.samples + (size_t)cur.z * g.slide * g.nro * g.nc_total * esz;
        const int slot = cur.slot + lane * NSUB + sub;
        const bool valid = slot < cur.cnt;
        const float4 e = valid ? __ldg(tab + slot) : make_float4(1.f, 1.f, 0.f, 0.f);
        float ict, ist, hwx, hwy;
        axis_terms(e.x, W, ict, hwx);
        axis_terms(e.y, W, ist, hwy);
        /* readout indices r with (r ct, r st) inside the tile's box grown by W (+ margin) */
        const float cx = 0.5f * (float)(XL + XH) * ict, cy = 0.5f * (float)(YL + YH) * ist;
        const float ex = fmaf(0.5f * (float)(XH - XL), fabsf(ict), hwx + 0.05f);
        const float ey = fmaf(0.5f * (float)(YH - YL), fabsf(ist), hwy + 0.05f);
        const float lo = fmaxf(fmaxf(cx - ex, cy - ey), -Rmax);
        const float hi = fminf(fminf(cx + ex, cy + ey), Rmax);
        int ra = 0, rb = -1;
        const int code = __float_as_int(e.w);
        if (valid && code != 0 && lo <= hi) { ra = (int)ceilf(lo); rb = (int)floorf(hi); }
        const char *src = samples + ((ptrdiff_t)__float_as_int(e.z) * g.nro + half_nro + ra) * (ptrdiff_t)(g.nc_total * esz);
        unsigned bytes = 0, lead = 0;
        if (rb >= ra) {
            lead = (unsigned)((uintptr_t)src & 15);
            bytes = (lead + (unsigned)(rb - ra + 1) * SAMP + 15u) & ~15u;
        }
        unsigned incl = bytes;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const unsigned fits = __ballot_sync(0xffffffffu, valid && incl <= (unsigned)cap);
        const int count = __popc(fits);                     /* a prefix of the lanes: incl is monotone */
        const unsigned bar = bar0 + 8 * buf;
        const unsigned dst = data0 + (unsigned)buf * (unsigned)cap + (incl - bytes);
        if (lane < count) {
            if (bytes) bulk_g2s(dst, src - lead, bytes, bar);
            const int len = bytes ? rb - ra + 1 : 0;
            const int pk = ((ra + 2048) & 0xfff) | (len << 12) | (code & (int)0x80000000);
            sts_f4(seg0 + (unsigned)(buf * SC_SLOTS + lane) * 16u,
                   make_float4(e.x, e.y, __int_as_float((int)(dst + lead) - ra * (int)SAMP), __int_as_float(pk)));
        }
        const unsigned total = __shfl_sync(0xffffffffu, incl, count > 0 ? count - 1 : 0);
        const bool last = cur.slot + count * NSUB + sub >= cur.cnt;      /* this warp's next entry would lie beyond the window */
        if (lane == 0)
            info_p[buf] = make_int4(cur.z, count, (cur.slot == 0 ? SC_FIRST : 0) | (last ? SC_LAST : 0)
                                    | (last && cur.z + 1 >= ze ? SC_FINAL : 0) | (full ? SC_FULL : 0), 0);
        __syncwarp();
        if (lane == 0) mbar_arrive_expect_tx(bar, count > 0 ? total : 0u);
        cur.slot += count * NSUB;
        if (last) { cur.z += 1; cur.slot = 0; cur.cnt = -1; }
    };

    stage(0);

    unsigned long long c2[TRONB_KB_DEG + 1];
#pragma unroll
    for (int m = 0; m <= TRONB_KB_DEG; ++m) c2[m] = *reinterpret_cast<const unsigned long long *>(&g.kb.c2[m]);
    float invW = g.kb.invW, sdc_as = g.sdc_as, sdc_bs = g.sdc_bs, Wk = g.kb.W;
    asm volatile("" : "+f"(invW), "+f"(sdc_as), "+f"(sdc_bs), "+f"(Wk));
#pragma unroll
    for (int m = 0; m <= TRONB_KB_DEG; ++m) asm volatile("" : "+l"(c2[m]));      /* in registers, not re-read from the constant bank per step */

    /* this lane's place in a sample's 4 x 4 support, and in the pair of samples a warp takes per step */
    const int lx = lane & 3, ly = (lane >> 2) & 3, hf = lane >> 4;
    const float lxf = (float)(lx - 1), lyf = (float)(ly - 1);
    const int tX0 = XL, tY0 = YL;

    for (int it = 0;; ++it) {
        const int buf = it & 1;
        stage(buf ^ 1);                                     /* (every lane finished reading that buffer: __syncwarp below) */
        mbar_wait(bar0 + 8 * buf, (it >> 1) & 1);
        const int4 info = info_p[buf];
        if ((info.z & (SC_FIRST | SC_FULL)) == (SC_FIRST | SC_FULL)) {        /* a chain starts: clear the tile */
            for (unsigned o = 16u * (unsigned)lane; o < (unsigned)P::BYTES; o += 512u) sts_f4(plane + o, make_float4(0.f, 0.f, 0.f, 0.f));
            __syncwarp();
        }
        const unsigned seg = seg0 + (unsigned)(buf * SC_SLOTS) * 16u;
#pragma unroll 1
        for (int j = 0; j < info.y; ++j) {
            const float4 d = lds_f4(seg + 16u * (unsigned)j);
            const float ct = d.x, st = d.y;
            const int a0 = __float_as_int(d.z), pk = __float_as_int(d.w);
            const int len = (pk >> 12) & 0xfff;
            if (len == 0) continue;
            const int ra = (pk & 0xfff) - 2048;
            const int sgn = pk & (int)0x80000000;
            /* the two half-warps take samples >= 6 readout steps apart (disjoint supports); short runs: one half.
             * SHARE (nc = 16): both halves take the same sample, each its own eight channels */
            const bool two = !P::SHARE && len >= 12;
            const int h = two ? (len + 1) >> 1 : len;
            int r = ra + ((hf && two) ? h : 0);
            const int mine = P::SHARE ? len : (hf ? (two ? len - h : 0) : h);
            /* one step = one sample per half-warp; branch free (dead lanes compute along and skip the store), two
             * steps per trip so that the second's index arithmetic and weight overlap the first's shared-memory
             * round trip */
            const int rsafe = ra;                           /* a staged sample for lanes without one of their own */
#pragma unroll 1
            for (int i = 0; i < h; i += SC_UNROLL, r += SC_UNROLL) {
#pragma unroll
                for (int u = 0; u < SC_UNROLL; ++u) {
                    const bool live = i + u < mine;
                    const int re = live ? r + u : rsafe;
                    const float rf = (float)re;
                    const float px = mul_ftz(ct, rf), py = mul_ftz(st, rf);
                    float Xf = floorf(px) + lxf, Yf = floorf(py) + lyf;
                    float dx = fma_ftz(ct, rf, -Xf);                     /* tron.cu:514,516 as compiled */
                    float dy = fma_ftz(st, rf, -Yf);
                    if (lx == 3 && !(fabsf(dx) < Wk)) { Xf -= 4.f; dx = fma_ftz(ct, rf, -Xf); }
                    if (ly == 3 && !(fabsf(dy) < Wk)) { Yf -= 4.f; dy = fma_ftz(st, rf, -Yf); }
                    const int cxl = (int)Xf - tX0, cyl = (int)Yf - tY0;
                    bool ok = live && fabsf(dx) < Wk && fabsf(dy) < Wk && (unsigned)cxl < (unsigned)SC_T && (unsigned)cyl < (unsigned)TH;
                    const int cxc = cxl & (SC_T - 1), cyc = cyl & (TH - 1);
                    const unsigned ann = lds_u1(ann0 + 4u * (unsigned)(cyc * SC_T + cxc));
                    const int ar = abs(re);
                    ok = ok && ar >= (int)(ann & 0xffffu) && ar <= (int)(ann >> 16);   /* annulus, tron.cu:501-502,512,521 */
                    float w = kb_poly_xy_c2(dx, dy, invW, c2);
                    const float sdc = fmaf(sdc_as, fabsf(rf), sdc_bs);       /* tron.cu:412, times the output scale */
                    w *= (re == 0) ? sdc + sdc : sdc;                        /* both of the reference's loops visit r = 0 */
                    w = __int_as_float(__float_as_int(w) ^ sgn);             /* leaving spoke: subtract */
                    const unsigned sa = (unsigned)(a0 + re * (int)SAMP) + (P::SHARE ? (unsigned)hf * (unsigned)(LCH * (HALF ? 4 : 8)) : 0u);
                    float2 v[LCH];
                    lds_sample<LCH, HALF>(v, sa);
#pragma unroll
                    for (int c = 0; c < LCH / 2; ++c) {
                        const unsigned ca = plane + P::unit(cxc, cyc, hf, c);
                        float4 q = lds_f4(ca);
                        float2 q0 = make_float2(q.x, q.y), q1 = make_float2(q.z, q.w);
                        ffma2(q0, w, v[2 * c]); ffma2(q1, w, v[2 * c + 1]);
                        sts_f4_if(ca, make_float4(q0.x, q0.y, q1.x, q1.y), ok);
                    }
#ifdef SC_STEP_SYNC
                    __syncwarp();                           /* the next step's samples may tap the same cells */
#endif
                    /* (no barrier between steps by default: the warp is converged here, its shared-memory instructions
                     * are volatile and execute in program order, so the next step's loads see this step's stores) */
                }
            }
            __syncwarp();
        }
        if (info.z & SC_LAST) {
            /* the slice is complete: store the tile, one 128-byte row segment per coil plane and row */
            const int zl = info.x - g.z0;                   /* slice index inside this launch */
            const size_t plane_sz = (size_t)n * n;
            if (NSUB == 1) {
                __syncwarp();
                const int cx = lane & 15, x = x0 + cx, X = x - n / 2;
                const int cy0 = lane >> 4;
                const size_t pstride = plane_sz * sizeof(float2), rowstep = (size_t)2 * n * sizeof(float2);
                char *obase = (char *)(g.grid + (size_t)zl * g.nch * plane_sz + (size_t)(y0 + cy0) * n + x);
                /* which of this lane's rows are stored (tile inside the grid and the last annulus: all of them) */
                unsigned rows = 0;
#pragma unroll
                for (int k = 0; k < TH / 2; ++k) {
                    const int y = y0 + cy0 + 2 * k, Y = y - n / 2;
                    if (x < n && y < n && (store_all || X * X + Y * Y <= g.zero_r2)) rows |= 1u << k;
                }
#pragma unroll
                for (int c = 0; c < CH / 2; ++c) {          /* c-th pair of coil planes */
                    char *oa = obase + (size_t)(2 * c) * pstride, *ob = oa + pstride;
#pragma unroll 4
                    for (int k = 0; k < TH / 2; ++k) {
                        const float4 q = lds_f4(plane + P::unit(cx, cy0 + 2 * k, c / (LCH / 2), c % (LCH / 2)));
                        if (rows & (1u << k)) { __stcs((float2 *)oa, make_float2(q.x, q.y)); __stcs((float2 *)ob, make_float2(q.z, q.w)); }
                        oa += rowstep; ob += rowstep;
                    }
                }
            } else {
                bar_sync_block(1, SC_WARPS * 32);           /* every warp's partial tile of this slice is complete */
                const int t = threadIdx.x;
#pragma unroll 1
                for (int k = 0; k < TH / 8; ++k) {
                    const int cy = 8 * k + (t >> 4), cx = t & 15;
                    const int x = x0 + cx, y = y0 + cy;
                    const int X = x - n / 2, Y = y - n / 2;
                    const bool st_ok = x < n && y < n && (store_all || X * X + Y * Y <= g.zero_r2);
                    float2 *o = g.grid + (size_t)zl * g.nch * plane_sz + (size_t)y * n + x;
#pragma unroll
                    for (int c = 0; c < CH / 2; ++c) {
                        const unsigned off = (unsigned)SC_HDR + P::unit(cx, cy, c / (LCH / 2), c % (LCH / 2));
                        float4 q = lds_f4(smem_u32(smem_raw) + off);
#pragma unroll
                        for (int ww = 1; ww < SC_WARPS; ++ww) {
                            const float4 p = lds_f4(smem_u32(smem_raw) + (unsigned)ww * per_warp + off);
                            q.x += p.x; q.y += p.y; q.z += p.z; q.w += p.w;
                        }
                        if (st_ok) { __stcs(o + (size_t)(2 * c) * plane_sz, make_float2(q.x, q.y)); __stcs(o + (size_t)(2 * c + 1) * plane_sz, make_float2(q.z, q.w)); }
                    }
                }
                bar_sync_block(1, SC_WARPS * 32);           /* before anyone adds the next slice's taps */
            }
        }
        __syncwarp();                                       /* every lane is done with buffer `buf` */
        if (info.z & SC_FINAL) break;
    }
}

template <int CH, bool HALF>
__global__ void __launch_bounds__(SC_WARPS * 32, 3)
grid_scatter_kernel(const GridLaunch g, const ScatterPlan sp, const int cap, const int nchunk_near, const int nchunk_far,
                    const int store_all)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int zend = g.z0 + g.nslices;
    int b = blockIdx.x;
    const int near_blocks = sp.n_near * nchunk_near;
    if (b < near_blocks) {
        /* tiles next to DC: one block per (tile, short chain), the warps split the spokes */
        const int trank = b / nchunk_near, c = g.z0 / sp.chain_near + b % nchunk_near;
        const int zs = max(c * sp.chain_near, g.z0), ze = min((c + 1) * sp.chain_near, zend);
        if (zs >= ze) return;
        scatter_task<CH, HALF, SC_WARPS>(g, sp, smem_raw, cap, __ldg(sp.sched + trank), zs, ze, store_all);
        return;
    }
    b -= near_blocks;
    /* the other tiles: four per block (neighbours in the nearest-first order: similar work), one warp each */
    const int ntl = sp.n_far + (store_all ? sp.ntiles_empty : 0);
    const int quad = b / nchunk_far, c = g.z0 / sp.chain + b % nchunk_far;
    const int ti = quad * SC_WARPS + (threadIdx.x >> 5);
    if (ti >= ntl) return;
    const int zs = max(c * sp.chain, g.z0), ze = min((c + 1) * sp.chain, zend);
    if (zs >= ze) return;
    scatter_task<CH, HALF, 1>(g, sp, smem_raw, cap, __ldg(sp.sched + sp.n_near + ti), zs, ze, store_all);
}

/* ---------------------------------------------------------------------- */
/* launch                                                                  */
/* ---------------------------------------------------------------------- */
bool grid_scatter_applicable(const GridLaunch &g)
{
    if (getenv("TRON_NO_SCATTER") != nullptr) return false;  /* diagnostic switches are read per launch */
    if (!g.scat || !g.scat->ready) return false;
    if (!(g.kb.fast && g.nro == g.n && g.kb.W == 2.0f)) return false;    /* 4 x 4 supports, ridx = r */
    if (g.nch != g.nc_total || g.ch0 != 0) return false;     /* whole samples are staged */
    if (g.nch != 2 && g.nch != 4 && g.nch != 6 && g.nch != 16) return false;
    if (g.n % SC_T != 0 || g.n > 4096) return false;
    if (g.scat->th != (g.nch > 8 ? 8 : 16)) return false;
    if (((uintptr_t)g.samples) % 16 != 0 && !g.half_in) return false;
    return g.dbg == nullptr;
}

template <int CH, bool HALF>
static int launch_scatter_t(const GridLaunch &g, cudaStream_t s)
{
    using P = ScPlane<CH>;
    ScatterPlan sp = *g.scat;
    if (g.nslices < sp.short_below) {                        /* too few (tile, chain) tasks for the long schedule */
        sp.sched = sp.sched_short; sp.n_near = sp.n_near_short; sp.n_far = sp.n_far_short;
        sp.chain = sp.chain_short; sp.chain_near = sp.chain_near_short;
    }
    const unsigned samp = CH * (HALF ? 4u : 8u);
    /* longest run a spoke can have inside a tile's box: its diagonal (+ margins) */
    const float bw = (float)(SC_T - 1) + 2.f * g.kb.W + 0.2f, bh = (float)(P::TH - 1) + 2.f * g.kb.W + 0.2f;
    const unsigned longest = ((unsigned)(sqrtf(bw * bw + bh * bh) + 3.f) * samp + 31u) & ~15u;
    unsigned cap = getenv("TRON_SCATTER_CAP") ? (unsigned)atoi(getenv("TRON_SCATTER_CAP")) : 1536u;     /* 3 blocks per SM */
    if (cap < longest) cap = longest;
    cap = (cap + 127u) & ~127u;
    const size_t smem = (size_t)SC_WARPS * (SC_HDR + P::BYTES + 2 * (size_t)cap);
    if (smem > 200 * 1024) return -1;
    const int zend = g.z0 + g.nslices;
    const int nchunk_near = (zend - 1) / sp.chain_near - g.z0 / sp.chain_near + 1;
    const int nchunk_far = (zend - 1) / sp.chain - g.z0 / sp.chain + 1;
    const int store_all = g.zero_r2 == 0x7fffffff ? 1 : 0;
    const int ntl = sp.n_far + (store_all ? sp.ntiles_empty : 0);
    const long long blocks = (long long)sp.n_near * nchunk_near + (long long)((ntl + SC_WARPS - 1) / SC_WARPS) * nchunk_far;
    if (blocks > 0x7fffffffLL || blocks < 1) return -1;
    auto kern = grid_scatter_kernel<CH, HALF>;
    static bool attr_set = false;
    if (!attr_set || smem > 48 * 1024) {
        TRON_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    kern<<<(unsigned)blocks, SC_WARPS * 32, smem, s>>>(g, sp, (int)cap, nchunk_near, nchunk_far, store_all);
    TRON_CUDA(cudaGetLastError());
    return 0;
}

int launch_grid_scatter(const GridLaunch &g, cudaStream_t s)
{
    switch (g.nch) {
    case 2: return g.half_in ? launch_scatter_t<2, true>(g, s) : launch_scatter_t<2, false>(g, s);
    case 4: return g.half_in ? launch_scatter_t<4, true>(g, s) : launch_scatter_t<4, false>(g, s);
    case 6: return g.half_in ? launch_scatter_t<6, true>(g, s) : launch_scatter_t<6, false>(g, s);
    case 16: return g.half_in ? launch_scatter_t<16, true>(g, s) : launch_scatter_t<16, false>(g, s);
    }
    return -1;
}

} // namespace tronb
