/*
 * grid_tile.cu -- gridding with the samples staged through shared memory (TMA bulk copies).
 *
 * Same operator, same tap set as grid.cu (precompensate + gridradial2d, /root/reference/src/tron.cu:405-416
 * and 465-536); what changes is where a tap's sample comes from and who waits for whom.
 *
 * grid.cu reads every tap through L1 (`ld.global.nc`, one 16-byte request per lane and two coils): ncu shows
 * its warps waiting on those loads 41 % of the time (profiles/r01_ncu_grid_v12.txt).  Here every WARP owns a
 * footprint of 8 x 4 cells for several consecutive slice groups and runs its own copy pipeline:
 *
 *   1. the warp walks the footprint's window of the angle-sorted spoke table (plan-time footprint table: the
 *      union of its cells' windows), one spoke per lane, clips the spoke against the footprint's bounding box
 *      (+ kernel half-width) and gets the run of readout indices [ra, rb] that any of its cells could tap --
 *      a CONTIGUOUS piece of the spoke in memory (nudata[nchan*(nro*pe + ro) + ch], tron.cu:519);
 *   2. each lane issues ONE `cp.async.bulk` (global -> shared, completion on the warp's mbarrier) for its
 *      spoke's run and writes a 32-byte descriptor (cos, sin, 1/cos, 1/sin, half-widths, shared address, mask);
 *   3. the lanes then walk their own cells' windows as in grid.cu -- same candidates, same reference
 *      predicates -- but read descriptors and samples from shared memory (`ld.shared.v4`).
 *
 * Rounds (at most 32 spokes and `cap` bytes) are double buffered per warp: round k+1 is in flight while round
 * k is consumed, across group boundaries, so the copy latency is paid once per warp.  The warps of a block
 * never synchronise with each other (a first version with one pipeline per 16 x 8 tile and a block barrier per
 * round spent 40 % of its warp time at that barrier: profiles/r02_ncu_tile_v0.txt).  Tiles next to DC
 * (hundreds of spokes per window) take one group per block, the others `gper` groups.
 */
#include "grid_common.cuh"
#include <algorithm>
#include <stdlib.h>
#include <vector>

namespace tronb {

#define FOOT_W 8
#define FOOT_H 4

/* ---------------------------------------------------------------------- */
/* plan-time: per-footprint angular window = union of its cells' windows   */
/* ---------------------------------------------------------------------- */
__global__ void foot_window_kernel(int2 *win, const int2 *cells, int n, int nbins, int nfx, int nfy, int fw, int fh)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nfx * nfy) return;
    const int x0 = (t % nfx) * fw, y0 = (t / nfx) * fh;
    bool any = false, all = false;
    int ref = 0, lo = 0, hi = 0;
    for (int y = y0; y < y0 + fh && y < n; ++y)
        for (int x = x0; x < x0 + fw && x < n; ++x) {
            const int2 c = cells[(size_t)y * n + x];
            const int Rlo = c.x & 0xffff, Rhi = c.x >> 16;
            if (Rlo > Rhi) continue;
            const int lo16 = c.y & 0xffff;
            if (lo16 == CELL_ALL_SPOKES) { all = true; continue; }
            int b0 = (int)(short)lo16, b1 = c.y >> 16;
            if (!any) { any = true; ref = (b0 + b1) / 2; lo = b0; hi = b1; continue; }
            int mid = (b0 + b1) / 2;
            while (mid - ref > nbins / 2) { mid -= nbins; b0 -= nbins; b1 -= nbins; }
            while (ref - mid > nbins / 2) { mid += nbins; b0 += nbins; b1 += nbins; }
            lo = min(lo, b0); hi = max(hi, b1);
        }
    int2 w;
    if (all || (any && hi - lo + 1 >= nbins)) w = make_int2(CELL_ALL_SPOKES, 0);
    else if (!any) w = make_int2(0, -1);                       /* no cell of the footprint can receive a sample */
    else {
        while (lo >= nbins) { lo -= nbins; hi -= nbins; }
        while (lo < 0) { lo += nbins; hi += nbins; }
        w = make_int2(lo, hi);
    }
    win[t] = w;
}

/* windows of fw x fh footprints (8 x 4: one warp of grid_tile.cu; 16 x 16: one tile of grid_scatter.cu) */
int build_tile_windows(int2 **d_win, const int2 *cells, int n, int nbins, int fw, int fh, cudaStream_t s)
{
    const int nfx = (n + fw - 1) / fw, nfy = (n + fh - 1) / fh;
    TRON_CUDA(cudaMalloc(d_win, (size_t)nfx * nfy * sizeof(int2)));
    foot_window_kernel<<<(nfx * nfy + 127) / 128, 128, 0, s>>>(*d_win, cells, n, nbins, nfx, nfy, fw, fh);
    TRON_CUDA(cudaGetLastError());
    return 0;
}

/* ---------------------------------------------------------------------- */
/* plan-time: sliding-window difference tables                              */
/* ---------------------------------------------------------------------- */
/* Golden angles, windows of `win` spokes sliding by `slide`: slice z + 1 holds the spokes of slice z minus
 * the `slide` oldest plus the `slide` next ones, and a tap's weight depends on the absolute spoke index only
 * (tron.cu:509,630).  So   grid(z+1) = grid(z) - G(leaving spokes) + G(entering spokes).
 * Table `tab` lists, for slice group `tab` (gs slices), the 2*gs*slide spokes that enter or leave between
 * consecutive slices from the last slice of group tab-1 on, sorted by angle mod pi like the full tables:
 *   entry = (cos, sin, bits: spoke index relative to the group's first spoke (may be negative),
 *            bits: 1 << k | sign << 31)     k = slice of the group whose difference the spoke belongs to,
 *                                           sign set = the spoke LEAVES (its taps are subtracted). */
__global__ void delta_table_kernel(float4 *gx, int *lut, float *key_unsorted, float *key_sorted, int ne,
                                   int tab_stride, int skip, int nbins, int win, int slide, int gs, int nslices)
{
    const int tab = blockIdx.x;
    float *ku = key_unsorted + (size_t)tab * ne, *ks = key_sorted + (size_t)tab * ne;
    float4 *gxd = gx + (size_t)tab * 2 * ne;
    for (int e = threadIdx.x; e < ne; e += blockDim.x) {
        const int k = e / (2 * slide), rem = e % (2 * slide);
        const int pe = slide * (k - 1) + (rem % slide) + (rem < slide ? 0 : win);
        const float t = ref_angle_grid(pe, 0, skip + tab * tab_stride, 1);
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
        const int k = e / (2 * slide), rem = e % (2 * slide);
        const bool leaving = rem < slide;
        const int pe = slide * (k - 1) + (rem % slide) + (leaving ? 0 : win);
        const float t = ref_angle_grid(pe, 0, skip + tab * tab_stride, 1);
        int code = (1 << k) | (leaving ? (int)0x80000000 : 0);
        if (tab == 0 || tab * gs + k >= nslices) code = 0;       /* no previous slice / slice beyond the plan's last */
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
}

/* same number of angular bins as the full tables `full`: the per-cell and per-footprint bin windows serve both */
int build_delta_tables(SpokeTables &d, const SpokeTables &full, int ntab, int tab_stride, int skip, int win,
                       int slide, int gs, int nslices, cudaStream_t s)
{
    const int ne = 2 * gs * slide;
    d.ntab = ntab; d.nbins = full.nbins; d.npe = ne; d.gs = gs;
    float *scratch = nullptr;
    TRON_CUDA(cudaMalloc(&d.gx, (size_t)ntab * 2 * ne * sizeof(float4)));
    TRON_CUDA(cudaMalloc(&d.lut, (size_t)ntab * (d.nbins + 1) * sizeof(int)));
    TRON_CUDA(cudaMalloc(&scratch, (size_t)ntab * 2 * ne * sizeof(float)));
    delta_table_kernel<<<ntab, 256, 0, s>>>(d.gx, d.lut, scratch, scratch + (size_t)ntab * ne, ne, tab_stride, skip,
                                            d.nbins, win, slide, gs, nslices);
    TRON_CUDA(cudaGetLastError());
    TRON_CUDA(cudaStreamSynchronize(s));
    TRON_CUDA(cudaFree(scratch));
    return 0;
}

/* Tile schedule: tiles whose nearest cell lies within `near_r` of DC come first (nearest first) and take one
 * slice group per block, the rest follow row by row and take `gper` groups per block.  order[i] = ty << 16 | tx. */
int build_tile_schedule(int **d_order, int *n_near, int n, int th, float near_r)
{
    const int ntx = (n + 15) / 16, nty = (n + th - 1) / th, nt = ntx * nty;
    std::vector<std::pair<float, int>> nearv;
    std::vector<int> far;
    for (int t = 0; t < nt; ++t) {
        const int x0 = (t % ntx) * 16 - n / 2, y0 = (t / ntx) * th - n / 2;
        const float dx = x0 > 0 ? (float)x0 : (x0 + 15 < 0 ? (float)-(x0 + 15) : 0.f);
        const float dy = y0 > 0 ? (float)y0 : (y0 + th - 1 < 0 ? (float)-(y0 + th - 1) : 0.f);
        const float d2 = dx * dx + dy * dy;
        const int packed = ((t / ntx) << 16) | (t % ntx);
        if (d2 < near_r * near_r) nearv.push_back(std::make_pair(d2, packed)); else far.push_back(packed);
    }
    std::sort(nearv.begin(), nearv.end());
    std::vector<int> order;
    for (size_t i = 0; i < nearv.size(); ++i) order.push_back(nearv[i].second);
    order.insert(order.end(), far.begin(), far.end());
    *n_near = (int)nearv.size();
    TRON_CUDA(cudaMalloc(d_order, order.size() * sizeof(int)));
    TRON_CUDA(cudaMemcpy(*d_order, order.data(), order.size() * sizeof(int), cudaMemcpyHostToDevice));
    return 0;
}

#define TILE_ROUND_SLOTS 32
struct alignas(128) WarpShared {       /* per warp; its two data buffers (cap bytes each) follow the descriptors of all warps */
    unsigned long long bar[2];
    int4 info[2];                      /* group (launch-local), first slot, end slot, flags (1 first, 2 last of group, 4 final) */
    float4 seg_a[2][TILE_ROUND_SLOTS]; /* cos, sin, 1/cos, 1/sin */
    float4 seg_b[2][TILE_ROUND_SLOTS]; /* W/|cos| + margin, W/|sin| + margin, bits: shared address of sample r = 0, bits: slice mask */
};

/* the warp's staging cursor over (group, slot) */
struct StageCursor { int grp, slot, k0, cnt; };

template <int CH, int GS, bool HALF, int BT, int MB>
__global__ void __launch_bounds__(BT, MB)
grid_tile_kernel(const GridLaunch g, const int2 *__restrict__ foot_win, const int *__restrict__ order,
                 const int n_near, const int gper, const int cap)
{
    constexpr int TH = BT / 16;
    constexpr int WARPS = BT / 32;
    constexpr unsigned SAMP = CH * (HALF ? 4u : 8u);
    extern __shared__ __align__(128) unsigned char smem_raw[];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = g.n;
    WarpShared &sh = reinterpret_cast<WarpShared *>(smem_raw)[warp];
    const unsigned data0 = smem_u32(smem_raw) + (unsigned)(WARPS * sizeof(WarpShared)) + (unsigned)warp * 2u * (unsigned)cap;

    /* block -> (heavy block | tile, group range) */
    const int heavy_blocks = (g.nheavy + WARPS - 1) / WARPS;
    /* chains of `gper` groups are cut at multiples of gper of the plan's (shard-local) group index, so a slice
     * gets the same arithmetic whatever launch it travels in, as long as launches start on chain boundaries */
    const int ug0 = g.z0 / GS;
    const int chunk0 = ug0 / gper, nchunk = (ug0 + g.ngroups - 1) / gper - chunk0 + 1;
    int b = blockIdx.x;
    if (b < heavy_blocks * g.ngroups) {                    /* cells next to DC: one warp per cell (grid_common.cuh) */
        grid_heavy_path<CH, GS, HALF, BT, true>(g, b / g.ngroups, b % g.ngroups, 0);
        return;
    }
    b -= heavy_blocks * g.ngroups;
    int trank, grp_lo, grp_hi;
    if (b < n_near * g.ngroups) { trank = b / g.ngroups; grp_lo = b % g.ngroups; grp_hi = grp_lo + 1; }
    else {
        b -= n_near * g.ngroups;
        trank = n_near + b / nchunk;
        const int c = chunk0 + b % nchunk;
        grp_lo = max(c * gper - ug0, 0); grp_hi = min(g.ngroups, (c + 1) * gper - ug0);
    }
    const int tile = __ldg(order + trank);
    /* this warp's footprint and this lane's cell */
    const int x0 = (tile & 0xffff) * 16 + (warp & 1) * FOOT_W, y0 = (tile >> 16) * TH + (warp >> 1) * FOOT_H;
    if (x0 >= n || y0 >= n) return;
    const int XL = x0 - n / 2, YL = y0 - n / 2;
    const int XH = min(x0 + FOOT_W - 1, n - 1) - n / 2, YH = min(y0 + FOOT_H - 1, n - 1) - n / 2;
    {   /* the whole footprint beyond the last annulus: nothing is stored there */
        const int dx = XL > 0 ? XL : (XH < 0 ? -XH : 0), dy = YL > 0 ? YL : (YH < 0 ? -YH : 0);
        if (dx * dx + dy * dy > g.zero_r2) return;
    }
    const int nfx = (n + FOOT_W - 1) / FOOT_W;
    const int2 tw = __ldg(foot_win + (size_t)(y0 / FOOT_H) * nfx + x0 / FOOT_W);

    const int x = x0 + (lane & 7), y = y0 + (lane >> 3);
    const int X = x - n / 2, Y = y - n / 2;
    const bool stores = x < n && y < n && X * X + Y * Y <= g.zero_r2 && X * X + Y * Y > g.heavy_r2;
    int2 cw = make_int2(0, -1);                            /* packed bin window of the cell (empty) */
    int Rlo = 1, Rhi = 0;
    if (stores) {
        const int2 t = __ldg(g.cells + (size_t)y * n + x);
        Rlo = t.x & 0xffff; Rhi = t.x >> 16;
        if (Rlo <= Rhi) {
            const int lo16 = t.y & 0xffff;
            if (lo16 == CELL_ALL_SPOKES) cw = make_int2(CELL_ALL_SPOKES, 0);
            else {
                int b0 = (int)(short)lo16, b1 = t.y >> 16;
                if (b0 < 0) { b0 += g.nbins; b1 += g.nbins; }
                cw = make_int2(b0, b1);
            }
        }
    }

    const unsigned bar0 = smem_u32(&sh.bar[0]);
    if (lane == 0) {
        mbar_init(bar0, 1); mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    const size_t esz = HALF ? sizeof(__half2) : sizeof(float2);
    const int half_nro = g.nro >> 1;
    const float W = g.kb.W;
    const float Rmax = (float)(n / 2 - 1);

    /* ---- staging: next round of the cursor into buffer `buf` ---- */
    StageCursor cur; cur.grp = grp_lo; cur.slot = 0; cur.k0 = 0; cur.cnt = -1;
    auto stage = [&](int buf) {
        if (cur.grp >= grp_hi) return;
        const int ug = g.z0 / GS + cur.grp;
        const int tabi = g.tab_per_slice ? ug : 0;
        /* every group after the first of this block's chain is built from its predecessor (difference table) */
        const bool delta = g.tab_gx_d != nullptr && cur.grp != grp_lo;
        const int npe_t = delta ? g.npe_d : g.npe;
        if (cur.cnt < 0) {
            const TileWindow w = window_of(tw, (delta ? g.lut_d : g.lut) + (size_t)tabi * (g.nbins + 1), g.nbins, npe_t);
            cur.k0 = w.k0; cur.cnt = w.cnt;
        }
        const float4 *tab = (delta ? g.tab_gx_d : g.tab_gx) + (size_t)tabi * 2 * npe_t + cur.k0;
        const char *samples = (const char *)g.samples + (size_t)ug * GS * g.slide * g.nro * g.nc_total * esz;
        const int slot = cur.slot + lane;
        const bool valid = slot < cur.cnt;
        const float4 e = valid ? __ldg(tab + slot) : make_float4(1.f, 1.f, 0.f, 0.f);
        float ict, ist, hwx, hwy;
        axis_terms(e.x, W, ict, hwx);
        axis_terms(e.y, W, ist, hwy);
        /* readout indices r with (r ct, r st) inside the footprint's box grown by W (+ margin): every live tap of
         * every cell of the footprint satisfies |r ct - X| < W and |r st - Y| < W for some X, Y of it */
        const float cx = 0.5f * (float)(XL + XH) * ict, cy = 0.5f * (float)(YL + YH) * ist;
        const float ex = fmaf(0.5f * (float)(XH - XL), fabsf(ict), hwx + 0.05f);
        const float ey = fmaf(0.5f * (float)(YH - YL), fabsf(ist), hwy + 0.05f);
        const float lo = fmaxf(fmaxf(cx - ex, cy - ey), -Rmax);
        const float hi = fminf(fminf(cx + ex, cy + ey), Rmax);
        int ra = 0, rb = -1;
        const int mask = GS > 1 ? __float_as_int(e.w) : 1;
        if (valid && mask != 0 && lo <= hi) { ra = (int)ceilf(lo); rb = (int)floorf(hi); }
        /* byte range of the run, widened to 16-byte boundaries (fp16 storage: 24-byte samples) */
        const char *src = samples + ((ptrdiff_t)__float_as_int(e.z) * g.nro + half_nro + ra) * (ptrdiff_t)(g.nc_total * esz)
                          + (size_t)g.ch0 * esz;
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
            sh.seg_a[buf][lane] = make_float4(e.x, e.y, ict, ist);
            /* nothing staged (spoke outside every window of a partial group, or missing the box): negative
             * half-widths leave the cells no candidates */
            sh.seg_b[buf][lane] = make_float4(bytes ? hwx : -1.f, bytes ? hwy : -1.f,
                                              __int_as_float((int)(dst + lead) - ra * (int)SAMP), __int_as_float(mask));
        }
        const unsigned total = __shfl_sync(0xffffffffu, incl, count > 0 ? count - 1 : 0);
        const bool last = cur.slot + count >= cur.cnt;
        if (lane == 0)
            sh.info[buf] = make_int4(cur.grp, cur.slot, cur.slot + count,
                                     (cur.slot == 0 ? 1 : 0) | (last ? 2 : 0) | (last && cur.grp + 1 >= grp_hi ? 4 : 0)
                                     | (delta ? 8 : 0));
        __syncwarp();
        if (lane == 0) mbar_arrive_expect_tx(bar, count > 0 ? total : 0u);
        cur.slot += count;
        if (last) { cur.grp += 1; cur.slot = 0; cur.cnt = -1; }
    };

    stage(0);

    /* the window's coefficients as 64-bit operands of the packed evaluation */
    unsigned long long c2[TRONB_KB_DEG + 1];
#pragma unroll
    for (int m = 0; m <= TRONB_KB_DEG; ++m) c2[m] = *reinterpret_cast<const unsigned long long *>(&g.kb.c2[m]);
    float invW = g.kb.invW, sdc_as = g.sdc_as, sdc_bs = g.sdc_bs, Wk = g.kb.W;
    /* opaque to ptxas: otherwise it re-reads them from the constant bank inside the tap loop */
    asm volatile("" : "+f"(invW), "+f"(sdc_as), "+f"(sdc_bs), "+f"(Wk));

    float2 acc[GS][CH];
    int rel = 0, cnt = 0, npe_t = g.npe;                   /* this cell's window in the footprint's slot coordinates; table length */
    const float Xf = (float)X, Yf = (float)Y, Rhif = (float)Rhi;

    for (int it = 0;; ++it) {
        const int buf = it & 1;
        stage(buf ^ 1);                                     /* (every lane finished reading that buffer: __syncwarp below) */
        mbar_wait(bar0 + 8 * buf, (it >> 1) & 1);
        const int4 info = sh.info[buf];
        const bool delta = (info.w & 8) != 0;
        if (info.w & 1) {                                   /* first round of a group */
            /* difference group: slice 0 starts from the last slice of the previous group, the others collect
             * their own differences and are summed up at the end */
#pragma unroll
            for (int i = 0; i < CH; ++i) acc[0][i] = delta ? acc[GS - 1][i] : make_float2(0.f, 0.f);
#pragma unroll
            for (int s = 1; s < GS; ++s)
#pragma unroll
                for (int i = 0; i < CH; ++i) acc[s][i] = make_float2(0.f, 0.f);
            const int tabi = g.tab_per_slice ? g.z0 / GS + info.x : 0;
            const int *lut = (delta ? g.lut_d : g.lut) + (size_t)tabi * (g.nbins + 1);
            npe_t = delta ? g.npe_d : g.npe;
            const TileWindow wt = window_of(tw, lut, g.nbins, npe_t);
            const TileWindow wc = window_of(cw, lut, g.nbins, npe_t);
            rel = wc.k0 - wt.k0; if (rel < 0) rel += npe_t;
            cnt = wc.cnt;
        }
        /* descriptor of slot j sits at segA/segB + 16 j */
        unsigned segA = smem_u32(&sh.seg_a[buf][0]) - 16u * (unsigned)info.y;
        unsigned segB = smem_u32(&sh.seg_b[buf][0]) - 16u * (unsigned)info.y;
        asm volatile("" : "+r"(segA), "+r"(segB));          /* (not to be recomputed per spoke) */
        /* slots of this round inside the window: [rel, rel + cnt) and, if it wraps, [0, rel + cnt - npe) */
#pragma unroll 1
        for (int piece = 0; piece < 2; ++piece) {
            int ja, jb;
            if (piece == 0) { ja = max(rel, info.y); jb = min(min(rel + cnt, npe_t), info.z); }
            else { ja = info.y; jb = min(rel + cnt - npe_t, info.z); }
#pragma unroll 1
            for (int j = ja; j < jb; ++j) {
                const float4 ec = lds_f4(segA + 16u * (unsigned)j);
                const float4 ed = lds_f4(segB + 16u * (unsigned)j);
                /* candidate radii: integer points of {|r ct - X| < W} n {|r st - Y| < W} with a margin */
                const float cx = Xf * ec.z, cy = Yf * ec.w;
                float lo = fmaxf(fmaxf(cx - ed.x, cy - ed.y), -Rhif);
                float hi = fminf(fminf(cx + ed.x, cy + ed.y), Rhif);
                if (!(lo <= hi)) continue;
                const int r0 = (int)ceilf(lo), r1 = (int)floorf(hi);
                const int off = __float_as_int(ed.z), mask = __float_as_int(ed.w);
#pragma unroll 1
                for (int r = r0; r <= r1; ++r) {
                    if (abs(r) < Rlo) continue;            /* annulus, tron.cu:501-502,512,521 */
                    const float rf = (float)r;
                    const float dx = fma_ftz(ec.x, rf, -Xf);               /* tron.cu:514,516 as compiled */
                    if (!(fabsf(dx) < Wk)) continue;
                    const float dy = fma_ftz(ec.y, rf, -Yf);
                    if (!(fabsf(dy) < Wk)) continue;
                    float2 v[CH];
                    lds_sample<CH, HALF>(v, (unsigned)(off + r * (int)SAMP));
                    float w = kb_poly_xy_c2(dx, dy, invW, c2);
                    const float sdc = fmaf(sdc_as, fabsf(rf), sdc_bs);       /* tron.cu:412, times the output scale */
                    w *= (r == 0) ? sdc + sdc : sdc;       /* both of the reference's loops visit r = 0 */
                    if (GS > 1) w = __int_as_float(__float_as_int(w) ^ (mask & (int)0x80000000));   /* leaving spoke */
#pragma unroll
                    for (int s = 0; s < GS; ++s) {
                        if (GS == 1 || (mask >> s) & 1) {
#pragma unroll
                            for (int i = 0; i < CH; ++i) ffma2(acc[s][i], w, v[i]);
                        }
                    }
                }
            }
            if (rel + cnt <= npe_t) break;
        }
        if (info.w & 2) {
            if (GS > 1 && delta) {
#pragma unroll
                for (int s = 1; s < GS; ++s)
#pragma unroll
                    for (int i = 0; i < CH; ++i) { acc[s][i].x += acc[s - 1][i].x; acc[s][i].y += acc[s - 1][i].y; }
            }
            if (stores) store_cell<CH, GS>(g, acc, info.x, 0, x, y);
        }
        __syncwarp();                                       /* every lane is done with buffer `buf` */
        if (info.w & 4) break;
    }
}

/* ---------------------------------------------------------------------- */
/* launch                                                                  */
/* ---------------------------------------------------------------------- */
bool grid_tile_applicable(const GridLaunch &g)
{
    const bool off = getenv("TRON_NO_TILE") != nullptr;     /* diagnostic switches are read per launch */
    if (off || !g.tile_win8) return false;
    if (!(g.kb.fast && g.nro == g.n)) return false;         /* the plain case: fitted window, ridx = r */
    if (g.nch != g.nc_total || g.ch0 != 0) return false;    /* whole samples are staged */
    if (!(g.gs == 4 || g.gs == 1)) return false;
    if (g.nch != 2 && g.nch != 4 && g.nch != 6 && g.nch != 8) return false;
    const size_t esz = g.half_in ? 4 : 8;
    if (((uintptr_t)g.samples) % 16 != 0 && !g.half_in) return false;
    (void)esz;
    return g.dbg == nullptr;
}

template <int CH, int GS, bool HALF, int MB>
static int launch_tile_mb(const GridLaunch &g, long long blocks, size_t smem, int gper, unsigned cap, cudaStream_t s)
{
    constexpr int BT = 128;
    auto kern = grid_tile_kernel<CH, GS, HALF, BT, MB>;
    if (smem > 48 * 1024) TRON_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)blocks, BT, smem, s>>>(g, g.tile_win8, g.tile_sched8, g.n_near8, gper, (int)cap);
    TRON_CUDA(cudaGetLastError());
    return 0;
}

template <int CH, int GS, bool HALF>
static int launch_tile_t(GridLaunch g, cudaStream_t s)
{
    constexpr int BT = 128;
    constexpr int TH = BT / 16;
    g.ngroups = (g.z0 + g.nslices - 1) / GS - g.z0 / GS + 1;
    const int env_gper = getenv("TRON_TILE_GPER") ? atoi(getenv("TRON_TILE_GPER")) : 0;
    const int env_cap = getenv("TRON_TILE_CAP") ? atoi(getenv("TRON_TILE_CAP")) : 0;
    const int env_mb = getenv("TRON_TILE_MB") ? atoi(getenv("TRON_TILE_MB")) : 0;
    int gper = g.chain > 0 ? g.chain : (env_gper > 0 ? env_gper : 8);      /* difference tables: the plan fixed the chain length */
    if (gper < 1) gper = 1;
    const unsigned samp = CH * (HALF ? 4u : 8u);
    /* longest run a spoke can have inside a footprint's box: its diagonal (+ margins) */
    const float bw = (float)(FOOT_W - 1) + 2.f * g.kb.W + 0.2f, bh = (float)(FOOT_H - 1) + 2.f * g.kb.W + 0.2f;
    const unsigned longest = ((unsigned)(sqrtf(bw * bw + bh * bh) + 3.f) * samp + 31u) & ~15u;
    unsigned cap = env_cap > 0 ? (unsigned)env_cap : 4096u;
    if (cap < longest) cap = longest;
    cap = (cap + 127u) & ~127u;
    const size_t smem = (BT / 32) * (sizeof(WarpShared) + 2 * (size_t)cap);
    if (smem > 200 * 1024) return -1;
    const int ntiles = ((g.n + 15) / 16) * ((g.n + TH - 1) / TH);
    const int heavy_blocks = (g.nheavy + BT / 32 - 1) / (BT / 32);
    const int ug0 = g.z0 / GS;
    const int nchunk = (ug0 + g.ngroups - 1) / gper - ug0 / gper + 1;
    const long long blocks = (long long)(heavy_blocks + g.n_near8) * g.ngroups + (long long)(ntiles - g.n_near8) * nchunk;
    if (blocks > 0x7fffffffLL) return -1;
    constexpr int MB = CH * GS >= 32 ? 4 : (CH * GS >= 16 ? 5 : 6);
    if (CH == 6 && GS == 4 && !HALF) {                       /* the benchmark instantiation: blocks per SM switchable */
        if (env_mb == 4) return launch_tile_mb<CH, GS, HALF, (CH == 6 && GS == 4 && !HALF) ? 4 : MB>(g, blocks, smem, gper, cap, s);
        if (env_mb == 6) return launch_tile_mb<CH, GS, HALF, (CH == 6 && GS == 4 && !HALF) ? 6 : MB>(g, blocks, smem, gper, cap, s);
    }
    return launch_tile_mb<CH, GS, HALF, MB>(g, blocks, smem, gper, cap, s);
}

template <int CH>
static int launch_tile_c(const GridLaunch &g, cudaStream_t s)
{
    if (g.gs == 4) return g.half_in ? launch_tile_t<CH, 4, true>(g, s) : launch_tile_t<CH, 4, false>(g, s);
    return g.half_in ? launch_tile_t<CH, 1, true>(g, s) : launch_tile_t<CH, 1, false>(g, s);
}

int launch_grid_tile(const GridLaunch &g, cudaStream_t s)
{
    switch (g.nch) {
    case 2: return launch_tile_c<2>(g, s);
    case 4: return launch_tile_c<4>(g, s);
    case 6: return launch_tile_c<6>(g, s);
    case 8: return launch_tile_c<8>(g, s);
    }
    return -1;
}

} // namespace tronb
