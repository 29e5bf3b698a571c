/*
 * grid.cu -- adjoint interpolation ("gridding"): radial samples -> Cartesian grid.
 *
 * Replaces, in one kernel, the reference's precompensate + gridradial2d
 * (/root/reference/src/tron.cu:405-416 and 465-536).
 *
 * The reference is a gather with one thread per cell that tests EVERY spoke and
 * every radius of an annulus (99.5 % of its tests fail).  This kernel keeps the
 * gather (registers accumulate, no atomics, no sample presort) but derives the
 * candidates analytically:
 *
 *   - the spokes of a table are ordered by angle mod pi (plan-time table with an
 *     angular-bin LUT, built by spoke_table_kernel below);
 *   - a cell at radius R only visits the spokes whose line passes within
 *     W*sqrt(2) of it: |angle - atan2(Y,X)| <= asin(W sqrt2 / R)  (mod pi); the
 *     bin range of that window is slice independent and comes from a per-cell
 *     table (cell_table_kernel);
 *   - on each such spoke the candidate radii are the integer points of
 *     {r : |r ct - X| < W} n {r : |r st - Y| < W}, computed from 1/ct, 1/st
 *     with a conservative margin;
 *   - every candidate is then decided by the reference's own predicate,
 *     evaluated with the reference's own operations (refmath.cuh): the set of
 *     taps is identical by construction, only the visiting order differs.
 *
 * Sliding windows (-d): with golden angles the spoke angle depends on the
 * ABSOLUTE spoke index (tron.cu:509,630), so the tap (cell, spoke, r) and its
 * weight are the same in every slice whose window holds that spoke.  A thread
 * therefore walks the union of the windows of GS consecutive slices once and
 * feeds GS accumulator sets; the per-tap test/weight/load work is shared, only
 * the FMAs (packed FFMA2) are per slice.
 *
 * Reference quirks reproduced: support = square n annulus (Rlo..Rhi, SURVEY
 * F4); r = 0 counted twice when Rlo == 0 (F5); ridx = (r*nro)/nxos with C
 * truncation; golden angle from the absolute spoke index in f32 (F16).
 * Folded in: ramp density compensation a|ro - nro/2| + b (tron.cu:408-412) and
 * the 1/(nxos*npe) scale (tron.cu:532).
 *
 * Cells near DC see every spoke (hundreds of taps): one warp per cell, lanes
 * striding over spokes, partial sums combined by shuffles; those blocks are
 * numbered first, then the tiles nearest DC, so the longest work starts first.
 *
 * Output is planar: grid[slice][ch][row][col].
 */
#include "grid_common.cuh"
#include <algorithm>
#include <math.h>
#include <stdlib.h>
#include <utility>
#include <vector>

namespace tronb {

/* ---------------------------------------------------------------------- */
/* plan-time tables                                                        */
/* ---------------------------------------------------------------------- */

/* One table per slice group (golden angle) or one shared table (linear).
 * Entry k of a table: (cos, sin, 1/cos, 1/sin) of the k-th spoke by angle mod pi,
 * pe_sorted[k] = index of that spoke relative to the group's first spoke, with
 * bits 24.. = mask of the group's slices whose window [s*slide, s*slide+win) holds it. */
__global__ void spoke_table_kernel(float4 *cs, int *pe_sorted, float4 *gx, int *lut, float2 *cs_lin,
                                   float *key_unsorted, float *key_sorted,
                                   int npe, int npe_formula, int tab_stride, int skip, int golden, int adjoint,
                                   int nbins, int win, int slide, int gs, int nslices)
{
    const int tab = blockIdx.x;
    float *ku = key_unsorted + (size_t)tab * npe;
    float *ks = key_sorted + (size_t)tab * npe;
    float2 *lin = cs_lin + (size_t)tab * npe;
    for (int pe = threadIdx.x; pe < npe; pe += blockDim.x) {
        float t = adjoint ? ref_angle_grid(pe, npe_formula, skip + tab * tab_stride, golden)
                          : ref_angle_degrid(pe, npe_formula, skip, golden);
        lin[pe] = make_float2(cos_approx(t), sin_approx(t));
        float key = fmodf(t, PI_F);
        if (key < 0.f) key += PI_F;
        if (key >= PI_F) key -= PI_F;
        ku[pe] = key;
    }
    __syncthreads();
    if (!adjoint) return;
    float4 *csd = cs + (size_t)tab * npe;
    int *ped = pe_sorted + (size_t)tab * npe;
    float4 *gxd = gx + (size_t)tab * 2 * npe;             /* stored twice: a circular window never wraps */
    for (int pe = threadIdx.x; pe < npe; pe += blockDim.x) {
        float key = ku[pe];
        int rank = 0;
        for (int j = 0; j < npe; ++j) {
            float kj = ku[j];
            rank += (kj < key) || (kj == key && j < pe);
        }
        float2 c = lin[pe];
        float ic = fabsf(c.x) > 1e-18f ? 1.0f / c.x : copysignf(1e18f, c.x);
        float is = fabsf(c.y) > 1e-18f ? 1.0f / c.y : copysignf(1e18f, c.y);
        int mask = 0;
        for (int s = 0; s < gs; ++s)
            if (tab * gs + s < nslices && pe >= s * slide && pe < s * slide + win) mask |= 1 << s;
        csd[rank] = make_float4(c.x, c.y, ic, is);
        ped[rank] = pe | (mask << 24);
        gxd[rank] = gxd[rank + npe] = make_float4(c.x, c.y, __int_as_float(pe), __int_as_float(mask));
        ks[rank] = key;
    }
    __syncthreads();
    const float lut_scale = (float)nbins / PI_F;
    int *l = lut + (size_t)tab * (nbins + 1);
    for (int b = threadIdx.x; b <= nbins; b += blockDim.x) {
        int lo = 0, hi = npe;                     /* first k with bin(ks[k]) >= b */
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (angle_bin(ks[mid], lut_scale, nbins) >= b) hi = mid; else lo = mid + 1;
        }
        l[b] = lo;
    }
}

/* Per-cell, slice-independent geometry (tron.cu:498-502 and the angular window):
 *   .x = Rlo | Rhi << 16          (Rlo > Rhi: no taps)
 *   .y = (b0 & 0xffff) | b1 << 16 first/last angular bin (b0 may be negative, b1 may exceed nbins:
 *        the window wraps), or CELL_ALL_SPOKES in the low half when every spoke must be visited. */
__global__ void cell_table_kernel(int2 *cells, int n, float W, int nbins)
{
    for (int id = blockIdx.x * blockDim.x + threadIdx.x; id < n * n; id += gridDim.x * blockDim.x) {
        const int X = id % n - n / 2, Y = id / n - n / 2;
        float R = ref_hypotf((float)X, (float)Y);
        int Rhi = (int)fminf(floorf(R + W), (float)(n / 2 - 1));
        int Rlo = (int)fmaxf(ceilf(R - W), 0.f);
        if (Rlo > Rhi) { Rlo = 1; Rhi = 0; }
        /* a spoke reaches the cell only if its line passes within W*sqrt(2): |sin(angle diff)| < reach/R */
        const float reach = W * 1.41421368f + 2e-3f;
        float xr = reach / fmaxf(R, 1e-6f);
        int lo16 = CELL_ALL_SPOKES, b1 = 0;
        if (xr <= 0.7f) {
            float T = atan2f((float)Y, (float)X);
            if (T < 0.f) T += PI_F;
            if (T >= PI_F) T -= PI_F;
            float delta = xr * fmaf(0.25f * xr, xr, 1.0f) + 2e-4f;     /* >= asin(xr) for xr <= 0.7 */
            const float lut_scale = (float)nbins / PI_F;
            int b0 = (int)floorf((T - delta) * lut_scale);
            b1 = (int)floorf((T + delta) * lut_scale);
            if (b1 - b0 + 1 < nbins) lo16 = b0 & 0xffff; else b1 = 0;
        }
        cells[id] = make_int2(Rlo | (Rhi << 16), lo16 | (b1 << 16));
    }
}

/* tiles of 16 x th cells (one thread block each) ordered by the distance of their nearest cell from DC
 * (short launches: the longest tiles start first), or row by row (long launches: the expensive tiles next to DC
 * are spread over the launch instead of running all at once -- 5.83 -> 5.64 us/slice on cfg2) */
int build_tile_order(int **d_order, int n, int th, bool raster)
{
    int tx = (n + 15) / 16, nt = tx * ((n + th - 1) / th);
    std::vector<std::pair<float, int>> key(nt);
    for (int t = 0; t < nt; ++t) {
        int x0 = (t % tx) * 16 - n / 2, y0 = (t / tx) * th - n / 2;
        float dx = x0 > 0 ? (float)x0 : (x0 + 15 < 0 ? (float)-(x0 + 15) : 0.f);
        float dy = y0 > 0 ? (float)y0 : (y0 + th - 1 < 0 ? (float)-(y0 + th - 1) : 0.f);
        const float k = raster ? (float)t : dx * dx + dy * dy;
        key[t] = std::make_pair(k, t);
    }
    std::sort(key.begin(), key.end());
    std::vector<int> order(nt);
    for (int t = 0; t < nt; ++t) order[t] = ((key[t].second / tx) << 16) | (key[t].second % tx);   /* ty << 16 | tx */
    TRON_CUDA(cudaMalloc(d_order, nt * sizeof(int)));
    TRON_CUDA(cudaMemcpy(*d_order, order.data(), nt * sizeof(int), cudaMemcpyHostToDevice));
    return 0;
}

/* Cells with X^2 + Y^2 <= r2 are "heavy": their spoke window holds about
 * HEAVY_SPOKES or more spokes.  The warp-per-cell path exists to keep the cells next to DC (which see every
 * spoke) off the critical path of a launch; it costs a 48-register shuffle reduction per cell, so the
 * threshold follows the size of the launch: 24 spokes when a launch is small (one 512^2 slice: 31 vs 76 us
 * with 96), 96 when there is enough other work to hide the long cells (cfg2 at 64+ slices per launch:
 * 5.8 vs 6.2 us/slice with 24; cfg5 8-coil shard, 2048^2: 3.3 vs 4.4 ms).  Returns the list (packed y<<16 | x, nearest DC first). */
int build_heavy_cells(int **d_cells, int *nheavy, int *heavy_r2, int n, int npe, float W, int heavy_spokes)
{
    const int HEAVY_SPOKES = getenv("TRON_HEAVY") ? atoi(getenv("TRON_HEAVY")) : heavy_spokes;
    *d_cells = nullptr; *nheavy = 0; *heavy_r2 = -1;
    double arg = HEAVY_SPOKES * 3.14159265358979323846 / (2.0 * npe);
    if (arg >= 1.2) return 0;                              /* too few spokes for any cell to be heavy */
    double reach = (double)W * 1.41421368 + 2e-3;
    double rh = reach / sin(arg);
    if (rh > n / 2) rh = n / 2;
    int r2 = (int)floor(rh * rh);
    std::vector<std::pair<int, int>> cells;
    int ri = (int)ceil(rh) + 1;
    for (int Y = -ri; Y <= ri; ++Y)
        for (int X = -ri; X <= ri; ++X) {
            int x = X + n / 2, y = Y + n / 2;
            if (x < 0 || y < 0 || x >= n || y >= n) continue;
            if (X * X + Y * Y <= r2) cells.push_back(std::make_pair(X * X + Y * Y, (y << 16) | x));
        }
    if (cells.empty()) return 0;
    std::sort(cells.begin(), cells.end());
    std::vector<int> packed(cells.size());
    for (size_t i = 0; i < cells.size(); ++i) packed[i] = cells[i].second;
    TRON_CUDA(cudaMalloc(d_cells, packed.size() * sizeof(int)));
    TRON_CUDA(cudaMemcpy(*d_cells, packed.data(), packed.size() * sizeof(int), cudaMemcpyHostToDevice));
    *nheavy = (int)packed.size(); *heavy_r2 = r2;
    return 0;
}

static int pick_nbins(int npe)
{
    int nb = 64;
    while (nb < 2 * npe && nb < 8192) nb <<= 1;
    return nb;
}

/* npe: entries per table (union window of a slice group); npe_formula: the npe of the linear-angle
 * formula (tron.cu:509,555); tab_stride: spokes between the first spokes of consecutive tables */
int launch_build_tables(SpokeTables &t, int npe, int npe_formula, int ntab, int tab_stride, int skip, int golden,
                        int adjoint, int win, int slide, int gs, int nslices, int n, float W, cudaStream_t s)
{
    t.ntab = ntab; t.nbins = pick_nbins(npe); t.npe = npe; t.gs = gs;
    size_t ne = (size_t)ntab * npe;
    float *scratch = nullptr;
    TRON_CUDA(cudaMalloc(&t.cs_lin, ne * sizeof(float2)));
    if (adjoint) {
        TRON_CUDA(cudaMalloc(&t.cs, ne * sizeof(float4)));
        TRON_CUDA(cudaMalloc(&t.pe, ne * sizeof(int)));
        TRON_CUDA(cudaMalloc(&t.gx, 2 * ne * sizeof(float4)));
        TRON_CUDA(cudaMalloc(&t.lut, (size_t)ntab * (t.nbins + 1) * sizeof(int)));
        TRON_CUDA(cudaMalloc(&t.cells, (size_t)n * n * sizeof(int2)));
        int blocks = (n * n + 255) / 256; if (blocks > 4096) blocks = 4096;
        cell_table_kernel<<<blocks, 256, 0, s>>>(t.cells, n, W, t.nbins);
        TRON_CUDA(cudaGetLastError());
    }
    TRON_CUDA(cudaMalloc(&scratch, 2 * ne * sizeof(float)));
    int threads = npe >= 1024 ? 1024 : 256;
    spoke_table_kernel<<<ntab, threads, 0, s>>>(t.cs, t.pe, t.gx, t.lut, t.cs_lin, scratch, scratch + ne,
                                                npe, npe_formula, tab_stride, skip, golden, adjoint, t.nbins,
                                                win, slide, gs, nslices);
    TRON_CUDA(cudaGetLastError());
    TRON_CUDA(cudaStreamSynchronize(s));
    TRON_CUDA(cudaFree(scratch));
    return 0;
}

/* one launch: the heavy-cell blocks come first (longest critical path), then the tiles.
 * Grid: x = slice group, y = heavy block / tile rank, z = channel chunk (launch_grid_cg).
 * BT threads per block: 256 (16x16 tile), or 128 (16x8 tile) where the accumulators need the
 * registers: 5 blocks of 128 threads leave 96 registers per thread, 3 of 256 only 80. */
template <int CH, int GS, bool HALF, int BT, bool PLAIN, int MB>
__global__ void __launch_bounds__(BT, MB)
grid_gather_kernel(const GridLaunch g, const int rank0)
{
    constexpr int WARPS = BT / 32;
    const int heavy_blocks = (g.nheavy + WARPS - 1) / WARPS;
    const int rank = rank0 + (int)blockIdx.y, grp = blockIdx.x, chunk = blockIdx.z;
    const long long t0 = g.dbg ? clock64() : 0;
    if (rank < heavy_blocks) grid_heavy_path<CH, GS, HALF, BT, PLAIN>(g, rank, grp, chunk);
    else grid_tile_path<CH, GS, HALF, BT, PLAIN>(g, rank - heavy_blocks, grp, chunk);
    if (g.dbg) {                                          /* per-warp cycle counts (TRON_GRID_DEBUG) */
        __syncwarp();
        const long long t1 = clock64();
        const size_t b = ((size_t)chunk * (gridDim.y + rank0) + rank) * gridDim.x + grp;
        if ((threadIdx.x & 31) == 0 && b < (size_t)8 * 65536) g.dbg[b * 8 + (threadIdx.x >> 5)] = t1 - t0;
    }
}

template <int CH, int GS, bool HALF, bool PLAIN>
static int launch_grid_cghp(const GridLaunch &g, cudaStream_t s)
{
    constexpr int BT = CH * GS >= 16 ? 128 : 256;
    const int tiles = ((g.n + 15) / 16) * ((g.n + BT / 16 - 1) / (BT / 16));
    const int ranks = tiles + (g.nheavy + BT / 32 - 1) / (BT / 32);
    for (int r0 = 0; r0 < ranks; r0 += 65535) {            /* gridDim.y limit */
        dim3 grid(g.ngroups, std::min(65535, ranks - r0), g.nch / CH);
        /* 128-thread blocks: 6 per SM (80 registers, 57 words of spill on the cold paths) measured 2 % faster
         * than 5 per SM (96 registers, no spill) for the 6-channel 4-slice instantiation; the others keep 5 */
        constexpr int MB = BT == 128 ? (CH * GS >= 32 ? 4 : ((CH == 6 && GS == 4 && !HALF && PLAIN) ? 6 : 5)) : 4;
        grid_gather_kernel<CH, GS, HALF, BT, PLAIN, MB><<<grid, BT, 0, s>>>(g, r0);
        TRON_CUDA(cudaGetLastError());
    }
    return 0;
}

template <int CH, int GS>
static int launch_grid_cg(GridLaunch g, cudaStream_t s)
{
    g.ngroups = (g.z0 + g.nslices - 1) / GS - g.z0 / GS + 1;
    /* (which of the two heavy-cell lists is in g.heavy_* was decided when the plan was made: the choice changes
     * the summation order of the cells in between, and a slice must not depend on the launch it travels in) */
    if ((double)g.n * g.n * g.ngroups >= 4.0e6) {         /* (at 2 M the row order still loses to the tail: 7.1 vs 5.8 us) */
        if (g.tile_order_rows) g.tile_order = g.tile_order_rows;
        if (g.tile_order8_rows) g.tile_order8 = g.tile_order8_rows;
    }
    const bool plain = g.kb.fast && g.nro == g.n;
    if (g.half_in) return plain ? launch_grid_cghp<CH, GS, true, true>(g, s) : launch_grid_cghp<CH, GS, true, false>(g, s);
    return plain ? launch_grid_cghp<CH, GS, false, true>(g, s) : launch_grid_cghp<CH, GS, false, false>(g, s);
}

int launch_grid(const GridLaunch &g, cudaStream_t s)
{
    /* vector loads need an even first channel and an even channel count */
    size_t esz = g.half_in ? 4 : 8;
    bool aligned = (((uintptr_t)g.samples) % (2 * esz) == 0) && (g.nc_total % 2 == 0) && (g.ch0 % 2 == 0);
    if (g.nslices <= 0 || g.nch <= 0) return 0;
    const bool no_wide = getenv("TRON_NO_WIDE") != nullptr;      /* diagnostic switch, read per launch */
    if (!no_wide && grid_wide_applicable(g)) return launch_grid_wide(g, s);   /* nc >= 16: lanes = channels */
    if (aligned && grid_scatter_applicable(g)) {           /* tiles accumulated in shared memory, sample driven */
        const int rc = launch_grid_scatter(g, s);
        if (rc >= 0) return rc;
    }
    if (aligned && grid_tile_applicable(g)) {              /* samples staged in shared memory by bulk copies */
        const int rc = launch_grid_tile(g, s);
        if (rc >= 0) return rc;                            /* < 0: geometry outside that kernel's limits */
    }
    if (g.gs == 4) {                                      /* sliding windows share taps across 4 slices */
        if (!aligned || g.nch % 2) return launch_grid_cg<1, 4>(g, s);
        if (g.nch % 6 == 0) return launch_grid_cg<6, 4>(g, s);
        /* 8 or 16 channels (cfg4): 8 per thread (126 registers, 4 blocks/SM) halves the tap evaluations of the
         * 4-channel chunking -- measured 8.9 -> 7.5 us per 16-channel frame */
        if (g.nch % 8 == 0) return launch_grid_cg<8, 4>(g, s);
        if (g.nch % 4 == 0) return launch_grid_cg<4, 4>(g, s);
        return launch_grid_cg<2, 4>(g, s);
    }
    if (g.gs != 1) { set_error("unsupported slice group size %d", g.gs); return TRON_EINVAL; }
    if (!aligned || g.nch % 2) return launch_grid_cg<1, 1>(g, s);
    if (g.nch % 8 == 0) return launch_grid_cg<8, 1>(g, s);
    if (g.nch % 6 == 0) return launch_grid_cg<6, 1>(g, s);
    if (g.nch % 4 == 0) return launch_grid_cg<4, 1>(g, s);
    return launch_grid_cg<2, 1>(g, s);
}

/* ---------------------------------------------------------------------- */
/* layout helpers (tests / legacy surface)                                 */
/* ---------------------------------------------------------------------- */
__global__ void interleave_kernel(float2 *dst, const float2 *planar, int nch, int n)
{
    size_t plane = (size_t)n * n;
    size_t total = plane * nch;
    const float2 *src = planar + (size_t)blockIdx.y * total;
    float2 *d = dst + (size_t)blockIdx.y * total;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        size_t cell = i / nch; int ch = (int)(i - cell * nch);
        d[i] = src[(size_t)ch * plane + cell];
    }
}

__global__ void deinterleave_kernel(float2 *planar, const float2 *src, int nch, int n)
{
    size_t plane = (size_t)n * n;
    size_t total = plane * nch;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int ch = (int)(i / plane); size_t cell = i - (size_t)ch * plane;
        planar[i] = src[cell * nch + ch];
    }
}

int launch_interleave(float2 *dst, const float2 *planar, int nch, int n, int nslices, cudaStream_t s)
{
    dim3 grid(1184, nslices);
    interleave_kernel<<<grid, 256, 0, s>>>(dst, planar, nch, n);
    TRON_CUDA(cudaGetLastError());
    return 0;
}

int launch_deinterleave(float2 *planar, const float2 *src, int nch, int n, cudaStream_t s)
{
    deinterleave_kernel<<<1184, 256, 0, s>>>(planar, src, nch, n);
    TRON_CUDA(cudaGetLastError());
    return 0;
}

} // namespace tronb
