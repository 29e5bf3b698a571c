/* tron_internal.h -- plan object and kernel launch interfaces (not installed). */
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/tron.h"
#include "refmath.cuh"

namespace tronb {

void set_error(const char *fmt, ...);
int  cuda_fail(cudaError_t e, const char *what, const char *file, int line);
#define TRON_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) \
    return tronb::cuda_fail(e__, #call, __FILE__, __LINE__); } while (0)

/* Entry points run on the plan's device and leave the caller's current device as they found it
 * (hosts such as torch keep their own notion of the current device). */
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; if (dev != prev) cudaSetDevice(dev); }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard &) = delete;
    DeviceGuard &operator=(const DeviceGuard &) = delete;
};

/* One sorted-spoke table per slice group (golden angle) or one shared table (linear). */
struct SpokeTables {
    float4 *cs = nullptr;     /* [ntab][npe]  (cos, sin, 1/cos, 1/sin), sorted by angle mod pi */
    int    *pe = nullptr;     /* [ntab][npe]  group-relative spoke index | slice mask << 24 */
    float4 *gx = nullptr;     /* [ntab][2*npe] (cos, sin, bits: group-relative spoke index, bits: slice mask), same order,
                                 stored twice back to back so a circular window is a plain range */
    int    *lut = nullptr;    /* [ntab][nbins+1] first sorted entry of each angular bin */
    float2 *cs_lin = nullptr; /* [ntab][npe]  (cos, sin) in acquisition order (degridding) */
    int2   *cells = nullptr;  /* [n][n] slice-independent cell geometry (band, angular bins) */
    int ntab = 0, nbins = 0;
    int npe = 0;              /* entries per table = union window of a slice group */
    int gs = 1;               /* slices per group */
};

/* grid_scatter.cu: per-slice sorted spoke tables (full window / difference from the previous slice), 16 x 16 tile
 * windows and the tile schedule.  Device pointers + scalars: passed to the kernel by value. */
struct ScatterPlan {
    float4 *tab_full = nullptr; int *lut_full = nullptr;     /* [ntab][2*win], [ntab][nbins+1]; ntab = slices (golden) or 1 */
    float4 *tab_delta = nullptr; int *lut_delta = nullptr;   /* [slices][2*ne_delta], ...; null: no difference tables */
    int2 *tile_win = nullptr;                                /* packed angular-bin window per tile */
    int *sched = nullptr;                                    /* tiles (ty << 16 | tx): n_near, then n_far, then ntiles_empty */
    int win = 0, ne_delta = 0, per_slice = 0;
    int th = 16;                                             /* tile height the windows and schedules were built for */
    int chain = 1, chain_near = 1;                           /* slices per chain for far / near tiles */
    int n_near = 0, n_far = 0, ntiles_empty = 0;
    int ready = 0;
    /* the schedule above serves long launches; launches of fewer than `short_below` slices have too few
     * (tile, chain) tasks to fill the GPU and take this one: shorter chains, more tiles on the split path */
    int *sched_short = nullptr;
    int chain_short = 1, chain_near_short = 1, n_near_short = 0, n_far_short = 0, short_below = 0;
};

struct GridLaunch {               /* everything the gridding kernel needs */
    const void *samples;          /* first spoke of shard-local slice 0 */
    float2 *grid;                 /* [nslices][nch][n][n] */
    const float4 *tab_cs; const int *tab_pe; const float4 *tab_gx; const int *lut; const int2 *cells;
    const int *tile_order;        /* [tiles] 16x16 tiles, heaviest (nearest DC) first */
    const int *tile_order8;       /* same for 16x8 tiles (128-thread blocks) */
    const int *tile_order_rows, *tile_order8_rows;   /* both lists row by row (long launches) */
    const int *heavy_cells; int nheavy; int heavy_r2;   /* cells with X^2+Y^2 <= heavy_r2: one warp each */
    const int *heavy_cells_big; int nheavy_big; int heavy_r2_big;   /* the shorter list for launches with much other work */
    int tab_per_slice;            /* 1: table index = slice group, 0: shared */
    int nbins;
    int n, nro, nc_total, ch0, nch;
    int npe;                      /* entries per spoke table (union window of a group) */
    int gs, ngroups;              /* slices per group; groups touched by this launch */
    int z0, nslices, slide;
    KbParams kb;
    float sdc_a, sdc_b, scale;
    float sdc_as, sdc_bs;         /* sdc_a * scale, sdc_b * scale: the gather folds the output scale into the weight */
    int half_in;
    long long *dbg;               /* optional per-warp cycle counts [blocks][8] */
    const float4 *tab_gx_d = nullptr;  /* grid_tile.cu: sliding-window difference tables (same layout as tab_gx), or null */
    const int *lut_d = nullptr; int npe_d = 0;
    int chain = 0;                     /* groups per chain when difference tables are in use (fixed by the plan) */
    const int2 *tile_win8 = nullptr;   /* grid_tile.cu: per 8x4 warp footprint, packed angular-bin window (union of its cells) */
    const int *tile_sched8 = nullptr;  /* grid_tile.cu: tile schedule, the n_near8 tiles next to DC first */
    int n_near8 = 0;
    const ScatterPlan *scat = nullptr; /* grid_scatter.cu (host pointer; the launcher passes the struct by value) */
    int wide_prefetch = 0;        /* grid_wide.cu: 0 none, 1 L1, 2 L2 prefetch of a sample when its list entry is written */
    int zero_r2;                  /* cells with X^2 + Y^2 > zero_r2 can hold no sample and are NOT stored (the FFT pass
                                     that follows does not fetch them); INT_MAX: every cell is stored */
};

struct DegridLaunch {
    void *samples;                /* [npe][nro][nc_total] */
    const float2 *grid;           /* planar [nch][n][n] */
    const float2 *cs;             /* [npe] (cos, sin) acquisition order */
    int n, nro, npe, nc_total, ch0, nch;
    KbParams kb;
    int half_out;
    int nimg = 1;                 /* grids per launch: samples [nimg][npe][nro][nc_total], grid [nimg][nch][n][n] */
    int cs_stride = 0;            /* entries between the spoke tables of consecutive grids (0: shared) */
    int pair_spokes = 0;          /* degrid_wide.cu: spokes pe, pe + 1 are neighbours in angle (linear order): one warp takes both */
};

int launch_grid(const GridLaunch &g, cudaStream_t s);
bool grid_tile_applicable(const GridLaunch &g);
int launch_grid_tile(const GridLaunch &g, cudaStream_t s);
int build_tile_windows(int2 **d_win, const int2 *cells, int n, int nbins, int fw, int fh, cudaStream_t s);
int build_tile_schedule(int **d_order, int *n_near, int n, int th, float near_r);
int build_delta_tables(SpokeTables &d, const SpokeTables &full, int ntab, int tab_stride, int skip, int win,
                       int slide, int gs, int nslices, cudaStream_t s);
bool grid_scatter_applicable(const GridLaunch &g);
int launch_grid_scatter(const GridLaunch &g, cudaStream_t s);
int scatter_plan_build(ScatterPlan &sp, const int2 *cells, int nbins, int n, int nslices, int win, int slide, int skip,
                       int golden, float W, int nc, cudaStream_t s);
void scatter_plan_free(ScatterPlan &sp);
bool grid_wide_applicable(const GridLaunch &g);
int launch_grid_wide(const GridLaunch &g, cudaStream_t s);
int launch_degrid(const DegridLaunch &d, cudaStream_t s);
bool degrid_wide_applicable(const DegridLaunch &d);
int launch_degrid_wide(const DegridLaunch &d, float2 *scratch, cudaStream_t s);
int launch_build_tables(SpokeTables &t, int npe, int npe_formula, int ntab, int tab_stride, int skip, int golden,
                        int adjoint, int win, int slide, int gs, int nslices, int n, float W, cudaStream_t s);
int build_tile_order(int **d_order, int n, int th, bool raster);
int build_heavy_cells(int **d_cells, int *nheavy, int *heavy_r2, int n, int npe, float W, int heavy_spokes);
int launch_interleave(float2 *dst, const float2 *planar, int nch, int n, int nslices, cudaStream_t s);
int launch_deinterleave(float2 *planar, const float2 *src, int nch, int n, cudaStream_t s);

/* FFT stage */
struct FftPlan {
    int n = 0;                    /* transform length (nxos) */
    int nkeep = 0;                /* nx */
    int nfac = 0; int fac[16];    /* radix schedule */
    float2 *tw = nullptr;         /* [n] exp(+2 pi i k / n) */
    int lines = 0;                /* lines per CTA (generic path) */
    int pow2 = 0;                 /* register radix-8 fast path available */
    size_t smem = 0;
};
int fft_plan_init(FftPlan &f, int n, int nkeep);
void fft_plan_free(FftPlan &f);

struct AdjFftLaunch {
    const float2 *grid;           /* planar [nslices][nch][n][n] */
    float2 *tmp;                  /* [nslices][nch][nkeep][n] */
    void *out;                    /* see mode */
    const float *deapod;          /* [nkeep][nkeep] reciprocal weights */
    int nslices, nch, nc_total, ch0;
    int mode;                     /* 0 rss (complex64 out), 1 single-channel complex, 2 per-coil interleaved,
                                     3 partial sum of squares (float) */
    int half_out;
    int zero_r2 = 0x7fffffff;     /* grid cells with X^2 + Y^2 > zero_r2 are known to be zero and are not read */
    int *sync = nullptr;          /* 2 * nslices counters: enables the single-launch path (fft.cu: p2w_adj_fused) */
    int ring = 0;                 /* slices the intermediate `tmp` holds on that path (a ring of slots) */
};
int launch_adj_fft(const FftPlan &f, const AdjFftLaunch &a, cudaStream_t s);
bool adj_fft_single_launch(const FftPlan &f, const AdjFftLaunch &a);

struct FwdFftLaunch {
    const void *img;              /* [nx][nx][nc_total] channel-interleaved */
    float2 *tmp;                  /* [nch][n][nkeep] */
    float2 *grid;                 /* planar [nch][n][n] */
    const float *deapod;          /* [nkeep][nkeep] reciprocal weights of the padded region */
    int nch, nc_total, ch0;
    int half_in;
    int nimg = 1;                 /* images per launch: img [nimg][nx][nx][nc_total], tmp/grid [nimg][nch]... */
};
int launch_fwd_fft(const FftPlan &f, const FwdFftLaunch &a, cudaStream_t s);
int launch_deapod_tables(float *adj_tab, float *fwd_tab, int nx, int nxos, float W, float gridos, cudaStream_t s);

/* coil combination of channel-interleaved per-coil images (combine.cu) */
int launch_coil_combine(void *out, const float2 *coil, size_t npix, int nc, int mode, int half_out, cudaStream_t s);
int launch_walsh(void *out, const float2 *coil, int nimg, int nc, int npatch, int nslices, int half_out, cudaStream_t s);

} // namespace tronb

struct tron_plan;
namespace tronb {
/* plan.cu */
GridLaunch make_grid_launch(const tron_plan *p, const void *d_samples, float2 *d_grid, int z0, int nb);
/* cgnr.cu: per-coil images of slices [z0, z0 + nb) into plan->d_coil (plain adjoint or CGNR) */
int run_percoil_batch(tron_plan *p, const void *d_in, int z0, int nb, cudaStream_t s);
size_t cg_part_doubles(int batch);
}

struct tron_plan {
    tron_config cfg;
    tron_geometry g;
    int device = 0;
    int nch = 0;                         /* channels this plan owns */
    int nslices = 0;                     /* slices this plan owns */
    int batch = 1;
    int percoil = 0;                     /* adjoint through per-coil images: Walsh combine and/or CGNR */
    cudaStream_t stream = nullptr;       /* compute */
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    cudaStream_t s_grid = nullptr, s_fft = nullptr;   /* low / high priority compute streams */
    cudaEvent_t ev_grid[2] = {nullptr, nullptr}, ev_fft[2] = {nullptr, nullptr}, ev_user = nullptr;
    int overlap = 0;                     /* gridding of batch k+1 overlaps the FFT passes of batch k */
    cudaEvent_t ev_in = nullptr, ev_done[2] = {nullptr, nullptr}, ev_t[4] = {nullptr, nullptr, nullptr, nullptr};
    tronb::KbParams kb;
    tronb::SpokeTables tabs;
    tronb::FftPlan fft;
    float *deapod_adj = nullptr, *deapod_fwd = nullptr;
    int *tile_order = nullptr, *tile_order8 = nullptr, *heavy_cells = nullptr;
    int *tile_order_rows = nullptr, *tile_order8_rows = nullptr;
    long long *grid_dbg = nullptr;       /* TRON_GRID_DEBUG: per-warp cycles of the last gridding launch */
    int nheavy = 0, heavy_r2 = -1;
    int *heavy_cells_big = nullptr; int nheavy_big = 0, heavy_r2_big = -1;
    int heavy_big = 0;                   /* which of the two lists this plan's launches use (fixed per plan: one summation order) */
    int2 *tile_win8 = nullptr; int *tile_sched8 = nullptr; int n_near8 = 0;   /* grid_tile.cu */
    tronb::SpokeTables tabs_d;           /* grid_tile.cu: sliding-window difference tables (empty: not in use) */
    tronb::ScatterPlan scat;             /* grid_scatter.cu: tiles accumulated in shared memory (not ready: not in use) */
    int chain = 0;                       /* slice groups per chain: the first is gridded in full, the others from differences */
    int zero_r2 = 0x7fffffff;            /* adjoint: cells beyond this squared radius never receive a sample */
    float2 *d_grid = nullptr, *d_tmp = nullptr;     /* batch work buffers */
    int work_slices = 0;                            /* slices per launch they hold (grown on demand, plan.cu) */
    int *fft_sync = nullptr; int fft_ring = 0;      /* single-launch FFT stage: counters, slices of d_tmp used as a ring */
    float2 *d_gridi = nullptr;                      /* forward, nc >= 32: channel-interleaved copy of the grid */
    float2 *d_coil = nullptr;                       /* per-coil images of a batch (Walsh combine, CGNR iterate x) */
    float2 *cg_r = nullptr, *cg_v = nullptr;        /* CGNR: residual and A p, [batch][npe1work][nro][nc] */
    float2 *cg_z = nullptr, *cg_p = nullptr;        /* CGNR: B r and search direction, [batch][nx][nx][nc] */
    double *cg_part = nullptr;                      /* CGNR: partial inner products */
    void *d_in = nullptr, *d_out = nullptr;         /* device staging for the host API */
    size_t in_bytes = 0, out_bytes = 0;
    size_t in_elem_bytes = 8, out_elem_bytes = 8;
    int last_launches = 0;
    int stage_timing = 0;                /* TRON_STAGE_TIMING: per-stage events (diagnostic) */
    float last_ms[3] = {0, 0, 0};
};
