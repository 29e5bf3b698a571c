/*
 * plan.cu -- geometry, plan object and the host pipeline of libtron_b200.
 *
 * Replaces the reference's host layer:
 *   geometry derivation in main()        /root/reference/src/tron.cu:905-961
 *   tron_init / tron_shutdown            tron.cu:579-620
 *   tron_nufft_adj_radial2d / _radial2d  tron.cu:623-649
 *   recon_radial2d (slice loop)          tron.cu:726-786
 *
 * Differences in structure (not in results):
 *   - geometry lives in a plan, not in file-static globals;
 *   - the acquisition is uploaded ONCE and the sliding window (-d) is addressed
 *     on the device; the reference re-uploads every overlapping window from
 *     pageable memory (tron.cu:738-748);
 *   - slices are processed in batches per launch (3 launches per batch instead
 *     of 8 per slice), uploads/downloads overlap compute on separate streams;
 *   - buffers, tables and streams are created once per plan, not per call.
 */
#include "tron_internal.h"

#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <vector>

namespace tronb {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap; va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line)
{
    set_error("CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
    return e == cudaErrorMemoryAllocation ? TRON_ENOMEM : TRON_ECUDA;
}

} // namespace tronb

using namespace tronb;

extern "C" const char *tron_last_error(void) { return tronb::g_err; }
extern "C" int tron_version(void) { return TRON_B200_VERSION; }

extern "C" void tron_config_defaults(tron_config *c)
{
    memset(c, 0, sizeof *c);
    c->gridos = 2.f; c->kernwidth = 2.f; c->data_undersamp = 1.f;     /* tron.cu:67-69 */
    c->device = -1;
}

/* tron.cu:905-961, same int/float conversions and truncations */
extern "C" int tron_geometry_compute(const tron_config *c, tron_geometry *g)
{
    memset(g, 0, sizeof *g);
    if (c->niter < 0) { set_error("niter must be >= 0"); return TRON_EINVAL; }
    if (c->niter > 0 && !c->adjoint) { set_error("-i (CGNR) applies to the adjoint direction only (tron.cu:753-755)"); return TRON_EINVAL; }
    if (c->coil_combine != 0 && c->coil_combine != 1) { set_error("coil_combine must be 0 (sum of squares) or 1 (Walsh)"); return TRON_EINVAL; }
    if (c->coil_combine == 1 && (c->walsh_npatch < 0 || c->walsh_npatch > 16)) { set_error("walsh_npatch must be in [0, 16]"); return TRON_EINVAL; }
    if (c->koosh) { set_error("-3 (koosh ball) has no kernels in the reference either"); return TRON_EUNSUPPORTED; }
    if (!(c->gridos > 0.f) || !(c->kernwidth > 0.f) || !(c->data_undersamp > 0.f)) { set_error("gridos, kernwidth and data_undersamp must be positive"); return TRON_EINVAL; }
    for (int i = 0; i < 5; ++i) if (c->dims[i] == 0 || c->dims[i] > 0x7fffffffULL) { set_error("dims[%d] = %llu out of range", i, (unsigned long long)c->dims[i]); return TRON_EINVAL; }
    {   /* five 31-bit factors can wrap 64 bits: the payload-size checks downstream rely on these products */
        unsigned __int128 prod = 1;
        for (int i = 0; i < 5; ++i) prod *= c->dims[i];
        if (prod > ((unsigned __int128)1 << 60)) { set_error("the array described by dims holds more than 2^60 elements"); return TRON_EINVAL; }
    }
    g->nc = (int)c->dims[0]; g->nt = (int)c->dims[1];
    g->out_dims[0] = 1;                                    /* tron.cu:899 */
    int slide = c->prof_slide;
    if (c->adjoint) {
        g->nro = (int)c->dims[2]; g->npe1 = (int)c->dims[3]; g->npe2 = (int)c->dims[4];
        g->nx = g->nro / 2; g->ny = g->nro / 2;
        g->nxos = (int)(g->nx * c->gridos); g->nyos = (int)(g->ny * c->gridos);
        if ((float)g->npe1 <= (float)g->nro * c->data_undersamp) g->npe1work = g->npe1;
        else g->npe1work = (int)((float)g->nro * c->data_undersamp);
        if (g->npe1work < 1) { set_error("npe1work = %d", g->npe1work); return TRON_EINVAL; }
        if (slide == 0) slide = g->npe1work;
        if (slide < 0) { set_error("prof_slide must be >= 0"); return TRON_EINVAL; }
        g->nz = 1 + (g->npe1 - g->npe1work) / slide;
        g->out_dims[1] = (uint64_t)g->nt; g->out_dims[2] = (uint64_t)g->nx;
        g->out_dims[3] = (uint64_t)g->ny; g->out_dims[4] = (uint64_t)g->nz;
        g->in_elems = (uint64_t)g->nc * g->nt * g->nro * g->npe1 * (uint64_t)g->npe2;
        g->out_elems = (uint64_t)g->nt * g->nx * g->ny * (uint64_t)g->nz;
    } else {
        g->nx = (int)c->dims[2]; g->ny = (int)c->dims[3]; g->nz = (int)c->dims[4];
        g->nxos = (int)(g->nx * c->gridos); g->nyos = (int)(g->ny * c->gridos);
        g->nro = (int)(c->gridos * g->nx);
        g->npe1work = (int)(c->data_undersamp * (float)g->nro);
        g->npe1 = g->npe1work; g->npe2 = 1;              /* prof_slide stays as given (tron.cu:936-961) */
        g->out_dims[1] = (uint64_t)g->nt; g->out_dims[2] = (uint64_t)g->nro;
        g->out_dims[3] = (uint64_t)g->npe1; g->out_dims[4] = (uint64_t)g->npe2;
        g->in_elems = (uint64_t)g->nc * g->nt * g->nx * g->ny * (uint64_t)g->nz;
        g->out_elems = (uint64_t)g->nc * g->nt * g->nro * g->npe1 * (uint64_t)g->npe2;
        if (g->nx != g->ny) { set_error("non-square images are not implemented (tron.cu:945)"); return TRON_EUNSUPPORTED; }
        if (g->nz != 1) { set_error("forward mode with dims[4] = %d: the reference reads every slice from offset 0 and overruns its output (tron.cu:750,776); only dims[4] = 1 is defined", g->nz); return TRON_EUNSUPPORTED; }
        if (g->npe1work < 1) { set_error("npe1work = %d", g->npe1work); return TRON_EINVAL; }
    }
    g->prof_slide = slide;
    if (!(g->nc % 2 == 0 || g->nc == 1)) { set_error("nc = %d: only a single or an even number of channels (tron.cu:963)", g->nc); return TRON_EINVAL; }
    if (g->nt != 1) { set_error("nt = %d: the reference builds its FFT plans for nc channels but runs nc*nt (tron.cu:599-601); only nt = 1 is defined", g->nt); return TRON_EUNSUPPORTED; }
    if (g->nx < 1 || g->nxos < g->nx || (g->nxos & 1)) { set_error("nx = %d, nxos = %d: need an even oversampled grid >= nx", g->nx, g->nxos); return TRON_EINVAL; }

    if (c->niter > 0 && g->nro != g->nxos) { set_error("-i (CGNR) needs nro == nxos (gridos 2): nro = %d, nxos = %d", g->nro, g->nxos); return TRON_EUNSUPPORTED; }
    if (c->adjoint && c->coil_combine == 1 && g->nc > 1 && (c->per_coil_out || c->sos_partial)) { set_error("the Walsh combine excludes per_coil_out and sos_partial"); return TRON_EINVAL; }

    g->slice_begin = c->slice_begin; g->slice_end = c->slice_end;
    if (g->slice_begin == 0 && g->slice_end == 0) g->slice_end = c->adjoint ? g->nz : 1;
    if (!c->adjoint) { g->slice_begin = 0; g->slice_end = 1; }
    if (g->slice_begin < 0 || g->slice_end > (c->adjoint ? g->nz : 1) || g->slice_begin >= g->slice_end) { set_error("bad slice shard [%d,%d) of %d", g->slice_begin, g->slice_end, g->nz); return TRON_EINVAL; }
    g->coil_begin = c->coil_begin; g->coil_end = c->coil_end;
    if (g->coil_begin == 0 && g->coil_end == 0) g->coil_end = g->nc;
    if (g->coil_begin < 0 || g->coil_end > g->nc || g->coil_begin >= g->coil_end) { set_error("bad coil shard [%d,%d) of %d", g->coil_begin, g->coil_end, g->nc); return TRON_EINVAL; }
    int nch = g->coil_end - g->coil_begin;
    if (nch != g->nc && c->adjoint && (c->niter > 0 || c->coil_combine == 1)) { set_error("CGNR and the Walsh combine need every coil of a slice: coil shards are not supported"); return TRON_EUNSUPPORTED; }
    if (g->nc > 1 && ((g->coil_begin & 1) || (nch & 1))) { set_error("coil shards must start at an even channel and hold an even count"); return TRON_EINVAL; }
    /* A shard of the coils writes only its own channels of a channel-interleaved output (forward samples,
     * per-coil images) and cannot form the root of a partial sum: the one defined coil-sharded product is the
     * partial sum of squares, reduced across GPUs by tron_coil_reduce (comm.cu).  A forward transform is
     * coil-sharded by cutting dims[0] of its input instead (the channels are independent, tron.cu:540-577). */
    if (nch != g->nc && !(c->adjoint && c->sos_partial && !c->per_coil_out)) { set_error("coil shards [%d,%d) of %d need adjoint + sos_partial (partial sum of squares for tron_coil_reduce); shard a forward transform by its input's dims[0]", g->coil_begin, g->coil_end, g->nc); return TRON_EUNSUPPORTED; }

    uint64_t spoke = (uint64_t)g->nc * g->nt * g->nro;
    if (c->adjoint) {
        int ns = g->slice_end - g->slice_begin;
        g->shard_in_offset = spoke * (uint64_t)g->slice_begin * slide;
        g->shard_in_elems = spoke * ((uint64_t)(ns - 1) * slide + g->npe1work);
        uint64_t per = (uint64_t)g->nx * g->ny * (c->per_coil_out ? (uint64_t)g->nc : 1);
        g->shard_out_offset = per * g->slice_begin;
        g->shard_out_elems = per * ns;
    } else {
        g->shard_in_offset = 0; g->shard_in_elems = g->in_elems;
        g->shard_out_offset = 0; g->shard_out_elems = g->out_elems;
    }
    return TRON_OK;
}

static void plan_release(tron_plan *p)
{
    if (!p) return;
    cudaFree(p->tabs_d.gx); cudaFree(p->tabs_d.lut);
    scatter_plan_free(p->scat);
    cudaFree(p->tabs.cs); cudaFree(p->tabs.pe); cudaFree(p->tabs.gx); cudaFree(p->tabs.lut); cudaFree(p->tabs.cs_lin); cudaFree(p->tabs.cells);
    fft_plan_free(p->fft);
    cudaFree(p->deapod_adj); cudaFree(p->deapod_fwd); cudaFree(p->tile_order); cudaFree(p->tile_order8); cudaFree(p->tile_order_rows); cudaFree(p->tile_order8_rows); cudaFree(p->heavy_cells); cudaFree(p->heavy_cells_big); cudaFree(p->grid_dbg); cudaFree(p->tile_win8); cudaFree(p->tile_sched8);
    cudaFree(p->d_grid); cudaFree(p->d_tmp); cudaFree(p->d_gridi); cudaFree(p->d_in); cudaFree(p->d_out);
    cudaFree(p->fft_sync);
    cudaFree(p->d_coil); cudaFree(p->cg_r); cudaFree(p->cg_v); cudaFree(p->cg_z); cudaFree(p->cg_p); cudaFree(p->cg_part);
    if (p->stream) cudaStreamDestroy(p->stream);
    if (p->copy_in) cudaStreamDestroy(p->copy_in);
    if (p->copy_out) cudaStreamDestroy(p->copy_out);
    if (p->ev_in) cudaEventDestroy(p->ev_in);
    if (p->s_grid) cudaStreamDestroy(p->s_grid);
    if (p->s_fft) cudaStreamDestroy(p->s_fft);
    for (int i = 0; i < 2; ++i) { if (p->ev_grid[i]) cudaEventDestroy(p->ev_grid[i]); if (p->ev_fft[i]) cudaEventDestroy(p->ev_fft[i]); }
    if (p->ev_user) cudaEventDestroy(p->ev_user);
    for (int i = 0; i < 2; ++i) if (p->ev_done[i]) cudaEventDestroy(p->ev_done[i]);
    for (int i = 0; i < 4; ++i) if (p->ev_t[i]) cudaEventDestroy(p->ev_t[i]);
    delete p;
}

#define HOST_BATCH_MAX 64

static int pick_batch(const tron_plan *p)
{
    if (p->cfg.batch_slices > 0) return p->cfg.batch_slices < p->nslices ? p->cfg.batch_slices : p->nslices;
    const char *e = getenv("TRON_BATCH");
    if (e && atoi(e) > 0) return atoi(e) < p->nslices ? atoi(e) : p->nslices;
    size_t per = (size_t)p->nch * p->g.nxos * ((size_t)p->g.nxos + p->g.nx) * sizeof(float2);
    if (p->percoil) per += (size_t)p->g.nc * p->g.nx * p->g.nx * sizeof(float2);
    if (p->cfg.niter > 0) per += 2 * (size_t)p->g.nc * ((size_t)p->g.nx * p->g.nx + (size_t)p->g.nro * p->g.npe1work) * sizeof(float2);
    /* launches of >= 32 slices reach the kernels' asymptotic throughput (profiles/r01_grid_only_timing.txt);
     * longer launches still shave the drain of each kernel (measured on cfg2, device resident: 10.31 ms per
     * step at 64 slices per launch, 10.02 at 128); the work buffers are bounded to ~6 GB of the 180 GB.
     * The host pipeline caps its launches at HOST_BATCH_MAX (copy/compute overlap wants them shorter). */
    /* Round 2: with the scatter kernel's (tile, chain) tasks a 512-slice launch has a shorter tail than two of 256 and
     * allows chains of 64 (7.8 instead of 8.0-8.1 ms per cfg2 step; TRON_BATCH x TRON_SCATTER_CHAIN sweep in
     * profiles/r02_batch_chain_sweep.txt): up to 512 slices within 12 GB. */
    size_t budget = (size_t)12288 << 20, freeb = 0, totalb = 0;
    if (cudaMemGetInfo(&freeb, &totalb) == cudaSuccess && freeb / 3 < budget) budget = freeb / 3;   /* shared GPUs */
    size_t b = budget / (per ? per : 1);
    if (b < 1) b = 1;
    if (b > 512) b = 512;
    if ((int)b > p->nslices) b = p->nslices;
    return (int)b;
}

/* the FFT, degridding and combine launches put slices x channels (or slices) into gridDim.y (<= 65535) */
static int clamp_batch_to_launch_limits(const tron_plan *p, int batch)
{
    const int per = p->nch > 0 ? p->nch : 1;
    const int lim = 65535 / per;
    if (batch > lim) batch = lim;
    return batch < 1 ? 1 : batch;
}

/* oversampled grid + pass-A output for `slices` slices per launch (never shrinks) */
static int ensure_work_buffers(tron_plan *p, int slices)
{
    if (slices <= p->work_slices) return TRON_OK;
    const int n = p->g.nxos;
    cudaFree(p->d_grid); cudaFree(p->d_tmp);
    p->d_grid = nullptr; p->d_tmp = nullptr; p->work_slices = 0;
    TRON_CUDA(cudaMalloc(&p->d_grid, (size_t)(p->overlap ? 2 : 1) * slices * p->nch * n * n * sizeof(float2)));
    TRON_CUDA(cudaMalloc(&p->d_tmp, (size_t)slices * p->nch * n * p->g.nx * sizeof(float2)));
    p->work_slices = slices;
    return TRON_OK;
}

extern "C" int tron_plan_create(tron_plan **out, const tron_config *cfg)
{
    *out = nullptr;
    tron_geometry g;
    int rc = tron_geometry_compute(cfg, &g);
    if (rc) return rc;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        cudaGetLastError();
        set_error("no CUDA device: libtron_b200 has no CPU path");
        return TRON_ENODEV;
    }
    int cur = 0;
    cudaGetDevice(&cur);
    if (cfg->device >= ndev) { set_error("device %d of %d", cfg->device, ndev); return TRON_ENODEV; }
    /* -g (tron.cu:837-839) picks the plan's device; the caller's current device is restored on return
     * (the CLI, like the reference's main(), selects it for the process itself) */
    DeviceGuard guard(cfg->device >= 0 ? cfg->device : cur);
    tron_plan *p = new tron_plan();
    p->cfg = *cfg; p->g = g;
    cudaGetDevice(&p->device);
    p->nch = g.coil_end - g.coil_begin;
    p->nslices = g.slice_end - g.slice_begin;
    p->kb = make_kb(cfg->kernwidth);              /* plan-time polynomial fit, refmath.cuh */
    /* paths that need every coil image of a slice at once (combine.cu, cgnr.cu) */
    p->percoil = cfg->adjoint && (cfg->niter > 0 || (cfg->coil_combine == 1 && g.nc > 1));
    p->in_elem_bytes = cfg->half_in ? 4 : 8;
    p->out_elem_bytes = cfg->half_out ? 4 : 8;
    if (cfg->adjoint && cfg->sos_partial && g.nc > 1) p->out_elem_bytes = 4;
    p->in_bytes = g.shard_in_elems * p->in_elem_bytes;
    p->out_bytes = g.shard_out_elems * p->out_elem_bytes;

    /* TRON_PLAN_TRACE: wall-clock milliseconds of each stage of plan creation on stderr */
    const bool ptrace = getenv("TRON_PLAN_TRACE") != nullptr;
    struct timespec pt0; clock_gettime(CLOCK_MONOTONIC, &pt0);
    auto pmark = [&](const char *what) {
        if (!ptrace) return;
        cudaDeviceSynchronize();
        struct timespec t1; clock_gettime(CLOCK_MONOTONIC, &t1);
        fprintf(stderr, "tron plan trace: %-28s %8.3f ms\n", what, (t1.tv_sec - pt0.tv_sec) * 1e3 + (t1.tv_nsec - pt0.tv_nsec) * 1e-6);
        pt0 = t1;
    };
#define PLAN_TRY(call) do { int rc__ = (call); if (rc__) { plan_release(p); return rc__; } } while (0)
#define PLAN_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { \
        int rc__ = cuda_fail(e__, #call, __FILE__, __LINE__); plan_release(p); return rc__; } } while (0)

    PLAN_CUDA(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    PLAN_CUDA(cudaStreamCreateWithFlags(&p->copy_in, cudaStreamNonBlocking));
    PLAN_CUDA(cudaStreamCreateWithFlags(&p->copy_out, cudaStreamNonBlocking));
    PLAN_CUDA(cudaEventCreateWithFlags(&p->ev_in, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) PLAN_CUDA(cudaEventCreateWithFlags(&p->ev_done[i], cudaEventDisableTiming));
    {
        int least = 0, greatest = 0;
        PLAN_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        PLAN_CUDA(cudaStreamCreateWithPriority(&p->s_grid, cudaStreamNonBlocking, least));
        PLAN_CUDA(cudaStreamCreateWithPriority(&p->s_fft, cudaStreamNonBlocking, greatest));
        for (int i = 0; i < 2; ++i) {
            PLAN_CUDA(cudaEventCreateWithFlags(&p->ev_grid[i], cudaEventDisableTiming));
            PLAN_CUDA(cudaEventCreateWithFlags(&p->ev_fft[i], cudaEventDisableTiming));
        }
        PLAN_CUDA(cudaEventCreateWithFlags(&p->ev_user, cudaEventDisableTiming));
    }
    for (int i = 0; i < 4; ++i) PLAN_CUDA(cudaEventCreate(&p->ev_t[i]));
    pmark("streams and events");

    const int n = g.nxos;
    if (cfg->adjoint) {
        /* golden angles + overlapping windows: 4 consecutive slices share their taps (grid.cu) */
        int gs = 1;
        if (cfg->golden_angle && p->nslices > 1 && 2 * g.prof_slide <= g.npe1work) gs = 4;
        const char *eg = getenv("TRON_GROUP");
        if (eg && cfg->golden_angle && (atoi(eg) == 1 || atoi(eg) == 4)) gs = atoi(eg);
        if (cfg->niter > 0) gs = 1;                       /* residuals of overlapping windows are not shared */
        int nun = (gs - 1) * g.prof_slide + g.npe1work;                 /* union window of a group */
        int ntab = cfg->golden_angle ? (p->nslices + gs - 1) / gs : 1;
        int skip = cfg->skip_angles + (cfg->golden_angle ? g.slice_begin * g.prof_slide : 0);
        PLAN_TRY(launch_build_tables(p->tabs, nun, g.npe1work, ntab, gs * g.prof_slide, skip, cfg->golden_angle, 1,
                                     g.npe1work, g.prof_slide, gs, p->nslices, n, cfg->kernwidth, p->stream));
    } else {
        PLAN_TRY(launch_build_tables(p->tabs, g.npe1work, g.npe1work, 1, 0, cfg->skip_angles, cfg->golden_angle, 0,
                                     g.npe1work, 0, 1, 1, n, cfg->kernwidth, p->stream));
    }
    pmark("spoke and cell tables");
    PLAN_TRY(fft_plan_init(p->fft, n, g.nx));
    pmark("fft plan");
    if (cfg->adjoint && p->fft.pow2 && !getenv("TRON_NO_ZERO_SKIP")) {
        /* a cell receives samples only if ceil(R - W) <= nxos/2 - 1 (tron.cu:498-502): beyond
         * R = nxos/2 - 1 + W (+0.5 of margin for the float hypot) the grid is identically zero; the
         * gridding kernel skips those stores and FFT pass A does not fetch them (same integer test) */
        const float rz = (float)(n / 2 - 1) + cfg->kernwidth + 0.5f;
        const double r2 = (double)rz * (double)rz;
        p->zero_r2 = r2 < 2.0e9 ? (int)r2 : 0x7fffffff;
    }
    if (cfg->adjoint) {
        PLAN_TRY(build_tile_order(&p->tile_order, n, 16, false));
        PLAN_TRY(build_tile_order(&p->tile_order8, n, 8, false));
        PLAN_TRY(build_tile_order(&p->tile_order_rows, n, 16, true));
        PLAN_TRY(build_tile_order(&p->tile_order8_rows, n, 8, true));
        /* two heavy-cell lists: launches with little other work need more of the DC neighbourhood on the
         * warp-per-cell path than long ones (grid.cu: build_heavy_cells, launch_grid_cg picks per launch) */
        PLAN_TRY(build_heavy_cells(&p->heavy_cells, &p->nheavy, &p->heavy_r2, n, p->tabs.npe, cfg->kernwidth, 24));
        PLAN_TRY(build_heavy_cells(&p->heavy_cells_big, &p->nheavy_big, &p->heavy_r2_big, n, p->tabs.npe, cfg->kernwidth, 96));
        /* enough cell-groups per launch to hide the long cells: only the innermost ones take the warp path.
         * Decided once per plan (from the launch length the device pipeline uses), never per launch. */
        {
            const int gsz = p->tabs.gs > 0 ? p->tabs.gs : 1;
            const int per_launch = p->nslices < 256 ? p->nslices : 256;
            p->heavy_big = (double)n * n * ((per_launch + gsz - 1) / gsz) >= 2.0e6;
            if (getenv("TRON_HEAVY_BIG")) p->heavy_big = atoi(getenv("TRON_HEAVY_BIG")) != 0;
        }
        pmark("tile orders, heavy cells");
        PLAN_TRY(build_tile_windows(&p->tile_win8, p->tabs.cells, n, p->tabs.nbins, 8, 4, p->stream));
        /* sliding windows: all but the first slice group of a chain are gridded from what enters and leaves
         * the window (grid_tile.cu); worth it when that is clearly fewer spokes than a group's union window */
        {
            const int gsz = p->tabs.gs, ne = 2 * gsz * g.prof_slide;
            const bool plain = p->kb.fast && g.nro == g.nxos && p->nch == g.nc && (g.nc == 2 || g.nc == 4 || g.nc == 6 || g.nc == 8);
            const char *ed = getenv("TRON_TILE_DELTA");
            if (cfg->golden_angle && gsz == 4 && plain && p->nslices > gsz && 5 * ne <= 4 * p->tabs.npe
                && !getenv("TRON_NO_TILE") && !(ed && atoi(ed) == 0)) {
                p->chain = getenv("TRON_TILE_GPER") && atoi(getenv("TRON_TILE_GPER")) > 0 ? atoi(getenv("TRON_TILE_GPER")) : 8;
                const int skip = cfg->skip_angles + g.slice_begin * g.prof_slide;
                PLAN_TRY(build_delta_tables(p->tabs_d, p->tabs, p->tabs.ntab, gsz * g.prof_slide, skip, g.npe1work,
                                            g.prof_slide, gsz, p->nslices, p->stream));
            }
        }
        pmark("tile windows, delta tables");
        /* tiles accumulated in shared memory, sample driven (grid_scatter.cu): the default where it applies */
        {
            const bool fits = p->kb.fast && cfg->kernwidth == 2.0f && g.nro == g.nxos && p->nch == g.nc
                              && (g.nc == 2 || g.nc == 4 || g.nc == 6 || g.nc == 16) && n % 16 == 0 && n <= 4096 && g.npe1work <= 4096;
            if (fits && !getenv("TRON_NO_SCATTER")) {
                const int skip = cfg->skip_angles + (cfg->golden_angle ? g.slice_begin * g.prof_slide : 0);
                PLAN_TRY(scatter_plan_build(p->scat, p->tabs.cells, p->tabs.nbins, n, p->nslices, g.npe1work, g.prof_slide,
                                            skip, cfg->golden_angle, cfg->kernwidth, g.nc, p->stream));
            }
        }
        PLAN_TRY(build_tile_schedule(&p->tile_sched8, &p->n_near8, n, 8,
                                     getenv("TRON_TILE_NEAR") ? (float)atof(getenv("TRON_TILE_NEAR")) : 32.f));
    }
    pmark("scatter plan, tile schedule");
    PLAN_CUDA(cudaMalloc(&p->deapod_adj, (size_t)g.nx * g.nx * sizeof(float)));
    PLAN_CUDA(cudaMalloc(&p->deapod_fwd, (size_t)g.nx * g.nx * sizeof(float)));
    PLAN_TRY(launch_deapod_tables(p->deapod_adj, p->deapod_fwd, g.nx, n, cfg->kernwidth, cfg->gridos, p->stream));

    pmark("deapodisation tables");
    p->batch = cfg->adjoint ? clamp_batch_to_launch_limits(p, pick_batch(p)) : 1;
    if (p->nch > 65535) { set_error("nc = %d channels exceed the launch limits (65535)", p->nch); plan_release(p); return TRON_EUNSUPPORTED; }
    if (cfg->kernwidth > 7.5f) { set_error("kernel half-width %.2f > 7.5: outside the kernels' tap windows", cfg->kernwidth); plan_release(p); return TRON_EUNSUPPORTED; }
    if (cfg->adjoint && cfg->coil_combine == 1 && g.nc > 128) { set_error("the Walsh combine supports at most 128 channels (nc = %d)", g.nc); plan_release(p); return TRON_EUNSUPPORTED; }
    if (cfg->adjoint && p->tabs.gs > 1) {                /* launches start on group (chain) boundaries */
        int q = p->tabs.gs * (p->chain > 0 ? p->chain : 1);
        if (p->scat.ready && p->scat.chain > q) q = p->scat.chain;
        p->batch = ((p->batch + q - 1) / q) * q;
        while (p->batch > q && (long long)p->batch * p->nch > 65535) p->batch -= q;
    }
    p->stage_timing = getenv("TRON_STAGE_TIMING") != nullptr;
    if (cfg->adjoint && getenv("TRON_GRID_DEBUG")) {
        PLAN_CUDA(cudaMalloc(&p->grid_dbg, (size_t)8 * 65536 * 8 * sizeof(long long)));
        PLAN_CUDA(cudaMemset(p->grid_dbg, 0, (size_t)8 * 65536 * 8 * sizeof(long long)));
    }
    p->overlap = cfg->adjoint && !p->percoil && p->nslices > p->batch && getenv("TRON_OVERLAP") != nullptr;   /* measured slower on B200: off */
    /* batch work buffers: sized for the host pipeline's launches now (<= HOST_BATCH_MAX slices: the cold span
     * plan + recon + destroy then allocates and frees ~1.2 GB instead of ~4.8 GB), grown to the device path's
     * launch length (p->batch) on its first use (ensure_work_buffers) */
    {
        int first = p->batch;
        if (cfg->adjoint && !p->percoil && first > HOST_BATCH_MAX) {
            int q = 1;
            if (p->tabs.gs > 1) { q = p->tabs.gs * (p->chain > 0 ? p->chain : 1); if (p->scat.ready && p->scat.chain > q) q = p->scat.chain; }
            first = (HOST_BATCH_MAX / q) * q;
            if (first < q) first = q;
            if (first > p->batch) first = p->batch;
        }
        int rc_w = ensure_work_buffers(p, first);
        if (rc_w) { plan_release(p); return rc_w; }
    }
    if (cfg->adjoint && !p->percoil && !p->overlap && getenv("TRON_FFT_FUSED")) {
        /* single-launch FFT stage (fft.cu: p2w_adj_fused): the intermediate lives in a ring of slices small
         * enough for the L2.  Measured on cfg2: HBM traffic of the stage 1605 -> 715 MB per 64 slices, but
         * 10.6-11.0 instead of 10.3 ms per step (the merged kernel is issue bound at 4 blocks/SM): off by default */
        const size_t slot = (size_t)p->nch * n * g.nx * sizeof(float2);
        int ring = getenv("TRON_FFT_RING") ? atoi(getenv("TRON_FFT_RING")) : (int)(((size_t)48 << 20) / (slot ? slot : 1));
        if (ring < 2) ring = 2;
        if (ring > p->batch) ring = p->batch;
        p->fft_ring = ring;
        PLAN_CUDA(cudaMalloc(&p->fft_sync, 2 * (size_t)p->batch * sizeof(int)));
    }
    if (!cfg->adjoint && ((p->nch >= 32 && p->nch % 32 == 0) || ((p->nch == 8 || p->nch == 16) && cfg->kernwidth >= 3.f)))
        PLAN_CUDA(cudaMalloc(&p->d_gridi, (size_t)p->nch * n * n * sizeof(float2)));
    if (p->percoil) {
        const size_t N = (size_t)p->batch * g.nc * g.nx * g.nx, ns = (size_t)p->batch * g.nc * g.nro * g.npe1work;
        PLAN_CUDA(cudaMalloc(&p->d_coil, N * sizeof(float2)));
        if (cfg->niter > 0) {
            PLAN_CUDA(cudaMalloc(&p->cg_z, N * sizeof(float2)));
            PLAN_CUDA(cudaMalloc(&p->cg_p, N * sizeof(float2)));
            PLAN_CUDA(cudaMalloc(&p->cg_r, ns * sizeof(float2)));
            PLAN_CUDA(cudaMalloc(&p->cg_v, ns * sizeof(float2)));
            PLAN_CUDA(cudaMalloc(&p->cg_part, cg_part_doubles(p->batch) * sizeof(double)));
        }
    }
    PLAN_CUDA(cudaStreamSynchronize(p->stream));
    pmark("work buffers");
#undef PLAN_TRY
#undef PLAN_CUDA
    *out = p;
    return TRON_OK;
}

extern "C" int tron_plan_destroy(tron_plan *p)
{
    if (!p) return TRON_OK;
    DeviceGuard guard(p->device);
    cudaDeviceSynchronize();
    plan_release(p);
    return TRON_OK;
}

extern "C" int tron_plan_geometry(const tron_plan *p, tron_geometry *g)
{
    if (!p || !g) { set_error("null plan"); return TRON_EINVAL; }
    *g = p->g;
    return TRON_OK;
}

namespace tronb {
GridLaunch make_grid_launch(const tron_plan *p, const void *d_samples, float2 *d_grid, int z0, int nb)
{
    const tron_geometry &g = p->g;
    GridLaunch L;
    L.samples = d_samples; L.grid = d_grid;
    L.tab_cs = p->tabs.cs; L.tab_pe = p->tabs.pe; L.tab_gx = p->tabs.gx; L.lut = p->tabs.lut; L.cells = p->tabs.cells;
    L.tile_order = p->tile_order; L.tile_order8 = p->tile_order8;
    L.tile_order_rows = p->tile_order_rows; L.tile_order8_rows = p->tile_order8_rows;
    if (p->heavy_big) { L.heavy_cells = p->heavy_cells_big; L.nheavy = p->nheavy_big; L.heavy_r2 = p->heavy_r2_big; }
    else { L.heavy_cells = p->heavy_cells; L.nheavy = p->nheavy; L.heavy_r2 = p->heavy_r2; }
    L.tile_win8 = p->tile_win8; L.tile_sched8 = p->tile_sched8; L.n_near8 = p->n_near8;
    L.scat = p->scat.ready ? &p->scat : nullptr;
    L.tab_gx_d = p->tabs_d.gx; L.lut_d = p->tabs_d.lut; L.npe_d = p->tabs_d.npe; L.chain = p->chain;
    L.heavy_cells_big = p->heavy_cells_big; L.nheavy_big = p->nheavy_big; L.heavy_r2_big = p->heavy_r2_big;
    L.tab_per_slice = p->tabs.ntab > 1 ? 1 : 0;
    L.nbins = p->tabs.nbins;
    L.n = g.nxos; L.nro = g.nro; L.npe = p->tabs.npe; L.gs = p->tabs.gs; L.ngroups = 0;
    L.nc_total = g.nc * g.nt; L.ch0 = g.coil_begin; L.nch = p->nch;
    L.z0 = z0; L.nslices = nb; L.slide = g.prof_slide;
    L.kb = p->kb;
    /* tron.cu:408-409 and 532 */
    L.sdc_a = (2.f - 2.f / (float)g.npe1work) / (float)g.nro;
    L.sdc_b = 1.f / (float)g.npe1work;
    L.scale = 1.f / (float)g.nxos / (float)g.npe1work;
    L.sdc_as = L.sdc_a * L.scale; L.sdc_bs = L.sdc_b * L.scale;
    L.half_in = p->cfg.half_in;
    L.dbg = p->grid_dbg;
    L.zero_r2 = 0x7fffffff;                              /* stage API: every cell is stored */
    return L;
}
} // namespace tronb

static int adjoint_mode(const tron_plan *p)
{
    if (p->cfg.per_coil_out) return 2;
    if (p->g.nc == 1) return 1;                      /* tron.cu:265-266 */
    return p->cfg.sos_partial ? 3 : 0;
}

/* the two halves of one batch of adjoint slices */
static int launch_batch_grid(tron_plan *p, const void *d_in, float2 *d_grid, int z0, int nb, cudaStream_t s)
{
    GridLaunch L = make_grid_launch(p, d_in, d_grid, z0, nb);
    L.zero_r2 = p->zero_r2;
    p->last_launches += 1;
    return launch_grid(L, s);
}

static int launch_batch_fft(tron_plan *p, void *d_out, const float2 *d_grid, int z0, int nb, cudaStream_t s)
{
    const tron_geometry &g = p->g;
    AdjFftLaunch a;
    a.grid = d_grid; a.tmp = p->d_tmp; a.deapod = p->deapod_adj;
    a.nslices = nb; a.nch = p->nch; a.nc_total = g.nc * g.nt; a.ch0 = g.coil_begin;
    a.mode = adjoint_mode(p); a.half_out = p->cfg.half_out;
    a.zero_r2 = p->zero_r2;
    a.sync = p->fft_sync; a.ring = p->fft_ring;
    size_t per = (size_t)g.nx * g.ny * (a.mode == 2 ? (size_t)g.nc : 1);
    a.out = (char *)d_out + (size_t)z0 * per * p->out_elem_bytes;
    p->last_launches += adj_fft_single_launch(p->fft, a) ? 1 : 2;
    return launch_adj_fft(p->fft, a, s);
}

/* One batch through the per-coil path: per-coil images (plain adjoint or CGNR, cgnr.cu), then the
 * coil combine the configuration asks for (combine.cu). */
static int launch_batch_percoil(tron_plan *p, void *d_out, const void *d_in, int z0, int nb, cudaStream_t s)
{
    const tron_geometry &g = p->g;
    int rc = run_percoil_batch(p, d_in, z0, nb, s);
    if (rc) return rc;
    const int mode = adjoint_mode(p);
    const size_t per = (size_t)g.nx * g.ny * (mode == 2 ? (size_t)g.nc : 1);
    void *out = (char *)d_out + (size_t)z0 * per * p->out_elem_bytes;
    p->last_launches += 1;
    if (p->cfg.coil_combine == 1 && g.nc > 1)
        return launch_walsh(out, p->d_coil, g.nx, g.nc, p->cfg.walsh_npatch, nb, p->cfg.half_out, s);
    return launch_coil_combine(out, p->d_coil, (size_t)nb * g.nx * g.ny, g.nc, mode, p->cfg.half_out, s);
}

/* All adjoint slices of the plan.  Gridding (instruction-issue bound) runs on a low-priority
 * stream, the FFT passes (shared-memory / HBM bound) of the previous batch on a high-priority
 * one, so blocks of both kinds are resident on the SMs at the same time; the oversampled grid is
 * double buffered.  Host mode (h_in != NULL) uploads each spoke once, in the order the batches
 * need them, and downloads trail the FFT stream.  `user` (device mode) is made to wait for the
 * start of its own prior work and for the end of ours. */
static int run_adjoint_all(tron_plan *p, void *d_out, const void *d_in, cudaStream_t user,
                           const void *h_in, void *h_out)
{
    const tron_geometry &g = p->g;
    const bool host = h_in != nullptr;
    const bool overlap = p->overlap && !p->stage_timing;
    cudaStream_t sg = overlap ? p->s_grid : (host ? p->stream : user);
    cudaStream_t sf = overlap ? p->s_fft : sg;
    if (!host && overlap) {
        TRON_CUDA(cudaEventRecord(p->ev_user, user));
        TRON_CUDA(cudaStreamWaitEvent(sg, p->ev_user, 0));
        TRON_CUDA(cudaStreamWaitEvent(sf, p->ev_user, 0));
    }
    const size_t spoke_bytes = (size_t)g.nc * g.nt * g.nro * p->in_elem_bytes;
    const size_t slice_out_bytes = (size_t)g.nx * g.ny * (p->cfg.per_coil_out ? (size_t)g.nc : 1) * p->out_elem_bytes;
    size_t spokes_up = 0;
    int i = 0;
    int gs = (p->tabs.gs > 0 ? p->tabs.gs : 1) * (p->chain > 0 ? p->chain : 1);   /* launch granularity */
    if (p->scat.ready && p->scat.chain > gs) gs = p->scat.chain;
    /* slices per launch: the plan's batch on resident data, at most HOST_BATCH_MAX (TRON_HOST_BATCH) in host mode */
    static const int hbm_env = getenv("TRON_HOST_BATCH") ? atoi(getenv("TRON_HOST_BATCH")) : 0;
    const int hbm = hbm_env >= gs ? hbm_env : HOST_BATCH_MAX;
    const int hb = host ? (p->batch < hbm ? p->batch : ((hbm / gs) * gs > 0 ? (hbm / gs) * gs : gs)) : p->batch;
    {
        const int rc_w = ensure_work_buffers(p, hb < p->batch ? hb : p->batch);
        if (rc_w) return rc_w;
    }
    const size_t grid_elems = (size_t)p->work_slices * p->nch * g.nxos * g.nxos;
    /* TRON_HOST_TRACE: when did the last upload, the last kernel and the last download finish? */
    static const bool trace = getenv("TRON_HOST_TRACE") != nullptr;
    cudaEvent_t tr[4] = {nullptr, nullptr, nullptr, nullptr};
    if (trace && host) {
        for (int k = 0; k < 4; ++k) cudaEventCreate(&tr[k]);
        cudaDeviceSynchronize();
        cudaEventRecord(tr[0], p->copy_in);
    }
    for (int z0 = 0, nb = 0; z0 < p->nslices; z0 += nb, i ^= 1) {
        nb = p->nslices - z0 < p->batch ? p->nslices - z0 : p->batch;
        if (host && nb > hb) nb = hb;
        if (host && hb >= 2 * gs) {
            /* host mode ramps the batch size up at the start and down at the end, so that the first
             * upload and the last download (which nothing overlaps) are short */
            int ramp = hb;
            if (z0 < hb) ramp = z0 == 0 ? hb / 8 : (z0 < hb / 2 ? hb / 4 : hb / 2);
            const int rem = p->nslices - z0;
            if (rem <= hb) ramp = rem > hb / 4 ? rem / 2 : rem;
            ramp = ((ramp + gs - 1) / gs) * gs;
            if (ramp >= gs && ramp < nb) nb = ramp;
        }
        float2 *gridbuf = p->d_grid + (overlap ? (size_t)i * grid_elems : 0);
        if (host) {
            size_t need = (size_t)(z0 + nb - 1) * g.prof_slide + g.npe1work;
            if (need > spokes_up) {
                TRON_CUDA(cudaMemcpyAsync((char *)d_in + spokes_up * spoke_bytes,
                                          (const char *)h_in + spokes_up * spoke_bytes,
                                          (need - spokes_up) * spoke_bytes, cudaMemcpyHostToDevice, p->copy_in));
                spokes_up = need;
                TRON_CUDA(cudaEventRecord(p->ev_in, p->copy_in));
                TRON_CUDA(cudaStreamWaitEvent(sg, p->ev_in, 0));
            }
        }
        if (overlap) TRON_CUDA(cudaStreamWaitEvent(sg, p->ev_fft[i], 0));   /* buffer i free again */
        if (p->stage_timing) cudaEventRecord(p->ev_t[0], sg);
        int rc = p->percoil ? launch_batch_percoil(p, d_out, d_in, z0, nb, sg)
                            : launch_batch_grid(p, d_in, gridbuf, z0, nb, sg);
        if (rc) return rc;
        if (p->stage_timing) cudaEventRecord(p->ev_t[1], sg);
        if (overlap) {
            TRON_CUDA(cudaEventRecord(p->ev_grid[i], sg));
            TRON_CUDA(cudaStreamWaitEvent(sf, p->ev_grid[i], 0));
        }
        if (!p->percoil) {
            rc = launch_batch_fft(p, d_out, gridbuf, z0, nb, sf);
            if (rc) return rc;
        }
        if (p->stage_timing) {                     /* diagnostic mode: serialises host and device */
            cudaEventRecord(p->ev_t[2], sf);
            cudaEventSynchronize(p->ev_t[2]);
            float t0 = 0, t1 = 0;
            cudaEventElapsedTime(&t0, p->ev_t[0], p->ev_t[1]);
            cudaEventElapsedTime(&t1, p->ev_t[1], p->ev_t[2]);
            p->last_ms[0] += t0; p->last_ms[1] += t1;
        }
        if (overlap || host) TRON_CUDA(cudaEventRecord(p->ev_fft[i], sf));
        if (host) {
            TRON_CUDA(cudaStreamWaitEvent(p->copy_out, p->ev_fft[i], 0));
            TRON_CUDA(cudaMemcpyAsync((char *)h_out + (size_t)z0 * slice_out_bytes,
                                      (char *)d_out + (size_t)z0 * slice_out_bytes,
                                      (size_t)nb * slice_out_bytes, cudaMemcpyDeviceToHost, p->copy_out));
        }
    }
    if (host) {
        if (trace) { cudaEventRecord(tr[1], p->copy_in); cudaEventRecord(tr[2], sf); cudaEventRecord(tr[3], p->copy_out); }
        TRON_CUDA(cudaStreamSynchronize(p->copy_out));
        TRON_CUDA(cudaStreamSynchronize(sf));
        TRON_CUDA(cudaStreamSynchronize(sg));
        if (trace) {
            float a = 0, b = 0, c = 0;
            cudaEventElapsedTime(&a, tr[0], tr[1]); cudaEventElapsedTime(&b, tr[0], tr[2]); cudaEventElapsedTime(&c, tr[0], tr[3]);
            fprintf(stderr, "tron host trace: uploads done %.3f ms, kernels done %.3f ms, downloads done %.3f ms\n", a, b, c);
            for (int k = 0; k < 4; ++k) cudaEventDestroy(tr[k]);
        }
    } else if (overlap) {
        TRON_CUDA(cudaEventRecord(p->ev_user, sf));
        TRON_CUDA(cudaStreamWaitEvent(user, p->ev_user, 0));
    }
    return TRON_OK;
}

static int run_degrid(tron_plan *p, const DegridLaunch &d, cudaStream_t s)
{
    if (p->d_gridi && degrid_wide_applicable(d) && !getenv("TRON_NO_WIDE")) return launch_degrid_wide(d, p->d_gridi, s);
    return launch_degrid(d, s);
}

static int run_forward(tron_plan *p, void *d_out, const void *d_in, cudaStream_t s)
{
    const tron_geometry &g = p->g;
    FwdFftLaunch f;
    f.img = d_in; f.tmp = p->d_tmp; f.grid = p->d_grid; f.deapod = p->deapod_fwd;
    f.nch = p->nch; f.nc_total = g.nc * g.nt; f.ch0 = g.coil_begin; f.half_in = p->cfg.half_in;
    int rc = launch_fwd_fft(p->fft, f, s);
    if (rc) return rc;
    DegridLaunch d;
    d.samples = d_out; d.grid = p->d_grid; d.cs = p->tabs.cs_lin;
    d.n = g.nxos; d.nro = g.nro; d.npe = g.npe1work;
    d.nc_total = g.nc * g.nt; d.ch0 = g.coil_begin; d.nch = p->nch;
    d.kb = p->kb; d.half_out = p->cfg.half_out;
    d.pair_spokes = !p->cfg.golden_angle;
    rc = run_degrid(p, d, s);
    p->last_launches += 3;
    return rc;
}

extern "C" int tron_recon_device(tron_plan *p, void *d_out, const void *d_in, void *stream)
{
    if (!p || !d_out || !d_in) { set_error("null argument"); return TRON_EINVAL; }
    DeviceGuard guard(p->device);
    cudaStream_t s = (cudaStream_t)stream;
    p->last_launches = 0;
    p->last_ms[0] = p->last_ms[1] = p->last_ms[2] = 0.f;
    if (!p->cfg.adjoint) return run_forward(p, d_out, d_in, s);
    return run_adjoint_all(p, d_out, d_in, s, nullptr, nullptr);
}

extern "C" int tron_recon_host(tron_plan *p, void *h_out, const void *h_in)
{
    if (!p || !h_out || !h_in) { set_error("null argument"); return TRON_EINVAL; }
    DeviceGuard guard(p->device);
    if (!p->d_in) TRON_CUDA(cudaMalloc(&p->d_in, p->in_bytes));
    if (!p->d_out) TRON_CUDA(cudaMalloc(&p->d_out, p->out_bytes));
    p->last_launches = 0;
    p->last_ms[0] = p->last_ms[1] = p->last_ms[2] = 0.f;
    if (!p->cfg.adjoint) {
        TRON_CUDA(cudaMemcpyAsync(p->d_in, h_in, p->in_bytes, cudaMemcpyHostToDevice, p->stream));
        int rc = run_forward(p, p->d_out, p->d_in, p->stream);
        if (rc) return rc;
        TRON_CUDA(cudaMemcpyAsync(h_out, p->d_out, p->out_bytes, cudaMemcpyDeviceToHost, p->stream));
        TRON_CUDA(cudaStreamSynchronize(p->stream));
        return TRON_OK;
    }
    return run_adjoint_all(p, p->d_out, p->d_in, nullptr, h_in, h_out);
}

/* ---------------- stage-level entry points ---------------- */
extern "C" int tron_grid_device(tron_plan *p, void *d_grid, const void *d_samples, int z0, int nslices, void *stream)
{
    if (!p || !p->cfg.adjoint) { set_error("tron_grid_device needs an adjoint plan"); return TRON_EINVAL; }
    if (z0 < 0 || nslices < 1 || z0 + nslices > p->nslices) { set_error("slice range [%d,%d) outside the plan's %d slices", z0, z0 + nslices, p->nslices); return TRON_EINVAL; }
    DeviceGuard guard(p->device);
    cudaStream_t s = (cudaStream_t)stream;
    GridLaunch L = make_grid_launch(p, d_samples, (float2 *)d_grid, z0, nslices);
    return launch_grid(L, s);
}

extern "C" int tron_grid_to_interleaved(tron_plan *p, void *d_dst, const void *d_grid, int nslices, void *stream)
{
    if (!p) { set_error("null plan"); return TRON_EINVAL; }
    DeviceGuard guard(p->device);
    cudaStream_t s = (cudaStream_t)stream;
    return launch_interleave((float2 *)d_dst, (const float2 *)d_grid, p->nch, p->g.nxos, nslices, s);
}

extern "C" int tron_degrid_device(tron_plan *p, void *d_samples, const void *d_grid, void *stream)
{
    if (!p || p->cfg.adjoint) { set_error("tron_degrid_device needs a forward plan"); return TRON_EINVAL; }
    DeviceGuard guard(p->device);
    cudaStream_t s = (cudaStream_t)stream;
    const tron_geometry &g = p->g;
    if (p->nch != g.nc) { set_error("tron_degrid_device does not support coil shards"); return TRON_EUNSUPPORTED; }
    int rc = launch_deinterleave(p->d_grid, (const float2 *)d_grid, p->nch, g.nxos, s);
    if (rc) return rc;
    DegridLaunch d;
    d.samples = d_samples; d.grid = p->d_grid; d.cs = p->tabs.cs_lin;
    d.n = g.nxos; d.nro = g.nro; d.npe = g.npe1work;
    d.nc_total = g.nc * g.nt; d.ch0 = 0; d.nch = p->nch;
    d.kb = p->kb; d.half_out = p->cfg.half_out;
    return run_degrid(p, d, s);
}

extern "C" int tron_plan_last_stage_ms(tron_plan *p, float ms[3])
{
    if (!p) { set_error("null plan"); return TRON_EINVAL; }
    for (int i = 0; i < 3; ++i) ms[i] = p->last_ms[i];
    return TRON_OK;
}

extern "C" int tron_plan_last_launches(const tron_plan *p) { return p ? p->last_launches : 0; }
extern "C" int tron_plan_batch_slices(const tron_plan *p) { return p ? p->batch : 0; }

/* diagnostic: copy out the per-warp cycle counts of the last gridding launch (TRON_GRID_DEBUG) */
extern "C" int tron_plan_grid_debug(tron_plan *p, long long *h_cycles, int nwarps)
{
    if (!p || !p->grid_dbg) { set_error("plan was not created with TRON_GRID_DEBUG set"); return TRON_EINVAL; }
    if (nwarps > 8 * 65536 * 8) nwarps = 8 * 65536 * 8;
    TRON_CUDA(cudaDeviceSynchronize());
    TRON_CUDA(cudaMemcpy(h_cycles, p->grid_dbg, (size_t)nwarps * sizeof(long long), cudaMemcpyDeviceToHost));
    return TRON_OK;
}

/* pinned host memory for ra_read_pinned (ra.c is plain C and does not see cudart) */
extern "C" int tron_pinned_alloc(void **p, size_t bytes)
{
    return cudaMallocHost(p, bytes) == cudaSuccess ? 0 : -1;
}
extern "C" void tron_pinned_free(void *p) { if (p) cudaFreeHost(p); }
