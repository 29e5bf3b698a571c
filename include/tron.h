/*
 * tron.h -- C ABI of libtron_b200: a B200-native (sm_100a) radial NUFFT engine
 * that is a drop-in for the hot path of davidssmith/TRON.
 *
 * Two layers are exported, both extern "C", plain pointers and sizes only:
 *
 *  1. The plan API (new).  The reference keeps its geometry in file-static
 *     globals that only main() sets (/root/reference/src/tron.cu:54-87,
 *     905-961), so its host entry points are not independently callable.
 *     A tron_config carries exactly what main() derives from the RA header and
 *     the command line; tron_plan_create() derives the geometry with the same
 *     integer truncations and owns every device buffer, table and stream.
 *
 *  2. The legacy symbols of the reference (tron.h:55-72, tron.cu:460-788):
 *     recon_radial2d / recon_radial_2d, tron_init, tron_shutdown,
 *     tron_nufft_adj_radial2d, tron_nufft_radial2d, and the two
 *     `extern "C" __global__` kernels gridradial2d / degridradial2d (kept as
 *     launch-configuration-agnostic compatibility shims, see tron_b200/csrc/
 *     legacy.cu).  They read a process-global configuration installed with
 *     tron_set_config(), the replacement for main()'s writes to the globals.
 *
 * Data layout at this boundary is the reference's: complex64 (or complex-half
 * when fp16 storage is selected), channel fastest, i.e. RA dims
 * [nc, nt, nro, npe1, npe2] column-major:
 *     samples[nc*(nro*pe + ro) + ch],   image[nx*row + col]   (adjoint output)
 *     image  [nc*(nx*row + col) + ch],  samples as above       (forward)
 *
 * Errors: every int-returning function returns 0 on success and a negative
 * TRON_E* code otherwise; tron_last_error() holds the message.  Nothing in
 * this library calls exit() or blocks on stdin (the reference does both,
 * tron.cu:92-100,138-147).  There is no CPU fallback: without a CUDA device
 * tron_plan_create() fails with TRON_ENODEV.
 */
#ifndef TRON_B200_TRON_H
#define TRON_B200_TRON_H

#include <stddef.h>
#include <stdint.h>

#include "ra.h"

#ifdef __cplusplus
extern "C" {
#endif

#define TRON_B200_VERSION 100

enum {
    TRON_OK = 0,
    TRON_EINVAL = -1,      /* bad configuration (e.g. odd nc != 1, nt != 1, nxos % 4) */
    TRON_ENODEV = -2,      /* no usable CUDA device */
    TRON_ECUDA = -3,       /* a CUDA runtime call failed */
    TRON_ENOMEM = -4,
    TRON_EUNSUPPORTED = -5 /* valid in the reference's CLI but not implemented (-3; -i outside its domain) */
};

/* What main() reads from the RA header and getopt (tron.cu:822-874, 905-961). */
typedef struct tron_config {
    uint64_t dims[5];       /* input RA dims */
    int   adjoint;          /* -a */
    int   golden_angle;     /* -G */
    float gridos;           /* -o  (default 2) */
    float kernwidth;        /* -k  (default 2) */
    float data_undersamp;   /* -u  (default 1) */
    int   prof_slide;       /* -d  (default 0 = npe1work) */
    int   skip_angles;      /* -s */
    int   niter;            /* -i  CGNR iterations (adjoint only, gridos 2; tron.cu:665-720, see cgnr.cu) */
    int   koosh;            /* -3  (must be 0: the reference has no 3-D kernels) */
    int   verbose;          /* -v */
    int   device;           /* -g  (-1 = leave the current device alone) */
    /* ---- extensions (all 0 = reference behaviour) ---- */
    int   half_in;          /* input payload is complex-half (fp16 storage) */
    int   half_out;         /* output payload is complex-half */
    int   slice_begin;      /* adjoint: this plan reconstructs slices [slice_begin, slice_end) */
    int   slice_end;        /*          0,0 = all nz slices */
    int   coil_begin;       /* this plan owns channels [coil_begin, coil_end) of the nc in dims[0] */
    int   coil_end;         /*          0,0 = all */
    int   sos_partial;      /* adjoint, nc>1: emit sum |z_c|^2 as float32[nx*ny] per slice (no sqrt),
                               for a cross-GPU sum when coils are sharded */
    int   batch_slices;     /* slices per kernel launch, 0 = auto */
    int   per_coil_out;     /* adjoint: skip the coil combine, emit [nx*ny*nc] per slice
                               (what tron_nufft_adj_radial2d returns in the reference) */
    int   coil_combine;     /* adjoint, nc>1: 0 = root sum of squares (tron.cu:764), 1 = adaptive Walsh
                               combine (coilcombinewalsh, tron.cu:270-302; its call at tron.cu:766 is
                               commented out in the reference) */
    int   walsh_npatch;     /* patch half-width of the Walsh combine (tron.cu:766 passes 1) */
} tron_config;

/* Derived geometry, same names as the reference's globals (tron.cu:76-79). */
typedef struct tron_geometry {
    int nc, nt, nro, npe1, npe2, npe1work;
    int nx, ny, nz, nxos, nyos;
    int prof_slide;
    int slice_begin, slice_end;       /* resolved shard */
    int coil_begin, coil_end;
    uint64_t out_dims[5];             /* RA dims of the full output (dims[0] = 1 as in tron.cu:899) */
    uint64_t in_elems;                /* complex elements of the full input */
    uint64_t out_elems;               /* complex elements of the full output */
    uint64_t shard_in_offset;         /* first input element this plan reads (full-array index) */
    uint64_t shard_in_elems;          /* contiguous input elements this plan reads */
    uint64_t shard_out_offset;        /* first output element this plan writes */
    uint64_t shard_out_elems;
} tron_geometry;

typedef struct tron_plan tron_plan;

void tron_config_defaults(tron_config *cfg);
/* host-only: works without a GPU */
int  tron_geometry_compute(const tron_config *cfg, tron_geometry *g);

int  tron_plan_create(tron_plan **plan, const tron_config *cfg);
int  tron_plan_destroy(tron_plan *plan);
int  tron_plan_geometry(const tron_plan *plan, tron_geometry *g);

/* Whole job on HOST buffers: H2D, kernels, D2H; returns when h_out is complete.
 * h_in / h_out point at the plan's shard (element shard_in_offset /
 * shard_out_offset of the full arrays).  Replaces recon_radial2d's span
 * (tron.cu:726-786) without its per-call plan/buffer creation. */
int  tron_recon_host(tron_plan *plan, void *h_out, const void *h_in);

/* Whole job on DEVICE-resident buffers, asynchronous on `stream`
 * (a cudaStream_t; NULL = the CUDA default stream, as everywhere in CUDA). */
int  tron_recon_device(tron_plan *plan, void *d_out, const void *d_in, void *stream);

/* Stage-level entry points (parity tests, profiling, roofline timing).
 * tron_grid_device: adjoint interpolation only (density compensation and the
 *   1/(nxos*npe) scale folded in) for `nslices` slices starting at shard-local
 *   slice z0; d_grid receives nslices*nc planes of nxos*nxos complex64, planar
 *   ([slice][ch][row][col]).  Replaces precompensate + gridradial2d
 *   (tron.cu:405-416, 465-536).
 * tron_grid_to_interleaved: planar -> the reference's [row][col][ch] order.
 * tron_degrid_device: forward interpolation of one oversampled grid in the
 *   reference's channel-interleaved order -> samples (tron.cu:540-577). */
int  tron_grid_device(tron_plan *plan, void *d_grid, const void *d_samples, int z0, int nslices, void *stream);
int  tron_grid_to_interleaved(tron_plan *plan, void *d_dst, const void *d_grid, int nslices, void *stream);
int  tron_degrid_device(tron_plan *plan, void *d_samples, const void *d_grid, void *stream);
/* Coil combination alone, on channel-interleaved per-coil images [nslices][nimg][nimg][nchan]
 * (complex64) -> [nslices][nimg][nimg] complex64.  No plan needed.
 * tron_coilcombine_sos_device:   coilcombinesos   (tron.cu:255-268)
 * tron_coilcombine_walsh_device: coilcombinewalsh (tron.cu:270-302, powit tron.cu:222-253); unlike the
 *   reference (MAXCHAN = 6, tron.h:51) any nchan <= 128 works. */
int  tron_coilcombine_sos_device(void *d_img, const void *d_coilimg, int nimg, int nchan, int nslices, void *stream);
int  tron_coilcombine_walsh_device(void *d_img, const void *d_coilimg, int nimg, int nchan, int npatch,
                                   int nslices, void *stream);
/* average device milliseconds of the last tron_recon_* call's stages:
 * ms[0] = gridding/degridding kernels, ms[1] = FFT passes, ms[2] = everything else on the stream */
int  tron_plan_last_stage_ms(tron_plan *plan, float ms[3]);
/* number of kernel launches issued by the last tron_recon_* call */
int  tron_plan_last_launches(const tron_plan *plan);
/* slices per launch of the device-resident pipeline (tron_recon_device); the host pipeline's batches are shorter */
int  tron_plan_batch_slices(const tron_plan *plan);
/* diagnostic (plan created with TRON_GRID_DEBUG set): per-warp cycle counts of the last gridding launch */
int  tron_plan_grid_debug(tron_plan *plan, long long *h_cycles, int nwarps);

/* ------------------------------------------------------------------ */
/* Coil-sharded root sum of squares: the one collective of the path     */
/* ------------------------------------------------------------------ */
/* The reference combines the coils of a slice on one GPU (coilcombinesos, tron.cu:255-268, called at
 * tron.cu:764) and has no communication at all (MULTI_GPU, tron.h:48-49, tron.cu:582-585).  With the coils
 * of a slice sharded over GPUs (tron_config.coil_begin/coil_end + sos_partial) every GPU ends with
 * float32[nx*ny] partial sums per slice; tron_coil_reduce is ONE ncclReduce(sum) of those to `root` over
 * NVLink followed by sqrt on the root, which leaves the same (sqrt(sum), 0) complex64 pixels (complex-half
 * with half_out) as tron.cu:263-264.  Asynchronous on `stream`; d_sos is used in place.
 *
 * One process per GPU: rank 0 fills 128 bytes with tron_comm_unique_id(), hands them to the other ranks
 * (any launcher transport), each rank calls tron_comm_create().  One process for all GPUs:
 * tron_comm_create_all() (ncclCommInitAll) and tron_coil_reduce_all() (one NCCL group). */
typedef struct tron_comm tron_comm;
#define TRON_COMM_ID_BYTES 128
int  tron_comm_unique_id(void *id, size_t bytes);
int  tron_comm_create(tron_comm **comm, const void *id, size_t bytes, int rank, int nranks, int device);
int  tron_comm_create_all(tron_comm **comms, int ndev, const int *devices);
int  tron_comm_destroy(tron_comm *comm);
int  tron_comm_rank(const tron_comm *comm);
int  tron_comm_size(const tron_comm *comm);
int  tron_coil_reduce(tron_comm *comm, void *d_img, void *d_sos, size_t npix, int root, int half_out, void *stream);
int  tron_coil_reduce_all(tron_comm **comms, int n, void *d_img_root, void **d_sos, size_t npix, int root,
                          int half_out, void **streams);

const char *tron_last_error(void);
int  tron_version(void);

/* ------------------------------------------------------------------ */
/* Legacy surface of the reference (tron.h:55-72, tron.cu:579-786)     */
/* ------------------------------------------------------------------ */
#if defined(__CUDACC__) || defined(__VECTOR_TYPES_H__)
typedef float2 tron_float2;
#else
typedef struct { float x, y; } tron_float2;
#endif

/* replaces main()'s assignment of the file-static globals */
int  tron_set_config(const tron_config *cfg);
void tron_init(void);                                                     /* tron.cu:579 */
void tron_shutdown(void);                                                 /* tron.cu:608 */
void tron_nufft_adj_radial2d(tron_float2 *d_out, tron_float2 *d_in, const int j);   /* tron.cu:623 */
void tron_nufft_radial2d(tron_float2 *d_out, tron_float2 *d_in, const int j);       /* tron.cu:639 */
void recon_radial2d(tron_float2 *h_outdata, const tron_float2 *h_indata);           /* tron.cu:726 */
void recon_radial_2d(tron_float2 *h_outdata, const tron_float2 *h_indata);          /* tron.h:71 spelling */
/* tron.cu:665: `niter` CGNR iterations on one window already on the device -> per-coil images
 * [nx*ny*nc] in d_out (the reference then coil-combines them, tron.cu:764) */
void tron_cgnr_radial2d(tron_float2 *d_out, tron_float2 *d_in, const int j, const int niter);
/* tron.cu:651-663: device-to-device copy on the plan's stream and z = y + alpha x */
void copy(tron_float2 *d_dst, tron_float2 *d_src, const size_t N, const int j);
int  tron_launch_Caxpy(void *d_z, const void *d_y, const void *d_x, float alpha, size_t N,
                       int blocks, int threads, void *stream);

/* host launchers for the two compatibility kernels below (FFI users have no <<<>>>) */
int  tron_launch_gridradial2d(void *udata, const void *nudata, int nxos, int nchan, int nro, int npe,
                              float kernwidth, float gridos, int skip_angles, int golden,
                              int blocks, int threads, void *stream);
int  tron_launch_degridradial2d(void *nudata, const void *udata, int n, int nrep, int nro, int npe,
                                float W, float gridos, int skip_angles, int golden,
                                int blocks, int threads, void *stream);

#ifdef __CUDACC__
/* Compatibility kernels with the reference's exact parameter lists (tron.h:58-67).
 * Correct for any <<<blocks, threads>>>. */
__global__ void gridradial2d(float2 *udata, const float2 *__restrict__ nudata, const int ngrid,
                             const int nchan, const int nro, const int npe, const float kernwidth,
                             const float grid_oversamp, const int skip_angles, const int flag_golden_angle);
__global__ void degridradial2d(float2 *nudata, const float2 *__restrict__ udata, const int nimg,
                               const int nchan, const int nro, const int npe, const float kernwidth,
                               const float gridos, const int skip_angles, const int flag_golden_angle);
__global__ void Caxpy(float2 *d_z, float2 *d_y, float2 *d_x, float alpha, const size_t N);   /* tron.cu:658 */
#endif

#ifdef __cplusplus
}
#endif
#endif /* TRON_B200_TRON_H */
