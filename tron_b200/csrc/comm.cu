/*
 * comm.cu -- the one collective of the hot path: coil-sharded root-sum-of-squares.
 *
 * The reference combines coils on one GPU (coilcombinesos, /root/reference/src/tron.cu:255-268, called at
 * tron.cu:764); its "multi GPU" switch is a compile-time macro without any communication (tron.h:48-49,
 * tron.cu:582-585).  When the coils of a slice are sharded over GPUs (BASELINE cfg5: 64 coils, 8 per GPU)
 * every GPU runs the whole adjoint pipeline for its coils and ends with a partial sum of squares
 * float32[nx*ny] per slice (tron_config.sos_partial, written by the last FFT pass); what is left of
 * coilcombinesos is
 *
 *        sum over GPUs  ->  ncclReduce(sum) to the root over NVLink      (4 MiB at 1024^2)
 *        sqrt           ->  on the root, same (sqrt(s), 0) pixel as tron.cu:263-264
 *
 * Two ways to get a communicator, both plain C:
 *   - one process per GPU (torchrun, MPI, ...): rank 0 calls tron_comm_unique_id(), ships the 128 bytes to
 *     the others by whatever means the launcher has, every rank calls tron_comm_create();
 *   - one process driving all GPUs (what the reference's MULTI_GPU loop would have needed):
 *     tron_comm_create_all() = ncclCommInitAll, and tron_coil_reduce_all() issues the grouped reduce.
 */
#include "tron_internal.h"

#include <dlfcn.h>
#include <nccl.h>
#include <string.h>
#include <mutex>

using namespace tronb;

struct tron_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1, device = 0;
};

/* NCCL is bound at first use, not at load time: a host process often carries its own libnccl.so.2 (torch bundles
 * one, newer than the system's), and two libraries with one SONAME cannot both be mapped -- whichever is already
 * in the process is the one to use (RTLD_NOLOAD first), otherwise the system's.  Types come from <nccl.h>. */
namespace {
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Reduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

const NcclApi *nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
        if (!h) return;
        api.handle = h;
#define BIND(name) *(void **)(&api.name) = dlsym(h, "nccl" #name)
        BIND(GetUniqueId); BIND(CommInitRank); BIND(CommInitAll); BIND(CommDestroy); BIND(Reduce);
        BIND(GroupStart); BIND(GroupEnd); BIND(GetErrorString);
#undef BIND
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommInitAll && api.CommDestroy && api.Reduce
                 && api.GroupStart && api.GroupEnd && api.GetErrorString;
    });
    return api.ok ? &api : nullptr;
}
} // namespace

#define TRON_NCCL_API(N) const NcclApi *N = nccl_api(); \
    if (!N) { set_error("libnccl.so.2 not found (or too old): the coil-sharded reduce needs NCCL"); return TRON_EUNSUPPORTED; }
#define TRON_NCCL(N, call) do { ncclResult_t r__ = (call); if (r__ != ncclSuccess) { \
        set_error("NCCL error %d (%s) in %s at %s:%d", (int)r__, N->GetErrorString(r__), #call, __FILE__, __LINE__); \
        return TRON_ECUDA; } } while (0)

namespace {
/* tron.cu:263-264: img = (sqrtf(sum), 0); sqrtf under --use_fast_math as in the reference build */
template <bool HALF>
__global__ void sos_finish_kernel(void *out, const float *__restrict__ sos, size_t npix)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
        const float2 v = make_float2(sqrtf(sos[i]), 0.f);
        if (HALF) ((__half2 *)out)[i] = __float22half2_rn(v);
        else ((float2 *)out)[i] = v;
    }
}

int finish_on_root(void *d_img, const float *d_sos, size_t npix, int half_out, cudaStream_t s)
{
    const int threads = 256;
    size_t want = (npix + threads - 1) / threads;
    const int blocks = (int)(want < 148 * 8 ? (want ? want : 1) : 148 * 8);
    if (half_out) sos_finish_kernel<true><<<blocks, threads, 0, s>>>(d_img, d_sos, npix);
    else sos_finish_kernel<false><<<blocks, threads, 0, s>>>(d_img, d_sos, npix);
    TRON_CUDA(cudaGetLastError());
    return TRON_OK;
}
} // namespace

extern "C" int tron_comm_unique_id(void *id, size_t bytes)
{
    if (!id || bytes < sizeof(ncclUniqueId)) { set_error("tron_comm_unique_id needs %zu bytes", sizeof(ncclUniqueId)); return TRON_EINVAL; }
    TRON_NCCL_API(N);
    ncclUniqueId u;
    TRON_NCCL(N, N->GetUniqueId(&u));
    memcpy(id, &u, sizeof u);
    return TRON_OK;
}

extern "C" int tron_comm_create(tron_comm **out, const void *id, size_t bytes, int rank, int nranks, int device)
{
    if (!out || !id || bytes < sizeof(ncclUniqueId) || nranks < 1 || rank < 0 || rank >= nranks) { set_error("bad communicator arguments"); return TRON_EINVAL; }
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { cudaGetLastError(); set_error("device %d of %d", device, ndev); return TRON_ENODEV; }
    TRON_NCCL_API(N);
    DeviceGuard guard(device);
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    tron_comm *c = new tron_comm();
    c->rank = rank; c->nranks = nranks; c->device = device;
    ncclResult_t r = N->CommInitRank(&c->comm, nranks, u, rank);
    if (r != ncclSuccess) { set_error("ncclCommInitRank: %s", N->GetErrorString(r)); delete c; return TRON_ECUDA; }
    *out = c;
    return TRON_OK;
}

extern "C" int tron_comm_create_all(tron_comm **out, int ndev, const int *devices)
{
    if (!out || ndev < 1 || ndev > 64) { set_error("bad communicator arguments"); return TRON_EINVAL; }
    ncclComm_t comms[64];
    int devs[64];
    TRON_NCCL_API(N);
    int prev = -1;
    cudaGetDevice(&prev);
    for (int i = 0; i < ndev; ++i) { devs[i] = devices ? devices[i] : i; out[i] = nullptr; }
    ncclResult_t r = N->CommInitAll(comms, ndev, devs);
    if (prev >= 0) cudaSetDevice(prev);
    if (r != ncclSuccess) { set_error("ncclCommInitAll: %s", N->GetErrorString(r)); return TRON_ECUDA; }
    for (int i = 0; i < ndev; ++i) {
        out[i] = new tron_comm();
        out[i]->comm = comms[i]; out[i]->rank = i; out[i]->nranks = ndev; out[i]->device = devs[i];
    }
    return TRON_OK;
}

extern "C" int tron_comm_destroy(tron_comm *c)
{
    if (!c) return TRON_OK;
    const NcclApi *N = nccl_api();
    if (c->comm && N) { DeviceGuard guard(c->device); N->CommDestroy(c->comm); }
    delete c;
    return TRON_OK;
}

extern "C" int tron_comm_rank(const tron_comm *c) { return c ? c->rank : -1; }
extern "C" int tron_comm_size(const tron_comm *c) { return c ? c->nranks : 0; }

/* One rank's part of the reduce, asynchronous on `stream`.  d_sos is reduced in place on the root. */
extern "C" int tron_coil_reduce(tron_comm *c, void *d_img, void *d_sos, size_t npix, int root, int half_out, void *stream)
{
    if (!c || !d_sos || npix == 0 || root < 0 || root >= c->nranks) { set_error("bad tron_coil_reduce arguments"); return TRON_EINVAL; }
    if (c->rank == root && !d_img) { set_error("the root needs an image buffer"); return TRON_EINVAL; }
    TRON_NCCL_API(N);
    DeviceGuard guard(c->device);
    cudaStream_t s = (cudaStream_t)stream;
    if (c->nranks > 1) TRON_NCCL(N, N->Reduce(d_sos, d_sos, npix, ncclFloat32, ncclSum, root, c->comm, s));
    if (c->rank == root) return finish_on_root(d_img, (const float *)d_sos, npix, half_out, s);
    return TRON_OK;
}

/* Single-process form: the partial sums of all `n` communicators (one per GPU) in one NCCL group. */
extern "C" int tron_coil_reduce_all(tron_comm **cs, int n, void *d_img_root, void **d_sos, size_t npix, int root,
                                    int half_out, void **streams)
{
    if (!cs || !d_sos || n < 1 || root < 0 || root >= n || !d_img_root) { set_error("bad tron_coil_reduce_all arguments"); return TRON_EINVAL; }
    TRON_NCCL_API(N);
    int prev = -1;
    cudaGetDevice(&prev);
    if (n > 1) {
        TRON_NCCL(N, N->GroupStart());
        for (int i = 0; i < n; ++i) {
            ncclResult_t r = N->Reduce(d_sos[i], d_sos[i], npix, ncclFloat32, ncclSum, root, cs[i]->comm,
                                       streams ? (cudaStream_t)streams[i] : nullptr);
            if (r != ncclSuccess) { N->GroupEnd(); set_error("ncclReduce: %s", N->GetErrorString(r)); cudaSetDevice(prev); return TRON_ECUDA; }
        }
        ncclResult_t r = N->GroupEnd();
        cudaSetDevice(prev);
        if (r != ncclSuccess) { set_error("ncclGroupEnd: %s", N->GetErrorString(r)); return TRON_ECUDA; }
    }
    cudaSetDevice(cs[root]->device);
    int rc = finish_on_root(d_img_root, (const float *)d_sos[root], npix, half_out, streams ? (cudaStream_t)streams[root] : nullptr);
    cudaSetDevice(prev);
    return rc;
}
