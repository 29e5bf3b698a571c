"""Does write-combined pinned memory speed up the host->device leg? (probe; not used by bench.py)"""
import ctypes as C, json, time
import torch
rt = C.CDLL("libcudart.so.12")
n = 497885184
res = {}
for name, flags in (("pinned", 0), ("write_combined", 4)):
    p = C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(p), C.c_size_t(n), C.c_uint(flags)) == 0
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    for rep in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        rt.cudaMemcpyAsync(C.c_void_p(d.data_ptr()), p, C.c_size_t(n), C.c_int(1), C.c_void_p(s))
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    res[name] = {"ms": dt * 1e3, "GBps": n / dt / 1e9}
    rt.cudaFreeHost(p)
print(json.dumps(res))
