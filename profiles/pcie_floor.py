"""PCIe floor of the end-to-end cfg2 step: 498 MB host->device and 501 MB device->host, pinned, on two
streams at once (what tron_recon_host overlaps with the kernels)."""
import json
import time

import torch

n_in, n_out = 497885184, 501219328
h_in = torch.empty(n_in, dtype=torch.uint8, pin_memory=True); d_in = torch.empty(n_in, dtype=torch.uint8, device="cuda")
h_out = torch.empty(n_out, dtype=torch.uint8, pin_memory=True); d_out = torch.empty(n_out, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
res = {}
for name, do_in, do_out in (("h2d_only", 1, 0), ("d2h_only", 0, 1), ("both", 1, 1)):
    for rep in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        if do_in:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if do_out:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    res[name] = {"ms": dt * 1e3, "GBps_in": do_in * n_in / dt / 1e9, "GBps_out": do_out * n_out / dt / 1e9}
print(json.dumps(res))
