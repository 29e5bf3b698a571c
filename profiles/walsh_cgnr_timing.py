"""Device timings of the SURVEY 8(f) rows N3/N4 on BASELINE cfg2 shapes (run on the GPU box):
Walsh combine alone (vs the reference kernel), adjoint pipeline with -w 1, CGNR -i 3."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import tron_b200 as t  # noqa: E402


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = {}
s = torch.cuda.current_stream().cuda_stream
for nc, nimg, ns in ((6, 256, 64), (32, 256, 16), (64, 512, 4)):
    coil = torch.randn(ns, nimg, nimg, nc, 2, device="cuda") * 1e-3
    img = torch.empty(ns, nimg, nimg, 2, device="cuda")
    for npatch in (1, 3):
        ms = timed(lambda: t.coilcombine_walsh_device(img.data_ptr(), coil.data_ptr(), nimg, nc, npatch, ns, s))
        byt = ns * nimg * nimg * 8 * (nc + 1)
        out["walsh_nc%d_n%d_p%d" % (nc, nimg, npatch)] = dict(us_per_slice=1e3 * ms / ns, gbs=byt / ms / 1e6)
try:
    from oracle.oracle import RefLib
    ref = RefLib()
    c = (np.random.randn(256, 256, 6) + 1j * np.random.randn(256, 256, 6)).astype(np.complex64)
    _, ms = ref.walsh(c, 256, 6, 1, reps=5)
    out["reference_walsh_nc6_n256_p1"] = dict(us_per_slice=1e3 * ms)
except Exception as e:  # noqa: BLE001
    out["reference_walsh"] = str(e)

dims = [6, 1, 512, 21 * 127 + 204, 1]                    # 128 slices of the cfg2 geometry
flags = dict(adjoint=True, golden=True, undersamp=0.4, prof_slide=21)
d_in = torch.randn(int(np.prod(dims)), 2, device="cuda")
for name, extra in (("sos", {}), ("walsh1", dict(coil_combine=1, walsh_npatch=1)), ("cgnr1", dict(niter=1)),
                    ("cgnr3", dict(niter=3)), ("cgnr3_walsh1", dict(niter=3, coil_combine=1, walsh_npatch=1))):
    with t.Plan(t.make_config(dims, **flags, **extra)) as p:
        d_out = torch.zeros(int(p.geom.shard_out_elems), 2, device="cuda")
        ms = timed(lambda: p.recon_device(d_out.data_ptr(), d_in.data_ptr(), s), reps=3, warm=1)
        out["cfg2_128slices_" + name] = dict(ms=ms, us_per_slice=1e3 * ms / p.geom.nz, launches=p.last_launches())
print(json.dumps(out, indent=1))
