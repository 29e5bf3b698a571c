"""The cold span (legacy recon_radial2d = plan create + recon + destroy per call) on cfg2, stage by stage.
python profiles/cold_span.py [after_ref]"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch, tron_b200 as t
from bench import WORKLOADS, make_input
dims, flags, _ = WORKLOADS["cfg2"]
torch.cuda.init(); torch.zeros(1, device="cuda")
cfg = t.make_config(dims, device=0, **flags)
g = t.geometry(cfg)
d_in, h_in = make_input(torch, int(g.in_elems), 0)
h_out = torch.zeros(int(g.out_elems) * 2, dtype=torch.float32, pin_memory=True)
L = t.load_library()
L.tron_set_config(C.byref(cfg))
if len(sys.argv) > 1:
    from oracle.oracle import RefLib
    ref = RefLib()
    ref.configure([6, 1, 512, 204, 1], True, golden=True, undersamp=0.5)
    ref.recon(h_in[:6 * 512 * 204 * 2].numpy().view(np.complex64))
os.environ["TRON_PLAN_TRACE"] = "1"
for rep in range(5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    L.recon_radial2d(C.c_void_p(h_out.data_ptr()), C.c_void_p(h_in.data_ptr()))
    print("cold call %d: %.2f ms" % (rep, (time.perf_counter() - t0) * 1e3), file=sys.stderr)
for rep in range(3):
    t0 = time.perf_counter(); p = t.Plan(cfg); t1 = time.perf_counter()
    p.recon_host_ptr(h_out.data_ptr(), h_in.data_ptr()); t2 = time.perf_counter()
    p.recon_host_ptr(h_out.data_ptr(), h_in.data_ptr()); t3 = time.perf_counter()
    p.close(); t4 = time.perf_counter()
    print("create %.2f  first recon %.2f  second recon %.2f  destroy %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3), file=sys.stderr)
