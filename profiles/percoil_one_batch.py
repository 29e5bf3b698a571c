"""One run of the per-coil adjoint paths on 128 cfg2 slices (for ncu launch lists):
python profiles/percoil_one_batch.py walsh|cgnr1|cgnr3|percoil"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, tron_b200 as t
mode = sys.argv[1] if len(sys.argv) > 1 else "walsh"
extra = dict(walsh=dict(coil_combine=1, walsh_npatch=1), cgnr1=dict(niter=1), cgnr3=dict(niter=3),
             percoil=dict(per_coil_out=True), sos={})[mode]
dims = [6, 1, 512, 21 * 127 + 204, 1]
p = t.Plan(t.make_config(dims, adjoint=True, golden=True, undersamp=0.4, prof_slide=21, **extra))
d_in = torch.randn(int(np.prod(dims)), 2, device="cuda")
d_out = torch.zeros(int(p.geom.shard_out_elems), 2, device="cuda")
for _ in range(2):
    p.recon_device(d_out.data_ptr(), d_in.data_ptr(), 0)
torch.cuda.synchronize()
