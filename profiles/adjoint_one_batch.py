"""One adjoint batch through the plan (for ncu captures): python profiles/adjoint_one_batch.py cfg2 64"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, tron_b200 as t
from bench import WORKLOADS
name = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dims, flags, desc = WORKLOADS[name]
dims = list(dims)
p0 = t.Plan(t.make_config(dims, device=0, **flags)); g0 = p0.geom
# a shard of 2*B slices keeps the capture short
cfg = t.make_config(dims, device=0, batch_slices=B, slices=(0, min(2 * B, g0.nz)), **flags)
p = t.Plan(cfg); g = p.geom
d_in = torch.randn(int(g.shard_in_elems)*2, device='cuda'); d_out = torch.zeros(int(g.shard_out_elems)*2, device='cuda')
for _ in range(2): p.recon_device(d_out.data_ptr(), d_in.data_ptr(), 0)
torch.cuda.synchronize()
