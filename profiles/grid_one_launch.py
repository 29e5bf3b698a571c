"""One workload, a few gridding launches of B slices (for ncu): python profiles/grid_one_launch.py cfg2 256 [n_launches]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, tron_b200 as t
from bench import WORKLOADS
name = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
nl = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dims, flags, desc = WORKLOADS[name]
p = t.Plan(t.make_config(dims, device=0, **flags)); g = p.geom
B = min(B, g.nz)
d_in = torch.randn(int(g.shard_in_elems) * 2, device='cuda')
d_grid = torch.empty(B * g.nc * g.nxos * g.nxos * 2, device='cuda')
st = torch.cuda.current_stream().cuda_stream
for i in range(nl):
    p.grid_device(d_grid.data_ptr(), d_in.data_ptr(), (i * B) % max(1, g.nz - B + 1), B, st)
torch.cuda.synchronize()
p.close()
