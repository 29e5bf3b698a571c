"""One gridding launch of B slices (for ncu captures): python profiles/grid_one_launch.py cfg2 64"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, tron_b200 as t
from bench import WORKLOADS
name = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dims, flags, desc = WORKLOADS[name]
p = t.Plan(t.make_config(dims, device=0, **flags)); g = p.geom
d_in = torch.randn(int(g.shard_in_elems)*2, device='cuda')
d_grid = torch.empty(B*g.nc*g.nxos*g.nxos*2, device='cuda')
st = torch.cuda.current_stream().cuda_stream
for rep in range(3):
    p.grid_device(d_grid.data_ptr(), d_in.data_ptr(), 64, B, st)
torch.cuda.synchronize()
