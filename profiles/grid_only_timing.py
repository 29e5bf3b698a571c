"""Time the gridding kernel alone (tron_grid_device) for several slices-per-launch values."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, tron_b200 as t
from bench import WORKLOADS
name = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
dims, flags, desc = WORKLOADS[name]
p = t.Plan(t.make_config(dims, device=0, **flags)); g = p.geom
d_in = torch.randn(int(g.shard_in_elems)*2, device='cuda')
for B in (4, 8, 16, 32, 64, 128):
    if B > g.nz: break
    d_grid = torch.empty(B*g.nc*g.nxos*g.nxos*2, device='cuda')
    st = torch.cuda.current_stream().cuda_stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for rep in range(3):
        torch.cuda.synchronize(); e0.record(); n=0
        for z0 in range(0, g.nz - B + 1, B):
            p.grid_device(d_grid.data_ptr(), d_in.data_ptr(), z0, B, st); n+=1
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1)/n)
    print('%s B %4d  ms/launch %.4f  us/slice %.2f' % (name, B, best, best/B*1e3), flush=True)
    del d_grid
