"""Per-warp cycle counts of one gridding launch (TRON_GRID_DEBUG): where is the critical path?"""
import ctypes as C, os, sys
os.environ['TRON_GRID_DEBUG'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, tron_b200 as t
from bench import WORKLOADS
dims, flags, desc = WORKLOADS['cfg2']
p = t.Plan(t.make_config(dims, device=0, **flags)); g = p.geom
d_in = torch.randn(int(g.shard_in_elems)*2, device='cuda')
B = 4
d_grid = torch.empty(B*g.nc*g.nxos*g.nxos*2, device='cuda')
for _ in range(3):
    p.grid_device(d_grid.data_ptr(), d_in.data_ptr(), 0, B, 0)
torch.cuda.synchronize()
nblocks = 1024 + 200
buf = np.zeros(nblocks*8, dtype=np.int64)
L = t.load_library()
L.tron_plan_grid_debug.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
assert L.tron_plan_grid_debug(p.handle, buf.ctypes.data_as(C.c_void_p), buf.size) == 0
cyc = buf.reshape(-1, 8)
blk = cyc.max(axis=1)
nz = np.nonzero(blk)[0]
print('blocks with data', len(nz), 'max warp cycles', cyc.max(), '=', cyc.max()/1.965e3, 'us')
order = np.argsort(-blk)[:12]
for b in order: print('block', b, 'warp cycles', cyc[b].tolist())
print('median block max', np.median(blk[nz]), 'sum of block max (SM-cycles)', blk.sum())
# heavy blocks are first
hb = (blk[:158]); print('heavy blocks: max', hb.max(), 'mean', hb.mean())
tb = blk[158:158+1024]; print('tile blocks: max', tb.max(), 'mean', tb.mean(), 'top5', np.sort(tb)[-5:])
