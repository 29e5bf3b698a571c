"""Gridding kernel variants on one workload: us/slice per variant (tron_grid_device, CUDA events) and whether the
tile kernel (grid_tile.cu) reproduces the L1-gather kernel (grid.cu) bit for bit.

    python profiles/grid_variants.py [cfg2] [B]
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, tron_b200 as t
from bench import WORKLOADS

name = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dims, flags, desc = WORKLOADS[name]
KNOBS = ('TRON_NO_TILE', 'TRON_TILE_GPER', 'TRON_TILE_CAP', 'TRON_TILE_NEAR', 'TRON_TILE_MB', 'TRON_TILE_DELTA',
         'TRON_NO_SCATTER', 'TRON_SCATTER_CHAIN', 'TRON_SCATTER_CHAIN_NEAR', 'TRON_SCATTER_NEAR', 'TRON_SCATTER_CAP',
         'TRON_SCATTER_SHORT_BELOW', 'TRON_SCATTER_CHAIN_SHORT', 'TRON_SCATTER_CHAIN_NEAR_SHORT', 'TRON_SCATTER_NEAR_SHORT')

def run(env, check=None):
    for k in KNOBS:
        os.environ.pop(k, None)
    os.environ.update(env)
    p = t.Plan(t.make_config(dims, device=0, **flags)); g = p.geom
    b = min(B, g.nz)
    gen = torch.Generator(device='cuda'); gen.manual_seed(7)
    d_in = torch.randn(int(g.shard_in_elems) * 2, device='cuda', generator=gen)
    d_grid = torch.zeros(b * g.nc * g.nxos * g.nxos * 2, device='cuda')
    st = torch.cuda.current_stream().cuda_stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for rep in range(4):
        torch.cuda.synchronize(); e0.record(); n = 0
        for z0 in range(0, g.nz - b + 1, b):
            p.grid_device(d_grid.data_ptr(), d_in.data_ptr(), z0, b, st); n += 1
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n)
    p.grid_device(d_grid.data_ptr(), d_in.data_ptr(), 0, b, st)
    torch.cuda.synchronize()
    out = d_grid.clone()
    p.close()
    same = None
    if check is not None:
        same = bool(torch.equal(out, check))
        if not same:
            d = (out.double() - check.double())
            same = 'rel %.3e, %d differ' % (float(d.norm() / check.double().norm()), int((out != check).sum()))
    bytes_per_slice = 8 * g.nc * (g.nro * g.npe1work + g.nxos * g.nxos)
    print(json.dumps({'env': env, 'B': b, 'ms_per_launch': best, 'us_per_slice': best / b * 1e3,
                      'GBps': bytes_per_slice * b / best / 1e6, 'same_as_l1_kernel': same}), flush=True)
    return out

ref = run({'TRON_NO_TILE': '1', 'TRON_NO_SCATTER': '1'})
variants = [dict(v.split('=') for v in a.split(',') if v) for a in sys.argv[3:]] or [
    {}, {'TRON_TILE_GPER': '4'}, {'TRON_TILE_GPER': '16'}, {'TRON_TILE_CAP': '8192'}, {'TRON_TILE_CAP': '16384'},
    {'TRON_TILE_NEAR': '16'}, {'TRON_TILE_NEAR': '64'}]
for v in variants:
    run(v, ref)
