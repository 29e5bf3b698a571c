import os, sys, json
sys.path.insert(0, '/root/repo'); sys.path.insert(0,'/root/repo/tests')
os.environ['TRON_STAGE_TIMING']='1'
import torch, tron_b200 as t
from bench import WORKLOADS
for name in sys.argv[1:] or ['cfg2']:
    dims, flags, desc = WORKLOADS[name]
    for batch in (8, 16, 32, 64):
        p = t.Plan(t.make_config(dims, device=0, batch_slices=batch, **flags)); g = p.geom
        d_in = torch.randn(int(g.shard_in_elems)*2, device='cuda'); d_out = torch.zeros(int(g.shard_out_elems)*2, device='cuda')
        for _ in range(2): p.recon_device(d_out.data_ptr(), d_in.data_ptr(), 0)
        torch.cuda.synchronize()
        ms = p.last_stage_ms()
        print(name, 'batch', batch, 'grid %.2f ms (%.2f us/slice)  fft %.2f ms (%.2f us/slice)' % (ms[0], ms[0]*1e3/g.nz, ms[1], ms[1]*1e3/g.nz), flush=True)
        p.close(); del d_in, d_out
