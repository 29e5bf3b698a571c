"""The device-resident step of a workload, twice (for an ncu launch list): python profiles/device_step.py [cfg2]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, tron_b200 as t
from bench import WORKLOADS
name = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
dims, flags, desc = WORKLOADS[name]
p = t.Plan(t.make_config(dims, device=0, **flags)); g = p.geom
d_in = torch.randn(int(g.shard_in_elems) * 2, device='cuda')
d_out = torch.zeros(int(g.shard_out_elems) * 2, device='cuda')
for _ in range(2):
    p.recon_device(d_out.data_ptr(), d_in.data_ptr(), torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print("slices per launch", p.batch_slices(), "launches per step", p.last_launches())
p.close()
