"""Device-resident timing of the forward (degridding) and adjoint paths on cfg1- and cfg5-like shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['TRON_STAGE_TIMING'] = '1'
import torch, tron_b200 as t
CASES = {
    'cfg1_fwd':  ([1, 1, 256, 256, 1], dict(adjoint=False)),
    'cfg1_adj':  ([1, 1, 512, 512, 1], dict(adjoint=True)),
    'cfg5_fwd8': ([8, 1, 1024, 1024, 1], dict(adjoint=False, kernwidth=6.0, half_in=True, half_out=True)),
    'cfg5_adj8': ([8, 1, 2048, 2048, 1], dict(adjoint=True, kernwidth=6.0, half_in=True)),
    'cfg5_fwd16': ([16, 1, 1024, 1024, 1], dict(adjoint=False, kernwidth=6.0, half_in=True, half_out=True)),
    'cfg5_adj16': ([16, 1, 2048, 2048, 1], dict(adjoint=True, kernwidth=6.0, half_in=True)),
    'cfg5_fwd32': ([32, 1, 1024, 1024, 1], dict(adjoint=False, kernwidth=6.0, half_in=True, half_out=True)),
    'cfg5_adj32': ([32, 1, 2048, 2048, 1], dict(adjoint=True, kernwidth=6.0, half_in=True)),
    'cfg5_fwd64': ([64, 1, 1024, 1024, 1], dict(adjoint=False, kernwidth=6.0, half_in=True, half_out=True)),
    'cfg5_adj64': ([64, 1, 2048, 2048, 1], dict(adjoint=True, kernwidth=6.0, half_in=True)),
}
for name in sys.argv[1:] or list(CASES):
    dims, flags = CASES[name]
    p = t.Plan(t.make_config(dims, device=0, **flags)); g = p.geom
    ib = 2 if flags.get('half_in') else 4
    ob = 2 if flags.get('half_out') else 4
    d_in = (torch.randn(int(g.shard_in_elems) * 2, device='cuda')).to(torch.float16 if ib == 2 else torch.float32)
    d_out = torch.zeros(int(g.shard_out_elems) * 2, device='cuda', dtype=torch.float16 if ob == 2 else torch.float32)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for rep in range(3):
        torch.cuda.synchronize(); e0.record()
        p.recon_device(d_out.data_ptr(), d_in.data_ptr(), 0)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    nsamp = g.nc * g.nro * g.npe1work * (g.nz if flags['adjoint'] else 1)
    print('%-11s %8.3f ms   %.2f G coil-samples/s   stages(ms) %s' % (name, best, nsamp / best / 1e6, [round(x, 3) for x in p.last_stage_ms()]), flush=True)
    p.close(); del d_in, d_out
