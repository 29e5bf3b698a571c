"""GPU parity tests: the CUDA path (through the C ABI) against
  (a) the unmodified reference compiled in place (oracle/_ref, same GPU), and
  (b) the CPU oracle (oracle/tron_oracle.c).

Tolerances are the north star's: relative L2 <= 1e-5 for fp32, <= 2e-3 with fp16
storage, and bit-exact sample<->cell index maps (checked with indicator probes
pushed through the reference kernels themselves).
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from util import PARITY_CASES, case_input, rel_l2, synth_complex

pytestmark = pytest.mark.gpu

TOL_F32 = 1e-5
TOL_F16 = 2e-3


def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def flags_to_cfg(dims, flags, **extra):
    import tron_b200 as t
    kw = dict(flags)
    kw.update(extra)
    return t.make_config(dims, **kw)


def run_ref(ref, dims, flags, h_in):
    ref.configure(dims, flags.get("adjoint", False), golden=flags.get("golden", False),
                  gridos=flags.get("gridos", 2.0), kernwidth=flags.get("kernwidth", 2.0),
                  undersamp=flags.get("undersamp", 1.0), prof_slide=flags.get("prof_slide", 0),
                  skip_angles=flags.get("skip_angles", 0))
    return ref.recon(h_in)


# ------------------------------------------------------------------ whole pipeline
@pytest.mark.parametrize("name", sorted(PARITY_CASES))
def test_pipeline_vs_reference(lib, reflib, name):
    import tron_b200 as t
    torch_cuda()
    dims, flags = PARITY_CASES[name]
    h_in = case_input(name)
    want = run_ref(reflib, dims, flags, h_in)
    with t.Plan(flags_to_cfg(dims, flags)) as p:
        assert [int(x) for x in p.geom.out_dims] == reflib.out_dims
        g = p.geom.as_dict()
        for k, v in reflib.geom.items():
            assert g[k] == v, k
        got = p.recon_host(h_in)
    assert got.shape == want.shape
    assert rel_l2(got, want) <= TOL_F32, rel_l2(got, want)


@pytest.mark.parametrize("name", sorted(PARITY_CASES))
def test_pipeline_vs_oracle(lib, oracle, name):
    import tron_b200 as t
    torch_cuda()
    dims, flags = PARITY_CASES[name]
    h_in = case_input(name)
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    oracle.set_trig_table(gold["ct"], gold["st"])       # SFU sin/cos of the spokes, see tron_oracle.c
    cfg = oracle.config(dims, flags.get("adjoint", False), golden=flags.get("golden", False),
                        gridos=flags.get("gridos", 2.0), kernwidth=flags.get("kernwidth", 2.0),
                        undersamp=flags.get("undersamp", 1.0), prof_slide=flags.get("prof_slide", 0),
                        skip_angles=flags.get("skip_angles", 0))
    want = oracle.recon(cfg, h_in)
    oracle.set_trig_table(None)
    with t.Plan(flags_to_cfg(dims, flags)) as p:
        got = p.recon_host(h_in)
    assert rel_l2(got, want) <= TOL_F32, rel_l2(got, want)


@pytest.mark.parametrize("dims,flags", [
    ([6, 1, 512, 60, 1], dict(adjoint=True, golden=True)),                                      # 512-point lines, RSS
    ([6, 1, 512, 90, 1], dict(adjoint=True, golden=True, undersamp=0.1, prof_slide=13)),        # sliding windows
    ([4, 1, 256, 48, 1], dict(adjoint=True, golden=True)),                                      # 256-point lines
    ([1, 1, 512, 64, 1], dict(adjoint=True)),                                                   # complex output
    ([2, 1, 512, 64, 1], dict(adjoint=True, gridos=1.0)),                                       # nothing cropped
    ([2, 1, 256, 256, 1], dict(adjoint=False, undersamp=0.1)),                                  # forward, 512-point lines
    ([1, 1, 128, 128, 1], dict(adjoint=False, golden=True, undersamp=0.2, skip_angles=3)),      # forward, 256-point lines
    ([4, 1, 256, 256, 1], dict(adjoint=False, golden=True, undersamp=0.05, gridos=1.0)),        # forward, no padding
])
@pytest.mark.parametrize("radix8", [False, True])
def test_pipeline_line_lengths_of_the_benchmarks(lib, reflib, dims, flags, radix8, monkeypatch):
    """The 256- and 512-point transforms have two implementations (two-stage 16x16 / 32x16 and
    radix-8, fft.cu); both must match the reference on the grids the benchmark configs use."""
    import tron_b200 as t
    torch_cuda()
    if radix8:
        monkeypatch.setenv("TRON_FFT_R8", "1")
    else:
        monkeypatch.setenv("TRON_FFT_P2W", "1")            # two-stage forward passes also for one or two planes
    h_in = synth_complex((int(np.prod(dims)),), stream=77)
    want = run_ref(reflib, dims, flags, h_in)
    with t.Plan(flags_to_cfg(dims, flags)) as p:
        got = p.recon_host(h_in)
    assert got.shape == want.shape
    assert rel_l2(got, want) <= TOL_F32, rel_l2(got, want)


@pytest.mark.parametrize("ring", ["2", "5", "64"])
def test_single_launch_fft_stage(lib, reflib, monkeypatch, ring):
    """TRON_FFT_FUSED: both FFT passes in one launch, the intermediate in a ring of slices that stays in
    L2; pass-B blocks wait on per-slice counters published by the pass-A blocks (fft.cu: p2w_adj_fused).
    Same numbers as the two-launch stage, bit for bit, whatever the ring length."""
    import tron_b200 as t
    torch_cuda()
    dims = [6, 1, 512, 150, 1]
    flags = dict(adjoint=True, golden=True, undersamp=0.1, prof_slide=7)          # 15 slices of 51 spokes
    h_in = synth_complex((int(np.prod(dims)),), stream=78)
    with t.Plan(flags_to_cfg(dims, flags)) as p:
        two = p.recon_host(h_in)
    monkeypatch.setenv("TRON_FFT_FUSED", "1")
    monkeypatch.setenv("TRON_FFT_RING", ring)
    with t.Plan(flags_to_cfg(dims, flags)) as p:
        one = p.recon_host(h_in)
        assert p.last_launches() < 3 * p.geom.nz
    assert np.array_equal(one, two)
    assert rel_l2(one, run_ref(reflib, dims, flags, h_in)) <= TOL_F32


def test_pipeline_more_than_six_coils(lib, reflib_wide):
    """nc > MAXCHAN needs the widened reference build (tron.h:51)."""
    import tron_b200 as t
    torch_cuda()
    dims = [16, 1, 64, 80, 1]
    flags = dict(adjoint=True, golden=True, undersamp=0.5, prof_slide=16)
    h_in = synth_complex((int(np.prod(dims)),), stream=40)
    want = run_ref(reflib_wide, dims, flags, h_in)
    with t.Plan(flags_to_cfg(dims, flags)) as p:
        got = p.recon_host(h_in)
    assert rel_l2(got, want) <= TOL_F32


def test_six_coil_chunks_combine_like_reference(lib, reflib):
    """with the stock reference, nc > 6 is run in <= 6-coil chunks and RSS-combined."""
    import tron_b200 as t
    torch_cuda()
    nc, nro, npe = 8, 64, 48
    s = synth_complex((npe, nro, nc), stream=41)
    flags = dict(adjoint=True, golden=True)
    sos = np.zeros(32 * 32, dtype=np.float64)
    for c0, c1 in ((0, 6), (6, 8)):
        chunk = np.ascontiguousarray(s[:, :, c0:c1])
        out = run_ref(reflib, [c1 - c0, 1, nro, npe, 1], flags, chunk.ravel())
        sos += np.abs(out.astype(np.complex128)) ** 2
    want = np.sqrt(sos)
    with t.Plan(flags_to_cfg([nc, 1, nro, npe, 1], flags)) as p:
        got = p.recon_host(s.ravel())
    assert np.all(got.imag == 0)
    assert rel_l2(got.real, want) <= TOL_F32


# ------------------------------------------------------------------ stage level
def _grid_mine(t, torch, samples, nchan, nro, npe, nslices=1, **flags):
    """samples: (npe_total, nro, nchan) complex64 -> (nslices, n, n, nchan) via tron_grid_device."""
    cfg = t.make_config([nchan, 1, nro, samples.shape[0], 1], adjoint=True, **flags)
    with t.Plan(cfg) as p:
        n = p.geom.nxos
        ns = p.geom.slice_end - p.geom.slice_begin
        d_s = torch.from_numpy(samples.view(np.float32).copy()).cuda()
        d_g = torch.empty((ns, nchan, n, n, 2), dtype=torch.float32, device="cuda")
        d_i = torch.empty((ns, n, n, nchan, 2), dtype=torch.float32, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        p.grid_device(d_g.data_ptr(), d_s.data_ptr(), 0, ns, st)
        p.grid_to_interleaved(d_i.data_ptr(), d_g.data_ptr(), ns, st)
        torch.cuda.synchronize()
        out = d_i.cpu().numpy().view(np.complex64)[..., 0]
        return out, p.geom.as_dict()


@pytest.mark.parametrize("golden,W,gridos,nchan", [(True, 2.0, 2.0, 2), (False, 2.0, 2.0, 6), (True, 3.0, 2.0, 1),
                                                    (True, 2.0, 1.5, 4), (True, 2.5, 2.0, 2)])
def test_grid_stage_vs_reference(lib, reflib, oracle, golden, W, gridos, nchan):
    import tron_b200 as t
    torch = torch_cuda()
    nro, npe, skip = 128, 72, 5
    s = synth_complex((npe, nro, nchan), stream=50 + nchan)
    mine, geom = _grid_mine(t, torch, s, nchan, nro, npe, golden=golden, kernwidth=W, gridos=gridos, skip_angles=skip)
    n = geom["nxos"]
    pre = oracle.precompensate(s, nchan, nro, npe)          # tron.cu:405-416 on the host
    want = reflib.grid(pre, n, nchan, nro, npe, W=W, gridos=gridos, skip=skip, golden=golden)
    assert np.array_equal(mine[0] != 0, want != 0), "tap support differs"
    assert rel_l2(mine[0], want) <= TOL_F32, rel_l2(mine[0], want)


def _indicator_probe_grid(t, torch, reflib, nro, npe, nchan, golden, W, gridos, skip, sample_ids):
    """Each channel carries ONE unit sample; the non-zero cells of that channel are exactly
    the cells that sample contributes to.  Returns the support from both implementations."""
    s = np.zeros((npe, nro, nchan), dtype=np.complex64)
    for ch, sid in enumerate(sample_ids):
        s[sid // nro, sid % nro, ch] = 1.0
    mine, geom = _grid_mine(t, torch, s, nchan, nro, npe, golden=golden, kernwidth=W, gridos=gridos, skip_angles=skip)
    want = reflib.grid(s, geom["nxos"], nchan, nro, npe, W=W, gridos=gridos, skip=skip, golden=golden)
    return mine[0], want


@pytest.mark.parametrize("golden,W,gridos", [(True, 2.0, 2.0), (False, 2.0, 2.0), (True, 2.0, 1.5), (True, 3.0, 2.0)])
def test_grid_index_map_bit_exact(lib, reflib, golden, W, gridos):
    """Sample -> cell index map, every sample of a small geometry, against the stock reference kernel."""
    import tron_b200 as t
    torch = torch_cuda()
    nro, npe, skip, nchan = 32, 12, 7, 6
    a = (2.0 - 2.0 / npe) / nro
    b = 1.0 / npe
    total = nro * npe
    ntaps = 0
    for base in range(0, total, nchan):
        ids = [min(base + c, total - 1) for c in range(nchan)]
        mine, want = _indicator_probe_grid(t, torch, reflib, nro, npe, nchan, golden, W, gridos, skip, ids)
        assert np.array_equal(mine != 0, want != 0), "index map differs for samples %s" % ids
        for ch, sid in enumerate(ids):
            sdc = a * abs((sid % nro) - nro // 2) + b       # folded density compensation
            m, w = mine[..., ch].real, want[..., ch].real * sdc
            nz = w != 0
            ntaps += int(nz.sum())
            if nz.any():
                assert np.max(np.abs(m[nz] - w[nz]) / np.abs(w[nz])) < 2e-5
    assert ntaps > 0


def test_grid_index_map_bit_exact_wide(lib, reflib_wide):
    """Same probe on a larger geometry, 64 samples per launch through the widened build."""
    import tron_b200 as t
    torch = torch_cuda()
    nro, npe, skip, nchan = 64, 40, 11, 64
    total = nro * npe
    rng = np.random.default_rng(7)
    order = rng.permutation(total)
    for base in range(0, total, nchan):
        ids = [int(order[min(base + c, total - 1)]) for c in range(nchan)]
        mine, want = _indicator_probe_grid(t, torch, reflib_wide, nro, npe, nchan, True, 2.0, 2.0, skip, ids)
        assert np.array_equal(mine != 0, want != 0), "index map differs for samples %s" % ids


@pytest.mark.parametrize("nchan,half,flags,env", [
    (6, False, dict(golden=True, undersamp=0.25, prof_slide=3), {}),                             # 4-slice groups, chains of 8 groups: sliding-window differences
    (6, False, dict(golden=True, undersamp=0.25, prof_slide=3), {"TRON_TILE_DELTA": "0"}),       # every group gridded in full
    (4, False, dict(golden=True, undersamp=0.25, prof_slide=2, skip_angles=5), {"TRON_TILE_GPER": "3"}),   # chains of 3 groups
    (6, False, dict(golden=True, undersamp=0.25, prof_slide=3), {"TRON_TILE_CAP": "1024", "TRON_TILE_DELTA": "0"}),      # many rounds per group
    (6, False, dict(golden=True, undersamp=0.25, prof_slide=3), {"TRON_TILE_CAP": "512"}),
    (6, False, dict(golden=True, undersamp=0.25, prof_slide=3), {"TRON_TILE_GPER": "3", "TRON_TILE_NEAR": "0", "TRON_TILE_DELTA": "0"}),
    (6, False, dict(golden=True, undersamp=0.25, prof_slide=3), {"TRON_TILE_GPER": "1", "TRON_TILE_NEAR": "1000"}),
    (2, False, dict(golden=True, undersamp=0.5, prof_slide=2, skip_angles=11), {"TRON_TILE_CAP": "512"}),
    (4, False, dict(golden=False, prof_slide=40, undersamp=0.4), {}),                            # linear angles: one shared table, gs = 1
    (8, False, dict(golden=True), {}),                                                           # one slice
    (6, True, dict(golden=True, undersamp=0.25, prof_slide=3), {}),                              # fp16 storage: 24-byte samples
    (2, True, dict(golden=True, undersamp=0.3, prof_slide=5), {"TRON_TILE_CAP": "256"}),
])
def test_tile_kernel_matches_l1_gather(lib, monkeypatch, nchan, half, flags, env):
    """grid_tile.cu (spoke runs staged in shared memory by cp.async.bulk, one copy pipeline per warp) visits the
    same taps as grid.cu (taps through L1) with the same weights; only cells whose spoke window wraps around the
    end of the sorted table are summed in another order (rounds go through the table front to back) -- and, when the
    sliding-window difference tables are in use, all cells of the slice groups that are built from their predecessor."""
    import tron_b200 as t
    torch = torch_cuda()
    nro, npe1 = 128, 150
    s = synth_complex((npe1, nro, nchan), stream=90 + nchan)
    raw = s.view(np.float32).astype(np.float16) if half else s.view(np.float32)
    outs = []
    for tile in (False, True):
        for k in ("TRON_NO_TILE", "TRON_TILE_CAP", "TRON_TILE_GPER", "TRON_TILE_NEAR", "TRON_TILE_DELTA"):
            monkeypatch.delenv(k, raising=False)
        monkeypatch.setenv("TRON_NO_SCATTER", "1")
        if tile:
            for k, v in env.items():
                monkeypatch.setenv(k, v)
        else:
            monkeypatch.setenv("TRON_NO_TILE", "1")
        with t.Plan(t.make_config([nchan, 1, nro, npe1, 1], adjoint=True, half_in=half, **flags)) as p:
            ns, n = p.geom.nz, p.geom.nxos
            d_s = torch.from_numpy(raw.copy()).cuda()
            d_g = torch.full((ns, nchan, n, n, 2), 7.0, dtype=torch.float32, device="cuda")
            p.grid_device(d_g.data_ptr(), d_s.data_ptr(), 0, ns, torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            outs.append(d_g.cpu().numpy())
    assert np.abs(outs[0]).max() > 0
    if env.get("TRON_TILE_DELTA") == "0" or flags.get("prof_slide", 0) in (0, 40):
        assert np.array_equal(outs[0] != 0, outs[1] != 0)
        assert rel_l2(outs[1], outs[0]) <= 1e-7
        assert np.mean(outs[0] != outs[1]) < 0.02
    else:
        # sliding-window differences: slice z+1 = slice z - leaving + entering spokes, re-anchored every chain;
        # same taps and weights, another order of additions.  A cell all of whose taps have left the window
        # again keeps their rounding residue instead of an exact zero (this sparse test geometry has such
        # cells; with the benchmark shapes every cell inside the last annulus always holds taps).
        assert rel_l2(outs[1], outs[0]) <= 5e-7
        empty = outs[0] == 0
        assert np.abs(outs[1][empty]).max() <= 1e-6 * np.abs(outs[0]).max()
        assert np.array_equal(outs[0][:1] != 0, outs[1][:1] != 0)          # the chain's first group is gridded in full


@pytest.mark.parametrize("nchan,half,nro,flags,env", [
    (6, False, 128, dict(golden=True, undersamp=0.25, prof_slide=3), {}),                                     # short-launch schedule: chains of 16 / 8 slices
    (6, False, 128, dict(golden=True, undersamp=0.25, prof_slide=3), {"TRON_SCATTER_SHORT_BELOW": "0"}),      # long-launch schedule: chains of 32 / 16
    (6, False, 128, dict(golden=True, undersamp=0.25, prof_slide=3), {"TRON_SCATTER_SHORT_BELOW": "0", "TRON_SCATTER_CHAIN": "6", "TRON_SCATTER_CHAIN_NEAR": "3"}),
    (6, False, 128, dict(golden=True, undersamp=0.25, prof_slide=3), {"TRON_SCATTER_CHAIN_SHORT": "4", "TRON_SCATTER_CHAIN_NEAR_SHORT": "2", "TRON_SCATTER_NEAR_SHORT": "0.3"}),
    (6, False, 128, dict(golden=True, undersamp=0.25, prof_slide=3), {"TRON_SCATTER_CHAIN": "1"}),            # every slice gridded in full
    (6, False, 128, dict(golden=True, undersamp=0.25, prof_slide=3), {"TRON_SCATTER_SHORT_BELOW": "0", "TRON_SCATTER_CHAIN": "5", "TRON_SCATTER_CHAIN_NEAR": "5"}),
    (6, False, 128, dict(golden=True, undersamp=0.25, prof_slide=3), {"TRON_SCATTER_NEAR_SHORT": "2"}),       # no tile on the split path
    (6, False, 128, dict(golden=True, undersamp=0.25, prof_slide=3), {"TRON_SCATTER_NEAR_SHORT": "0"}),       # every tile on the split path
    (6, False, 128, dict(golden=True, undersamp=0.25, prof_slide=3), {"TRON_SCATTER_CAP": "1536"}),           # many rounds per slice
    (4, False, 128, dict(golden=True, undersamp=0.25, prof_slide=2, skip_angles=5), {}),
    (2, False, 96, dict(golden=True, undersamp=0.5, prof_slide=2, skip_angles=11), {}),                       # 96^2 grid: 6 x 6 tiles
    (4, False, 128, dict(golden=False, prof_slide=40, undersamp=0.4), {}),                                    # linear angles: one shared table
    (6, False, 256, dict(golden=True), {}),                                                                   # one slice, 256^2
    (6, True, 128, dict(golden=True, undersamp=0.25, prof_slide=3), {}),                                      # fp16 storage: 24-byte samples
    (16, False, 128, dict(golden=True, undersamp=0.25, prof_slide=3), {}),                                    # 16 coils: half-warps share a sample, 16 x 8 tiles
    (16, False, 128, dict(golden=True, undersamp=0.25, prof_slide=3), {"TRON_SCATTER_SHORT_BELOW": "0"}),
    (16, False, 96, dict(golden=True, undersamp=0.4, prof_slide=5, skip_angles=3), {"TRON_SCATTER_NEAR_SHORT": "0"}),
    (16, True, 128, dict(golden=True, undersamp=0.25, prof_slide=3), {"TRON_SCATTER_CHAIN": "1"}),            # fp16 storage, every slice in full
    (16, False, 64, dict(golden=False), {}),                                                                  # one slice, linear angles
    (2, True, 128, dict(golden=True, undersamp=0.3, prof_slide=5), {"TRON_SCATTER_CAP": "512"}),
])
def test_scatter_kernel_matches_l1_gather(lib, monkeypatch, nchan, half, nro, flags, env):
    """grid_scatter.cu (tiles accumulated in shared memory, sample driven) applies the same taps with the same weights
    as grid.cu (one thread per cell, taps through L1); the order of the additions differs, and with sliding-window
    chains a slice is its predecessor plus / minus the spokes that entered / left."""
    import tron_b200 as t
    torch = torch_cuda()
    npe1 = 150
    s = synth_complex((npe1, nro, nchan), stream=95 + nchan)
    raw = s.view(np.float32).astype(np.float16) if half else s.view(np.float32)
    outs = []
    for scatter in (False, True):
        for k in ("TRON_NO_TILE", "TRON_NO_SCATTER", "TRON_SCATTER_CHAIN", "TRON_SCATTER_CHAIN_NEAR", "TRON_SCATTER_NEAR", "TRON_SCATTER_CAP",
                  "TRON_SCATTER_SHORT_BELOW", "TRON_SCATTER_CHAIN_SHORT", "TRON_SCATTER_CHAIN_NEAR_SHORT", "TRON_SCATTER_NEAR_SHORT"):
            monkeypatch.delenv(k, raising=False)
        if scatter:
            for k, v in env.items():
                monkeypatch.setenv(k, v)
        else:
            monkeypatch.setenv("TRON_NO_TILE", "1")
            monkeypatch.setenv("TRON_NO_SCATTER", "1")
        with t.Plan(t.make_config([nchan, 1, nro, npe1, 1], adjoint=True, half_in=half, **flags)) as p:
            ns, n = p.geom.nz, p.geom.nxos
            d_s = torch.from_numpy(raw.copy()).cuda()
            d_g = torch.full((ns, nchan, n, n, 2), 7.0, dtype=torch.float32, device="cuda")
            p.grid_device(d_g.data_ptr(), d_s.data_ptr(), 0, ns, torch.cuda.current_stream().cuda_stream)
            if ns > 3:                                     # a launch that starts and ends inside chains
                d_g[1:ns - 1] = 7.0
                p.grid_device(d_g[1:].data_ptr(), d_s.data_ptr(), 1, ns - 2, torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            outs.append(d_g.cpu().numpy())
    assert np.abs(outs[0]).max() > 0
    chains = flags.get("golden") and flags.get("prof_slide", 0) not in (0, 40) and env.get("TRON_SCATTER_CHAIN") != "1"
    per_slice = max(rel_l2(a, b) for a, b in zip(outs[1], outs[0]))
    if not chains:
        assert np.array_equal(outs[0] != 0, outs[1] != 0)
        assert per_slice <= 2e-7, per_slice
    else:
        assert per_slice <= 5e-7, per_slice
        empty = outs[0] == 0
        assert np.abs(outs[1][empty]).max() <= 1e-6 * np.abs(outs[0]).max()
        assert np.array_equal(outs[0][:1] != 0, outs[1][:1] != 0)          # a chain's first slice is gridded in full


def _degrid_mine(t, torch, grid, n_img, nchan, **flags):
    """grid: (n, n, nchan) complex64 in the reference's interleaved order -> (npe, nro, nchan)."""
    cfg = t.make_config([nchan, 1, n_img, n_img, 1], adjoint=False, **flags)
    with t.Plan(cfg) as p:
        g = p.geom
        assert g.nxos == grid.shape[0]
        d_g = torch.from_numpy(grid.view(np.float32).copy()).cuda()
        d_s = torch.empty((g.npe1work, g.nro, nchan, 2), dtype=torch.float32, device="cuda")
        p.degrid_device(d_s.data_ptr(), d_g.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        return d_s.cpu().numpy().view(np.complex64)[..., 0], g.as_dict()


@pytest.mark.parametrize("golden,W,nchan", [(False, 2.0, 1), (True, 2.0, 2), (True, 3.0, 4), (False, 2.5, 2), (True, 6.0, 2)])
def test_degrid_stage_vs_reference(lib, reflib, golden, W, nchan):
    import tron_b200 as t
    torch = torch_cuda()
    n_img, skip = 48, 3
    n = 2 * n_img
    u = synth_complex((n, n, nchan), stream=60 + nchan)
    mine, geom = _degrid_mine(t, torch, u, n_img, nchan, golden=golden, kernwidth=W, skip_angles=skip)
    want = reflib.degrid(u, n, nchan, geom["nro"], geom["npe1work"], W=W, gridos=2.0, skip=skip, golden=golden)
    assert rel_l2(mine, want) <= TOL_F32, rel_l2(mine, want)


@pytest.mark.parametrize("golden", [True, False])
def test_degrid_index_map_bit_exact(lib, reflib, golden):
    """Cell -> sample index map with indicator grids (one unit cell per channel)."""
    import tron_b200 as t
    torch = torch_cuda()
    n_img, nchan, skip = 12, 6, 2
    n = 2 * n_img
    for base in range(0, n * n, nchan):
        u = np.zeros((n, n, nchan), dtype=np.complex64)
        for ch in range(nchan):
            cid = min(base + ch, n * n - 1)
            u[cid // n, cid % n, ch] = 1.0
        mine, geom = _degrid_mine(t, torch, u, n_img, nchan, golden=golden, skip_angles=skip)
        want = reflib.degrid(u, n, nchan, geom["nro"], geom["npe1work"], W=2.0, gridos=2.0, skip=skip, golden=golden)
        assert np.array_equal(mine != 0, want != 0), "index map differs at cell %d" % base
        nz = want != 0
        if nz.any():
            assert np.max(np.abs(mine[nz] - want[nz]) / np.abs(want[nz])) < 2e-5


# ------------------------------------------------------------------ sharding, batching, legacy, CLI
def test_slice_shards_concatenate(lib):
    import tron_b200 as t
    torch_cuda()
    dims = [2, 1, 64, 200, 1]
    flags = dict(adjoint=True, golden=True, undersamp=0.25, prof_slide=5, skip_angles=2)
    h_in = synth_complex((int(np.prod(dims)),), stream=70)
    with t.Plan(flags_to_cfg(dims, flags, batch_slices=7)) as p:
        full = p.recon_host(h_in)
        nz = p.geom.nz
    parts = []
    for rank in range(3):
        lo, hi = t.shard_slices(nz, rank, 3)
        with t.Plan(flags_to_cfg(dims, flags, slices=(lo, hi), batch_slices=4)) as p:
            g = p.geom
            part = p.recon_host(h_in[int(g.shard_in_offset):int(g.shard_in_offset + g.shard_in_elems)])
            assert part.size == int(g.shard_out_elems)
            parts.append(part)
    # A slice's taps and weights do not depend on the shard; the ORDER of its additions does (slice groups and
    # their sorted spoke tables start at the shard's first slice), so shards agree to rounding, not bit for bit.
    cat = np.concatenate(parts)
    assert rel_l2(cat, full) <= 1e-6
    per_slice = [rel_l2(a, b) for a, b in zip(cat.reshape(nz, -1), full.reshape(nz, -1))]
    assert max(per_slice) <= 1e-6
    with t.Plan(flags_to_cfg(dims, flags, batch_slices=7)) as p:       # same plan shape again: bit-identical
        assert np.array_equal(p.recon_host(h_in), full)


def test_coil_shards_sum_of_squares(lib):
    import tron_b200 as t
    torch_cuda()
    dims = [8, 1, 64, 48, 1]
    flags = dict(adjoint=True, golden=True)
    h_in = synth_complex((int(np.prod(dims)),), stream=71)
    with t.Plan(flags_to_cfg(dims, flags)) as p:
        full = p.recon_host(h_in)
    sos = np.zeros(full.size, dtype=np.float32)
    for c0, c1 in ((0, 2), (2, 6), (6, 8)):
        with t.Plan(flags_to_cfg(dims, flags, coils=(c0, c1), sos_partial=True)) as p:
            sos += p.recon_host(h_in)
    assert rel_l2(np.sqrt(sos), full.real) <= 1e-6


def test_device_api_matches_host_api(lib):
    import tron_b200 as t
    torch = torch_cuda()
    dims, flags = PARITY_CASES["P2_slide"]
    h_in = case_input("P2_slide")
    with t.Plan(flags_to_cfg(dims, flags)) as p:
        want = p.recon_host(h_in)
        d_in = torch.from_numpy(h_in.view(np.float32).copy()).cuda()
        d_out = torch.zeros(want.size * 2, dtype=torch.float32, device="cuda")
        p.recon_device(d_out.data_ptr(), d_in.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert p.last_launches() > 0
        got = d_out.cpu().numpy().view(np.complex64)
    assert np.array_equal(got, want)


def test_legacy_recon_radial2d(lib):
    import tron_b200 as t
    torch_cuda()
    dims, flags = PARITY_CASES["P3_6ch"]
    h_in = case_input("P3_6ch")
    cfg = flags_to_cfg(dims, flags)
    with t.Plan(cfg) as p:
        want = p.recon_host(h_in)
    assert lib.tron_set_config(C.byref(cfg)) == 0
    got = np.zeros_like(want)
    lib.recon_radial2d(got.ctypes.data_as(C.c_void_p), h_in.ctypes.data_as(C.c_void_p))
    assert np.array_equal(got, want)


def test_legacy_kernels_match_reference(lib, reflib):
    """gridradial2d / degridradial2d compatibility kernels, odd launch shape."""
    torch = torch_cuda()
    nro, npe, nchan, n = 64, 30, 2, 64
    s = synth_complex((npe, nro, nchan), stream=72)
    want = reflib.grid(s, n, nchan, nro, npe, W=2.0, gridos=2.0, skip=3, golden=True)
    d_s = torch.from_numpy(s.view(np.float32).copy()).cuda()
    d_u = torch.zeros((n, n, nchan, 2), dtype=torch.float32, device="cuda")
    assert lib.tron_launch_gridradial2d(C.c_void_p(d_u.data_ptr()), C.c_void_p(d_s.data_ptr()), n, nchan, nro, npe,
                                        2.0, 2.0, 3, 1, 37, 96, None) == 0
    torch.cuda.synchronize()
    got = d_u.cpu().numpy().view(np.complex64)[..., 0]
    assert np.array_equal(got != 0, want != 0)
    assert rel_l2(got, want) <= TOL_F32
    u = synth_complex((n, n, nchan), stream=73)
    want = reflib.degrid(u, n, nchan, nro, npe, W=2.0, gridos=2.0, skip=3, golden=True)
    d_u = torch.from_numpy(u.view(np.float32).copy()).cuda()
    d_o = torch.zeros((npe, nro, nchan, 2), dtype=torch.float32, device="cuda")
    assert lib.tron_launch_degridradial2d(C.c_void_p(d_o.data_ptr()), C.c_void_p(d_u.data_ptr()), n, nchan, nro, npe,
                                          2.0, 2.0, 3, 1, 37, 96, None) == 0
    torch.cuda.synchronize()
    got = d_o.cpu().numpy().view(np.complex64)[..., 0]
    assert rel_l2(got, want) <= TOL_F32


def test_cli_matches_reference_binary(lib, tmp_path):
    import tron_b200 as t
    torch_cuda()
    ref_exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "tron_ref")
    if not os.path.isfile(ref_exe):
        pytest.skip("oracle/_ref/tron_ref not built")
    for name, args in (("P2_slide", ["-a", "-G", "-u", "0.25", "-d", "7", "-s", "3"]), ("P4_fwd", [])):
        dims, _ = PARITY_CASES[name]
        fin = str(tmp_path / (name + "_in.ra"))
        t.ra_write(fin, case_input(name), dims=dims)
        f_ref, f_new = str(tmp_path / "ref.ra"), str(tmp_path / "new.ra")
        subprocess.run([ref_exe] + args + [fin, f_ref], check=True, timeout=300)
        subprocess.run([t.CLI_PATH] + args + [fin, f_new], check=True, timeout=300)
        a, da, ta, ba = t.ra_read(f_ref)
        b, db, tb, bb = t.ra_read(f_new)
        assert (da, ta, ba) == (db, tb, bb)
        assert open(f_ref, "rb").read(88) == open(f_new, "rb").read(88)      # identical 5-D header bytes
        assert rel_l2(b, a) <= TOL_F32
    assert subprocess.run([t.CLI_PATH], capture_output=True).returncode == 1
    assert subprocess.run([t.CLI_PATH, "-h"], capture_output=True).returncode == 1
    assert subprocess.run([t.CLI_PATH, "-Z", "x.ra"], capture_output=True).returncode == 1


# ------------------------------------------------------------------ fp16 storage
def test_fp16_storage_adjoint(lib):
    import tron_b200 as t
    torch_cuda()
    dims, flags = PARITY_CASES["P3_6ch"]
    h_in = case_input("P3_6ch")
    with t.Plan(flags_to_cfg(dims, flags)) as p:
        want = p.recon_host(h_in)
    h16 = h_in.view(np.float32).astype(np.float16)
    with t.Plan(flags_to_cfg(dims, flags, half_in=True, half_out=True)) as p:
        got16 = p.recon_host(h16)
    got = got16.astype(np.float32).reshape(-1, 2)
    got = got[:, 0] + 1j * got[:, 1]
    assert rel_l2(got, want) <= TOL_F16, rel_l2(got, want)


def test_fp16_storage_forward(lib):
    import tron_b200 as t
    torch_cuda()
    dims, flags = PARITY_CASES["P5_fwd4"]
    h_in = case_input("P5_fwd4")
    with t.Plan(flags_to_cfg(dims, flags)) as p:
        want = p.recon_host(h_in)
    h16 = h_in.view(np.float32).astype(np.float16)
    with t.Plan(flags_to_cfg(dims, flags, half_in=True, half_out=True)) as p:
        got16 = p.recon_host(h16)
    got = got16.astype(np.float32).reshape(-1, 2)
    got = got[:, 0] + 1j * got[:, 1]
    assert rel_l2(got, want) <= TOL_F16, rel_l2(got, want)


# ------------------------------------------------------------------ full-size properties
def test_full_size_linearity_and_shard_consistency(lib):
    """BASELINE config 2 geometry (nc=6, nro=512, 204-spoke window, slide 21), a 40-slice stretch:
    per-coil linearity A(x+2y) = A(x) + 2A(y) and batch-size independence."""
    import tron_b200 as t
    torch_cuda()
    npe1 = 204 + 21 * 39
    dims = [6, 1, 512, npe1, 1]
    flags = dict(adjoint=True, golden=True, undersamp=0.4, prof_slide=21)
    x = synth_complex((int(np.prod(dims)),), stream=80)
    y = synth_complex((int(np.prod(dims)),), stream=81)
    with t.Plan(flags_to_cfg(dims, flags, per_coil_out=True, batch_slices=5)) as p:
        assert p.geom.nz == 40 and p.geom.npe1work == 204
        ax, ay = p.recon_host(x), p.recon_host(y)
        axy = p.recon_host((x + 2 * y).astype(np.complex64))
    assert rel_l2(axy, ax + 2 * ay) <= 2e-6
    with t.Plan(flags_to_cfg(dims, flags, batch_slices=16)) as p:
        rss = p.recon_host(x)
    want = np.sqrt((np.abs(ax.reshape(-1, 6).astype(np.complex128)) ** 2).sum(axis=1))
    assert rel_l2(rss.real, want) <= 1e-6


# ------------------------------------------------------------------ many channels (lanes = channels kernel)
@pytest.mark.parametrize("nc,flags", [
    (16, dict(adjoint=True, golden=True, undersamp=0.5, prof_slide=5)),          # 4-slice tap sharing, half-warp mode
    (32, dict(adjoint=True, golden=True)),                                       # one slice, 32 lanes = 32 channels
    (32, dict(adjoint=True, prof_slide=40, undersamp=0.5)),                      # linear angles, sliding, shared table
    (64, dict(adjoint=True, golden=True, kernwidth=3.0)),                        # two channels per lane
    (48, dict(adjoint=True, golden=True, gridos=1.5)),                           # ragged last channel chunk
    (32, dict(adjoint=True, golden=True, undersamp=0.25, prof_slide=3, skip_angles=9)),
    (16, dict(adjoint=True, golden=True, kernwidth=6.0)),                        # cfg5 shard on 4 GPUs: wide kernel for 16 coils
    (16, dict(adjoint=True, kernwidth=3.5, undersamp=0.5, prof_slide=20)),
    (8, dict(adjoint=True, kernwidth=6.0)),                                      # cfg5 shard on 8 GPUs: four entries per step
    (8, dict(adjoint=True, golden=True, kernwidth=3.0, undersamp=0.5, prof_slide=7)),   # ... with 4-slice tap sharing
])
def test_wide_channel_gridding_vs_reference(lib, reflib_wide, nc, flags):
    import tron_b200 as t
    torch_cuda()
    dims = [nc, 1, 64, 96, 1]
    h_in = synth_complex((int(np.prod(dims)),), stream=200 + nc)
    want = run_ref(reflib_wide, dims, flags, h_in)
    with t.Plan(flags_to_cfg(dims, flags)) as p:
        got = p.recon_host(h_in)
    assert rel_l2(got, want) <= TOL_F32, rel_l2(got, want)
    os.environ["TRON_NO_WIDE"] = "1"          # the thread-per-cell kernel must agree with the wide one
    try:
        import importlib
        with t.Plan(flags_to_cfg(dims, flags, per_coil_out=True)) as p:
            a = p.recon_host(h_in)
    finally:
        del os.environ["TRON_NO_WIDE"]
    with t.Plan(flags_to_cfg(dims, flags, per_coil_out=True)) as p:
        b = p.recon_host(h_in)
    assert rel_l2(b, a) <= 2e-6


def test_wide_channel_index_map_bit_exact(lib, reflib_wide):
    """Indicator probes through the wide kernel: 32 samples per launch, support must equal the reference's."""
    import tron_b200 as t
    torch = torch_cuda()
    nro, npe, skip, nchan = 32, 10, 4, 32
    total = nro * npe
    for base in range(0, total, nchan):
        ids = [min(base + c, total - 1) for c in range(nchan)]
        mine, want = _indicator_probe_grid(t, torch, reflib_wide, nro, npe, nchan, True, 2.0, 2.0, skip, ids)
        assert np.array_equal(mine != 0, want != 0), "index map differs for samples %s" % ids


def test_wide_channel_fp16_storage(lib):
    import tron_b200 as t
    torch_cuda()
    dims = [32, 1, 64, 80, 1]
    flags = dict(adjoint=True, golden=True, undersamp=0.5, prof_slide=16)
    h_in = synth_complex((int(np.prod(dims)),), stream=210)
    with t.Plan(flags_to_cfg(dims, flags)) as p:
        want = p.recon_host(h_in)
    h16 = h_in.view(np.float32).astype(np.float16)
    with t.Plan(flags_to_cfg(dims, flags, half_in=True)) as p:
        got = p.recon_host(h16)
    assert rel_l2(got, want) <= TOL_F16


@pytest.mark.parametrize("nc,flags", [
    (32, dict(adjoint=False, golden=True)),
    (32, dict(adjoint=False, undersamp=0.5, skip_angles=3)),                     # linear angles
    (64, dict(adjoint=False, golden=True, kernwidth=6.0)),                       # cfg5 kernel width, 2 channels per lane
    (32, dict(adjoint=False, golden=True, kernwidth=2.5, undersamp=0.3)),
    (16, dict(adjoint=False, golden=True, kernwidth=6.0)),                       # cfg5 shard on 4 GPUs: two rows per warp step
    (8, dict(adjoint=False, kernwidth=6.0, undersamp=0.7)),                      # cfg5 shard on 8 GPUs: four rows per warp step
    (16, dict(adjoint=False, golden=True, kernwidth=3.0, skip_angles=2)),
])
def test_wide_channel_degridding_vs_reference(lib, reflib, nc, flags):
    """Forward path with nc >= 32 (lanes = channels kernel) against the reference and the thread-per-sample kernel."""
    import tron_b200 as t
    torch_cuda()
    dims = [nc, 1, 48, 48, 1]
    h_in = synth_complex((int(np.prod(dims)),), stream=300 + nc)
    want = run_ref(reflib, dims, flags, h_in)
    with t.Plan(flags_to_cfg(dims, flags)) as p:
        got = p.recon_host(h_in)
    assert rel_l2(got, want) <= TOL_F32, rel_l2(got, want)
    assert np.array_equal(got != 0, want != 0)
    os.environ["TRON_NO_WIDE"] = "1"
    try:
        with t.Plan(flags_to_cfg(dims, flags)) as p:
            other = p.recon_host(h_in)
    finally:
        del os.environ["TRON_NO_WIDE"]
    assert rel_l2(got, other) <= 2e-6
    # one and two spokes per warp (the default pairs spokes only for linear angle order; forced on a golden-angle
    # order the pairs are not neighbours and the kernel must walk them spoke by spoke)
    for pair in ("0", "1"):
        os.environ["TRON_DEGRID_PAIR"] = pair
        try:
            with t.Plan(flags_to_cfg(dims, flags)) as p:
                other = p.recon_host(h_in)
        finally:
            del os.environ["TRON_DEGRID_PAIR"]
        assert rel_l2(other, want) <= TOL_F32, (pair, rel_l2(other, want))
        assert np.array_equal(other != 0, want != 0)


@pytest.mark.parametrize("nc,dims_tail,flags", [
    (32, [72, 40, 1], dict(adjoint=True, golden=True)),                  # 72 x 72 grid: the last 16 x 16 tiles are half outside
    (64, [40, 36, 1], dict(adjoint=True, kernwidth=3.0)),                # 40 x 40 grid, two channels per lane
    (8, [40, 36, 1], dict(adjoint=True, kernwidth=6.0)),
    (32, [30, 30, 1], dict(adjoint=False, undersamp=0.55)),              # 33 spokes: the last pair is half empty
    (64, [20, 20, 1], dict(adjoint=False, kernwidth=6.0, undersamp=0.825)),
])
def test_wide_kernels_on_ragged_shapes(lib, reflib, reflib_wide, nc, dims_tail, flags):
    """Grids that are not a multiple of the 16 x 16 tile (the shared-memory output transpose must not store past the
    edge) and spoke counts that are not a multiple of the spokes per warp / per block."""
    import tron_b200 as t
    torch_cuda()
    dims = [nc, 1] + dims_tail
    h_in = synth_complex((int(np.prod(dims)),), stream=900 + nc + dims_tail[0])
    want = run_ref(reflib_wide if flags["adjoint"] else reflib, dims, flags, h_in)
    with t.Plan(flags_to_cfg(dims, flags)) as p:
        got = p.recon_host(h_in)
    assert rel_l2(got, want) <= TOL_F32, rel_l2(got, want)
    if not flags["adjoint"]:
        assert np.array_equal(got != 0, want != 0)


def test_shepp_logan_round_trip(lib, reflib):
    """BASELINE config 1: forward (RUNME1 flags = defaults) then adjoint on a 256^2 Shepp-Logan phantom,
    both steps against the reference; the golden-angle pair (-G / -a -G) must give back the phantom."""
    import tron_b200 as t
    from util import shepp_logan
    torch_cuda()
    ph = shepp_logan(256)
    dims = [1, 1, 256, 256, 1]
    for golden in (False, True):
        f = dict(adjoint=False, golden=golden)
        want = run_ref(reflib, dims, f, ph.ravel())
        with t.Plan(flags_to_cfg(dims, f)) as p:
            data = p.recon_host(ph.ravel())
            assert [int(x) for x in p.geom.out_dims] == [1, 1, 512, 512, 1]
        assert rel_l2(data, want) <= TOL_F32
        a = dict(adjoint=True, golden=golden)
        want_img = run_ref(reflib, [1, 1, 512, 512, 1], a, want)
        with t.Plan(flags_to_cfg([1, 1, 512, 512, 1], a)) as p:
            img = p.recon_host(data)
        assert rel_l2(img, want_img) <= 2e-5          # two chained operators
        if golden:                                     # consistent pair: looks like the phantom
            rec = np.abs(img.reshape(256, 256)); ref_ph = np.abs(ph)
            c = np.corrcoef(rec.ravel(), ref_ph.ravel())[0, 1]
            assert c > 0.95, c


def _fuzz_cases(n, seed=20261017):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        adjoint = bool(rng.integers(0, 4) > 0)
        nc = int(rng.choice([1, 2, 4, 6]))
        golden = bool(rng.integers(0, 2))
        W = float(rng.choice([2.0, 2.0, 2.5, 3.0, 4.0]))
        gridos = float(rng.choice([2.0, 2.0, 2.0, 1.5, 1.25]))
        if adjoint:
            nro = int(rng.choice([32, 48, 64, 96, 128, 160, 256]))
            npe1work = int(rng.integers(3, 90))
            undersamp = (npe1work + 0.5) / nro
            nslide = int(rng.integers(1, 5))
            slide = int(rng.integers(1, npe1work + 1)) if nslide > 1 else 0
            npe1 = npe1work + (nslide - 1) * slide + int(rng.integers(0, max(slide, 1)))
            skip = int(rng.integers(0, 50))
            if int(nro / 2 * gridos) % 2:          # the reference's 4x4 cell remap needs nxos % 4 == 0
                gridos = 2.0
            if int(nro / 2 * gridos) % 4:
                gridos = 2.0
            dims = [nc, 1, nro, npe1, 1]
            flags = dict(adjoint=True, golden=golden, kernwidth=W, gridos=gridos, undersamp=undersamp,
                         prof_slide=slide, skip_angles=skip)
        else:
            nx = int(rng.choice([16, 24, 32, 48, 64, 100]))
            if int(nx * gridos) % 4:
                gridos = 2.0
            dims = [nc, 1, nx, nx, 1]
            flags = dict(adjoint=False, golden=golden, kernwidth=W, gridos=gridos,
                         undersamp=float(rng.choice([1.0, 0.5, 0.3])), skip_angles=int(rng.integers(0, 20)))
        out.append((i, dims, flags))
    return out


@pytest.mark.parametrize("i,dims,flags", _fuzz_cases(40), ids=lambda v: str(v) if isinstance(v, int) else None)
def test_fuzz_pipeline_vs_reference(lib, reflib, i, dims, flags):
    """Seeded random geometries (channel counts, line lengths incl. non-powers of two, kernel widths,
    oversampling ratios, sliding windows with ragged tails, skip offsets) against the reference itself."""
    import tron_b200 as t
    torch_cuda()
    h_in = synth_complex((int(np.prod(dims)),), stream=500 + i)
    want = run_ref(reflib, dims, flags, h_in)
    with t.Plan(flags_to_cfg(dims, flags)) as p:
        assert [int(x) for x in p.geom.out_dims] == reflib.out_dims
        got = p.recon_host(h_in)
    assert got.shape == want.shape
    assert rel_l2(got, want) <= TOL_F32, (dims, flags, rel_l2(got, want))
