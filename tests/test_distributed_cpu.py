"""world_size-2 gloo test (CPU) of the multi-GPU host logic: every rank derives its slice shard from
the C library's geometry, "reconstructs" it (the CPU oracle stands in for the GPU kernels here), and
the gathered slabs equal the single-process result -- no data-path collective is needed.
Also checks the max-over-ranks timing reduction bench.py uses."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    import tron_b200 as t
    from oracle.oracle import Oracle
    from util import synth_complex
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dims = [2, 1, 32, 60, 1]
        flags = dict(adjoint=True, golden=True, undersamp=0.25, prof_slide=4, skip_angles=1)
        full = t.geometry(t.make_config(dims, **flags))
        lo, hi = t.shard_slices(full.nz, rank, world)
        g = t.geometry(t.make_config(dims, slices=(lo, hi), **flags))
        x = synth_complex((int(np.prod(dims)),), stream=90)
        shard_in = x[int(g.shard_in_offset): int(g.shard_in_offset + g.shard_in_elems)]
        # the shard is itself a valid acquisition: npe1 = its spoke count, golden index offset by lo*slide
        o = Oracle()
        nspokes = int(g.shard_in_elems) // (g.nc * g.nro)
        cfg = o.config([g.nc, 1, g.nro, nspokes, 1], True, golden=True, undersamp=0.25, prof_slide=4,
                       skip_angles=1 + lo * g.prof_slide)
        assert cfg.nz == hi - lo and cfg.npe1work == g.npe1work
        part = o.recon(cfg, shard_in)
        assert part.size == int(g.shard_out_elems)
        # gather the slabs on rank 0 (result assembly only; the compute needed no exchange)
        out = torch.zeros(int(full.out_elems) * 2, dtype=torch.float32)
        out[int(g.shard_out_offset) * 2: int(g.shard_out_offset + g.shard_out_elems) * 2] = \
            torch.from_numpy(part.view(np.float32).copy())
        dist.reduce(out, dst=0, op=dist.ReduceOp.SUM)
        tmax = torch.tensor([1.0 + rank], dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        assert float(tmax) == float(world)
        if rank == 0:
            want = o.recon(o.config(dims, True, golden=True, undersamp=0.25, prof_slide=4, skip_angles=1), x)
            q.put(bool(np.array_equal(out.numpy().view(np.complex64), want)))
    finally:
        dist.destroy_process_group()


def test_two_rank_slice_sharding_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert q.get(timeout=10) is True


def test_coil_shard_bookkeeping(lib):
    import tron_b200 as t
    dims = [64, 1, 2048, 2048, 1]
    for world in (2, 4, 8):
        per = 64 // world
        seen = []
        for rank in range(world):
            g = t.geometry(t.make_config(dims, adjoint=True, kernwidth=6.0, coils=(rank * per, (rank + 1) * per),
                                         sos_partial=True))
            seen += list(range(g.coil_begin, g.coil_end))
            assert g.nxos == 2048 and g.nx == 1024
        assert seen == list(range(64))


def _coil_worker(rank, world, port, q):
    """Coil-sharded root sum of squares, host protocol on CPU: every rank forms the partial sum of squares of ITS
    coils (the CPU oracle's per-coil adjoint stands in for the GPU pipeline with sos_partial), one reduce(sum)
    to rank 0 -- the collective tron_coil_reduce issues as ncclReduce -- and sqrt there equal the one-process
    coilcombinesos image (tron.cu:255-268)."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    import tron_b200 as t
    from oracle.oracle import Oracle
    from util import synth_complex
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        nc, nro, npe = 4, 32, 24
        dims = [nc, 1, nro, npe, 1]
        per = nc // world
        g = t.geometry(t.make_config(dims, adjoint=True, golden=True, coils=(rank * per, (rank + 1) * per), sos_partial=True))
        assert (g.coil_begin, g.coil_end) == (rank * per, (rank + 1) * per)
        assert int(g.shard_out_elems) == g.nx * g.ny                       # one float per pixel and slice
        o = Oracle()
        x = synth_complex((npe, nro, nc), stream=91)
        mine = np.ascontiguousarray(x[:, :, g.coil_begin:g.coil_end])
        cfg = o.config([per, 1, nro, npe, 1], True, golden=True)
        coil_imgs = o.adj_coils(cfg, mine)                                  # (nx, nx, per) per-coil images
        sos = torch.from_numpy((np.abs(coil_imgs.astype(np.complex128)) ** 2).sum(axis=2).astype(np.float32).ravel())
        dist.reduce(sos, dst=0, op=dist.ReduceOp.SUM)                       # the one collective of the path
        if rank == 0:
            got = np.sqrt(sos.numpy())
            want = o.recon(o.config(dims, True, golden=True), x.ravel()).real
            err = float(np.linalg.norm(got - want) / np.linalg.norm(want))
            q.put(err)
    finally:
        dist.destroy_process_group()


def test_two_rank_coil_sharded_sum_of_squares_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_coil_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert q.get(timeout=10) <= 1e-6
