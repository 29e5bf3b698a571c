"""Coil-sharded root sum of squares through the library's NCCL reduce (include/tron.h: tron_comm_*,
tron_coil_reduce*; replaces coilcombinesos, /root/reference/src/tron.cu:255-268, when the coils of a slice
live on several GPUs).  The 2-GPU cases skip on a single-GPU box (run them with `gpurun --gpus 2`)."""
import ctypes as C

import numpy as np
import pytest

from util import rel_l2, synth_complex
from test_parity_gpu import flags_to_cfg, run_ref, torch_cuda

pytestmark = pytest.mark.gpu

DIMS = [8, 1, 128, 96, 1]
FLAGS = dict(adjoint=True, golden=True, undersamp=0.25, prof_slide=16)      # 5 slices of 32 spokes


def _partial_sos(t, torch, h_in, c0, c1, device):
    with torch.cuda.device(device):
        with t.Plan(flags_to_cfg(DIMS, FLAGS, coils=(c0, c1), sos_partial=True, device=device)) as p:
            sos = p.recon_host(h_in)
        return torch.from_numpy(np.ascontiguousarray(sos)).to("cuda:%d" % device)


def test_single_rank_communicator_finishes_the_partial_sum(lib, reflib_wide):
    import tron_b200 as t
    torch = torch_cuda()
    h_in = synth_complex((int(np.prod(DIMS)),), stream=700)
    want = run_ref(reflib_wide, DIMS, FLAGS, h_in)
    d_sos = _partial_sos(t, torch, h_in, 0, 2, 0) + _partial_sos(t, torch, h_in, 2, 8, 0)
    npix = d_sos.numel()
    d_img = torch.zeros(npix * 2, dtype=torch.float32, device="cuda:0")
    with t.Comm(t.comm_unique_id(), 0, 1, 0) as c:
        c.coil_reduce(d_img.data_ptr(), d_sos.data_ptr(), npix, root=0, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
    got = d_img.cpu().numpy().view(np.complex64)
    assert np.all(got.imag == 0)
    assert rel_l2(got, want) <= 1e-5, rel_l2(got, want)
    with pytest.raises(t.TronError):
        t.Comm(b"\0" * 16, 0, 1, 0)                       # short id


@pytest.mark.parametrize("root", [0, 1])
def test_two_gpu_coil_shards_reduce_to_the_one_gpu_image(lib, reflib_wide, root):
    """Single-process form (ncclCommInitAll): 4 coils per GPU, one grouped ncclReduce, sqrt on the root."""
    import tron_b200 as t
    torch = torch_cuda()
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    h_in = synth_complex((int(np.prod(DIMS)),), stream=701)
    with t.Plan(flags_to_cfg(DIMS, FLAGS, device=0)) as p:
        one_gpu = p.recon_host(h_in)
    want = run_ref(reflib_wide, DIMS, FLAGS, h_in)
    sos = [_partial_sos(t, torch, h_in, 0, 4, 0), _partial_sos(t, torch, h_in, 4, 8, 1)]
    npix = sos[0].numel()
    d_img = torch.zeros(npix * 2, dtype=torch.float32, device="cuda:%d" % root)
    for d in (0, 1):
        torch.cuda.synchronize(d)
    comms = (C.c_void_p * 2)()
    assert lib.tron_comm_create_all(comms, 2, None) == 0, lib.tron_last_error()
    try:
        assert [lib.tron_comm_rank(comms[i]) for i in range(2)] == [0, 1] and lib.tron_comm_size(comms[0]) == 2
        ptrs = (C.c_void_p * 2)(sos[0].data_ptr(), sos[1].data_ptr())
        rc = lib.tron_coil_reduce_all(comms, 2, C.c_void_p(d_img.data_ptr()), ptrs, npix, root, 0, None)
        assert rc == 0, lib.tron_last_error()
        for d in (0, 1):
            torch.cuda.synchronize(d)
    finally:
        for i in range(2):
            lib.tron_comm_destroy(comms[i])
    got = d_img.cpu().numpy().view(np.complex64)
    assert rel_l2(got, one_gpu) <= 1e-6, rel_l2(got, one_gpu)
    assert rel_l2(got, want) <= 1e-5
    assert torch.cuda.current_device() == 0, "entry points must leave the caller's current device alone"
