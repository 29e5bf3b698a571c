"""GPU parity at the BASELINE shapes themselves (task 1): the launches bench.py times are the
launches tested here, against the unmodified reference compiled in place (oracle/_ref) on the same GPU.

  cfg2  [6,1,512,204+21(k-1),1] -u 0.4 -d 21 -a -G, k = 32 / 64 / 256 slices per launch, device and host paths
  cfg3  [32,1,512,804 s,1]      -a -G -u 1.5703125 -d 804 (widened reference build, tron.h:51)
  cfg4  [16,1,256,128+21(k-1),1] -u 0.5 -d 21 -a -G
  cfg5  [64,1,256,256,1] -k 6 forward and [64,1,512,512,1] -a -k 6 adjoint, fp16 storage, 512^2 grids

Tolerances: rel-L2 <= 1e-5 (fp32), <= 2e-3 (fp16 storage), per slice as well as over the whole job.
"""
import numpy as np
import pytest

from util import rel_l2, synth_complex
from test_parity_gpu import TOL_F16, TOL_F32, flags_to_cfg, run_ref, torch_cuda

pytestmark = pytest.mark.gpu

CFG2 = dict(adjoint=True, golden=True, undersamp=0.4, prof_slide=21)


def _per_slice(got, want, nz):
    return max(rel_l2(a, b) for a, b in zip(np.asarray(got).reshape(nz, -1), np.asarray(want).reshape(nz, -1)))


def _device_recon(t, torch, p, h_in):
    d_in = torch.from_numpy(np.ascontiguousarray(h_in).view(np.float32).copy()).cuda()
    d_out = torch.zeros(int(p.geom.shard_out_elems) * 2, dtype=torch.float32, device="cuda")
    p.recon_device(d_out.data_ptr(), d_in.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return d_out.cpu().numpy().view(np.complex64)


@pytest.mark.parametrize("k", [32, 64, 256, 480])
def test_cfg2_launch_lengths_device_and_host(lib, reflib, k):
    """k slices of the cfg2 geometry in ONE device launch (the bench's are 512 + 444; from 448 slices on the plan
    uses chains of 64) and through the host pipeline's
    ramped batches; both against the reference, slice by slice, and against each other."""
    import tron_b200 as t
    torch = torch_cuda()
    dims = [6, 1, 512, 204 + 21 * (k - 1), 1]
    h_in = synth_complex((int(np.prod(dims)),), stream=600 + k)
    want = run_ref(reflib, dims, CFG2, h_in)
    with t.Plan(flags_to_cfg(dims, CFG2)) as p:
        assert p.geom.nz == k and p.geom.npe1work == 204 and p.geom.nxos == 512
        dev = _device_recon(t, torch, p, h_in)
        assert p.last_launches() == 3, "one gridding launch + two FFT passes for the whole job"
        host = p.recon_host(h_in)
    assert np.all(dev.imag == 0) and np.all(host.imag == 0)
    assert rel_l2(dev, want) <= TOL_F32 and rel_l2(host, want) <= TOL_F32
    assert _per_slice(dev, want, k) <= TOL_F32, _per_slice(dev, want, k)
    assert _per_slice(host, want, k) <= TOL_F32, _per_slice(host, want, k)
    assert _per_slice(host, dev, k) <= 1e-6, _per_slice(host, dev, k)


def test_cfg2_gridding_stage_256_slices_vs_reference_kernel(lib, reflib, oracle):
    """The launch bench.py's roofline times (tron_grid_device, 256 slices): slices picked from the start, the
    middle and the end of the launch against the reference's gridradial2d on the same windows."""
    import tron_b200 as t
    torch = torch_cuda()
    k, nc, nro, win, slide = 256, 6, 512, 204, 21
    npe1 = win + slide * (k - 1)
    s = synth_complex((npe1, nro, nc), stream=610)
    with t.Plan(t.make_config([nc, 1, nro, npe1, 1], **CFG2)) as p:
        n = p.geom.nxos
        d_s = torch.from_numpy(s.view(np.float32).copy()).cuda()
        d_g = torch.empty((k, nc, n, n, 2), dtype=torch.float32, device="cuda")
        p.grid_device(d_g.data_ptr(), d_s.data_ptr(), 0, k, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        for z in (0, 1, 31, 32, 35, 127, 128, 254, 255):
            mine = d_g[z].permute(1, 2, 0, 3).contiguous().cpu().numpy().view(np.complex64)[..., 0]
            w = np.ascontiguousarray(s[z * slide: z * slide + win])
            want = reflib.grid(oracle.precompensate(w, nc, nro, win), n, nc, nro, win, W=2.0, gridos=2.0,
                               skip=z * slide, golden=True)
            assert rel_l2(mine, want) <= TOL_F32, (z, rel_l2(mine, want))
            # support: every cell the reference fills is filled here (differences of sliding windows may leave
            # rounding residue, never a missing tap)
            assert not np.any((mine == 0) & (want != 0)), z


def test_cfg2_slice_shard_off_chain_boundaries(lib, reflib):
    """A shard that starts in the middle of the unsharded plan's chains (strong-scaling leg of bench.py)."""
    import tron_b200 as t
    torch_cuda()
    k = 96
    dims = [6, 1, 512, 204 + 21 * (k - 1), 1]
    h_in = synth_complex((int(np.prod(dims)),), stream=620)
    want = run_ref(reflib, dims, CFG2, h_in).reshape(k, -1)
    for lo, hi in ((0, 37), (37, 70), (70, 96)):
        with t.Plan(flags_to_cfg(dims, CFG2, slices=(lo, hi))) as p:
            g = p.geom
            part = p.recon_host(h_in[int(g.shard_in_offset):int(g.shard_in_offset + g.shard_in_elems)])
        assert _per_slice(part, want[lo:hi], hi - lo) <= TOL_F32


@pytest.mark.parametrize("nslices", [1, 3])
def test_cfg3_shape_vs_widened_reference(lib, reflib_wide, nslices):
    """32 coils, 512 readout, 804 spokes per slice (lanes = channels kernel)."""
    import tron_b200 as t
    torch = torch_cuda()
    dims = [32, 1, 512, 804 * nslices, 1]
    flags = dict(adjoint=True, golden=True, undersamp=1.5703125, prof_slide=804)
    h_in = synth_complex((int(np.prod(dims)),), stream=630 + nslices)
    want = run_ref(reflib_wide, dims, flags, h_in)
    with t.Plan(flags_to_cfg(dims, flags)) as p:
        assert p.geom.nz == nslices and p.geom.npe1work == 804
        host = p.recon_host(h_in)
        dev = _device_recon(t, torch, p, h_in)
    assert _per_slice(host, want, nslices) <= TOL_F32, _per_slice(host, want, nslices)
    assert _per_slice(dev, want, nslices) <= TOL_F32


def test_cfg4_stretch_vs_widened_reference(lib, reflib_wide):
    """16 coils, 128-spoke window sliding by 21 (swallowing-style), 48 frames of 128x128."""
    import tron_b200 as t
    torch = torch_cuda()
    k = 48
    dims = [16, 1, 256, 128 + 21 * (k - 1), 1]
    flags = dict(adjoint=True, golden=True, undersamp=0.5, prof_slide=21)
    h_in = synth_complex((int(np.prod(dims)),), stream=640)
    want = run_ref(reflib_wide, dims, flags, h_in)
    with t.Plan(flags_to_cfg(dims, flags)) as p:
        assert p.geom.nz == k and p.geom.npe1work == 128
        host = p.recon_host(h_in)
        dev = _device_recon(t, torch, p, h_in)
    assert _per_slice(host, want, k) <= TOL_F32, _per_slice(host, want, k)
    assert _per_slice(dev, want, k) <= TOL_F32


def _as_half_pairs(z):
    return np.ascontiguousarray(z).view(np.float32).astype(np.float16)


def _from_half_pairs(h):
    f = np.asarray(h).astype(np.float32).reshape(-1, 2)
    return (f[:, 0] + 1j * f[:, 1]).astype(np.complex64)


def test_cfg5_line_length_2048_forward_and_adjoint(lib, reflib, reflib_wide):
    """cfg5's own grid, 2048 x 2048 (1024 matrix, 2x): the 2048-point FFT passes (4 lines per block, single exchange
    buffer, channel-fastest block order from 8 channels on) against the reference, forward and adjoint, 8 coils,
    default kernel width (the reference's adjoint kernel visits every spoke for every cell: -k 6 at this size
    would take minutes)."""
    import tron_b200 as t
    torch_cuda()
    nc, nx = 8, 1024
    fdims = [nc, 1, nx, nx, 1]
    fflags = dict(adjoint=False, undersamp=0.0625)                      # 128 spokes of 2048 samples
    img = synth_complex((int(np.prod(fdims)),), stream=660)
    want_s = run_ref(reflib, fdims, fflags, img)
    with t.Plan(flags_to_cfg(fdims, fflags)) as p:
        assert p.geom.nxos == 2048 and p.geom.nro == 2048 and p.geom.npe1work == 128
        got_s = p.recon_host(img)
    assert rel_l2(got_s, want_s) <= TOL_F32, rel_l2(got_s, want_s)
    adims = [nc, 1, 2048, 128, 1]
    aflags = dict(adjoint=True, undersamp=0.0625)
    s = (want_s * (1.0 / np.abs(want_s).max())).astype(np.complex64)
    want_i = run_ref(reflib_wide, adims, aflags, s)
    with t.Plan(flags_to_cfg(adims, aflags)) as p:
        assert p.geom.nxos == 2048 and p.geom.nx == 1024 and p.geom.npe1work == 128
        got_i = p.recon_host(s)
    assert rel_l2(got_i, want_i) <= TOL_F32, rel_l2(got_i, want_i)
    with t.Plan(flags_to_cfg(adims, aflags, per_coil_out=True)) as p:   # per-coil images: pass B without the coil sum
        got_c = p.recon_host(s)
    assert rel_l2(np.sqrt((np.abs(got_c.reshape(-1, nc)) ** 2).sum(axis=1)), want_i.real) <= TOL_F32


def test_line_length_4096_forward(lib, reflib):
    """The longest power-of-two line (one line per block, 512 threads, single exchange buffer): 2048 matrix, one coil."""
    import tron_b200 as t
    torch_cuda()
    fdims = [1, 1, 2048, 2048, 1]
    fflags = dict(adjoint=False, undersamp=0.0078125)                   # 32 spokes of 4096 samples
    img = synth_complex((int(np.prod(fdims)),), stream=661)
    want_s = run_ref(reflib, fdims, fflags, img)
    with t.Plan(flags_to_cfg(fdims, fflags)) as p:
        assert p.geom.nxos == 4096 and p.geom.npe1work == 32
        got_s = p.recon_host(img)
    assert rel_l2(got_s, want_s) <= TOL_F32, rel_l2(got_s, want_s)


def test_cfg5_shape_fp16_forward_and_adjoint(lib, reflib, reflib_wide):
    """64 coils, kernel width 6, fp16 storage, 512^2 oversampled grid (a quarter of cfg5's matrix): forward and
    adjoint, each against the reference run on the same fp16-rounded input in fp32."""
    import tron_b200 as t
    torch_cuda()
    nc, nx = 64, 256
    fdims = [nc, 1, nx, nx, 1]
    fflags = dict(adjoint=False, kernwidth=6.0)
    img16 = _as_half_pairs(synth_complex((int(np.prod(fdims)),), stream=650))
    img = _from_half_pairs(img16)
    want_s = run_ref(reflib, fdims, fflags, img)
    with t.Plan(flags_to_cfg(fdims, fflags, half_in=True, half_out=True)) as p:
        assert p.geom.nxos == 512 and p.geom.nro == 512 and p.geom.npe1work == 512
        got16 = p.recon_host(img16)
    assert rel_l2(_from_half_pairs(got16), want_s) <= TOL_F16, rel_l2(_from_half_pairs(got16), want_s)
    with t.Plan(flags_to_cfg(fdims, fflags, half_in=True)) as p:       # fp32 output: only the input is rounded
        got32 = p.recon_host(img16)
    assert rel_l2(got32, want_s) <= TOL_F32, rel_l2(got32, want_s)

    adims = [nc, 1, 512, 512, 1]
    aflags = dict(adjoint=True, kernwidth=6.0)
    s16 = _as_half_pairs(want_s * (1.0 / np.abs(want_s).max()))
    s = _from_half_pairs(s16)
    want_i = run_ref(reflib_wide, adims, aflags, s)
    with t.Plan(flags_to_cfg(adims, aflags, half_in=True)) as p:
        got = p.recon_host(s16)
    assert rel_l2(got, want_i) <= TOL_F32, rel_l2(got, want_i)
    with t.Plan(flags_to_cfg(adims, aflags, half_in=True, half_out=True)) as p:
        got16 = p.recon_host(s16)
    assert rel_l2(_from_half_pairs(got16), want_i) <= TOL_F16
    # coil shards (8 coils per GPU in cfg5): partial sums of squares add up to the same image
    sos = np.zeros(want_i.size, dtype=np.float32)
    for c0 in range(0, nc, 16):
        with t.Plan(flags_to_cfg(adims, aflags, half_in=True, coils=(c0, c0 + 16), sos_partial=True)) as p:
            sos += p.recon_host(s16)
    assert rel_l2(np.sqrt(sos), want_i.real) <= TOL_F32


def test_float16_quirk_range_against_the_reference_routine(lib, reflib):
    """float_to_float16 (float16.cu:76-166) differs from IEEE round-to-nearest-even on exactly
    12 288 non-NaN inputs, all with 2^-25 < |f| < 2^-14 (sticky bits dropped for subnormal results).  The
    product's conversion runs over EVERY float of that range; wherever it leaves IEEE the reference's own
    routine (compiled into oracle/_ref) is asked, plus a dense sample of the inputs where both agree."""
    import ctypes as C
    to_half = lib.tron_float_to_half_array
    to_half.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    to_half.restype = None
    ref = reflib.lib.tronref_floatbits_to_halfbits
    one = lib.tron_floatbits_to_halfbits
    exceptions = 0
    mant = np.arange(0, 1 << 23, dtype=np.uint32)
    for e in range(101, 114):                                  # 2^-26 .. 2^-14: biased exponents 101 .. 113
        for sign in (0, 0x80000000):
            bits = mant | np.uint32(e << 23) | np.uint32(sign)
            mine = np.empty(bits.size, dtype=np.uint16)
            to_half(mine.ctypes.data, bits.ctypes.data, bits.size)
            ieee = bits.view(np.float32).astype(np.float16).view(np.uint16)
            diff = np.nonzero(mine != ieee)[0]
            exceptions += diff.size
            probe = np.unique(np.concatenate([diff, np.clip(diff + 1, 0, bits.size - 1), np.clip(diff - 1, 0, bits.size - 1),
                                              np.arange(0, bits.size, 4099)]))
            for j in probe.tolist():
                assert int(mine[j]) == ref(int(bits[j])), hex(int(bits[j]))
    assert exceptions == 12288, exceptions
    for b in (0x7f800001, 0x7fc00000, 0xffc12345, 0x7f800000, 0x477fe000, 0x477ff000, 0x47800000, 0x33000000, 0x33000001, 0, 0x80000000):
        assert one(b) == ref(b), hex(b)
